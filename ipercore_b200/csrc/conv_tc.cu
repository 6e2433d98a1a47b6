// tcgen05 / TMEM implicit-GEMM convolution with TMA im2col tile loads — the conv stack of the AttLWB-SPADE
// generator (attlwb_spade_resunet.py: Encoder :255-288, ResidualBlock :14-25, SPADE :52-99, SelfAttentionLWB 1x1
// projections :202-204, SkipDecoder :316-357, heads :605-613).  sm_100a only.
//
// GEMM view: D[M = 128 output pixels, N = Cout block] += A[M, K] * B[N, K]^T with K = (tap, cin) walked in
// 64-channel steps.  A is never materialised: for every (tap, 64-channel chunk) one TMA tile load fetches the
// spatial patch of the NHWC activation tensor shifted by the tap offset (out-of-bounds = zero fill = padding)
// straight into the 128-byte-swizzled K-major layout tcgen05 consumes.  Stride-2 convolutions use a 5-D view
// (2*pitch, W/2, 2, H/2, N) of the same tensor so that each tap is again a dense box; the transposed 4x4/s2
// convolution is four 2x2 phase convolutions with interleaved stores.
//
// The two 5x5 heads (64 -> 3 tanh, 64 -> 1 sigmoid) would be smem-bandwidth bound as a 25-tap N=16 GEMM, so they run
// as IPER_CONV_ROW5: K walks only the 5 vertical taps (5 TMA loads per tile), N = 32 holds (dx, out) pairs
// (D[p,(dx,o)] = sum_dy,c X[y+dy-2, p, c] W[o,c,dy,dx]) and the epilogue finishes the horizontal taps with a shift-add
// out[x,o] = sum_dx D[x+dx-2,(dx,o)] through shared memory; tiles are 128-pixel row segments overlapping by 4.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   : TMA producer (one lane)            smem ring of STAGES x {A planes, B planes}
//   warp 1   : TMEM allocator + MMA issuer (one lane issues tcgen05.mma, tcgen05.commit frees ring slots)
//   warps 2-5: epilogue — tcgen05.ld the fp32 accumulator (thread = pixel row), fused bias / ReLU / residual /
//              SPADE(instance-norm) / heads+composite, stores; double-buffered accumulators overlap tile i's
//              epilogue with tile i+1's MMAs.
// Split-fp16 mode (NS = 2): activations and weights carry hi and lo fp16 planes; each K step issues
// lo*hi + hi*lo + hi*hi into the same fp32 accumulator (≈ 22-bit operands) to meet the 1e-3 fp32 parity target.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "iper_b200.h"

namespace iper {

constexpr int GEMM_THREADS = 192;
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;               // fp16 elements = 128 bytes = one swizzle row
constexpr int A_TILE_BYTES = BLOCK_M * 128;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 196 * 1024;   // ring buffer budget; keeps one CTA per SM (TMEM is per-CTA 512 cols)

struct alignas(64) GemmArgs {
    CUtensorMap mapA[2];
    CUtensorMap mapB[2];
    int mode, ksize;
    int N, Ho, Wo;           // grid the M tiles walk (conv: output grid; convT: input grid)
    int oH, oW;              // stored output spatial dims
    int tw, th, tn;          // patch = tw x th x tn pixels = 128
    int tiles_x, tiles_y, tiles_nb, m_tiles, n_tiles, phases, total_tiles;
    int cin_chunks, num_k;
    int a_coff, a_pitch;
    int rows;
    int epi, relu;
    const float* bias;
    void* out; int out_planes; long long out_plane_stride; int out_pitch, out_coff;
    const __half* x; int x_planes; long long x_plane_stride; int x_pitch, x_coff;
    const float* mean_rstd; int spade_C;
    const float* bg; long long bg_batch_stride; float* img; float* mask; float* pred;
};

template <int BN, int NS>
struct Cfg {
    static constexpr int B_TILE_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = NS * (A_TILE_BYTES + B_TILE_BYTES);
    static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) < MAX_STAGES ? (SMEM_BUDGET / STAGE_BYTES) : MAX_STAGES;
    static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
    static_assert(STAGES >= 2, "need at least a double-buffered ring");
};

struct TileCoord {
    int phase, n_tile, pn0, py0, px0;
};
IPER_DEVINL TileCoord decode_tile(const GemmArgs& a, int tile) {
    TileCoord t;
    t.n_tile = tile % a.n_tiles;
    int r = tile / a.n_tiles;
    const int m = r % a.m_tiles;
    t.phase = r / a.m_tiles;
    t.px0 = (a.mode == IPER_CONV_ROW5) ? (m % a.tiles_x) * (BLOCK_M - 4) - 2 : (m % a.tiles_x) * a.tw;
    const int r2 = m / a.tiles_x;
    t.py0 = (r2 % a.tiles_y) * a.th;
    t.pn0 = (r2 / a.tiles_y) * a.tn;
    return t;
}

IPER_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// store 32 consecutive channels of one pixel as fp16 planes (hi, optional lo); 64 B contiguous per plane
IPER_DEVINL void store_planes32(const GemmArgs& a, size_t elem_off, const float (&v)[32]) {
    __half* base = reinterpret_cast<__half*>(a.out) + elem_off;
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        __half h0, l0, h1, l1;
        split_half(v[2 * i], h0, l0);
        split_half(v[2 * i + 1], h1, l1);
        hi[i] = pack_half2(h0, h1);
        lo[i] = pack_half2(l0, l1);
    }
    uint4* p0 = reinterpret_cast<uint4*>(base);
#pragma unroll
    for (int i = 0; i < 4; i++) p0[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
    if (a.out_planes > 1) {
        uint4* p1 = reinterpret_cast<uint4*>(base + a.out_plane_stride);
#pragma unroll
        for (int i = 0; i < 4; i++) p1[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
    }
}

// load 32 consecutive channels of one pixel from fp16 planes as fp32 (hi + lo)
IPER_DEVINL void load_planes32(const __half* x, int planes, long long plane_stride, size_t elem_off, float (&v)[32]) {
    const uint4* p0 = reinterpret_cast<const uint4*>(x + elem_off);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint4 u = __ldg(p0 + i);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
            v[8 * i + 2 * j] = f.x;
            v[8 * i + 2 * j + 1] = f.y;
        }
    }
    if (planes > 1) {
        const uint4* p1 = reinterpret_cast<const uint4*>(x + plane_stride + elem_off);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 u = __ldg(p1 + i);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
                v[8 * i + 2 * j] += f.x;
                v[8 * i + 2 * j + 1] += f.y;
            }
        }
    }
}

template <int BN, int NS>
__global__ void __launch_bounds__(GEMM_THREADS, 1) conv_gemm_kernel(const __grid_constant__ GemmArgs a) {
    using C = Cfg<BN, NS>;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 1024-byte aligned ring buffer (SWIZZLE_128B atoms are 1024 B)
    const uint32_t ring = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* ring_ptr = smem_dyn + (ring - smem_u32(smem_dyn));
    auto sA = [&](int stage, int p) -> uint8_t* { return ring_ptr + stage * C::STAGE_BYTES + p * A_TILE_BYTES; };
    auto sB = [&](int stage, int p) -> uint8_t* {
        return ring_ptr + stage * C::STAGE_BYTES + NS * A_TILE_BYTES + p * C::B_TILE_BYTES;
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        for (int p = 0; p < NS; p++) { tma_prefetch_desc(&a.mapA[p]); tma_prefetch_desc(&a.mapB[p]); }
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
                const TileCoord t = decode_tile(a, tile);
                const int brow = t.phase * a.rows + t.n_tile * BN;
                for (int kb = 0; kb < a.num_k; kb++) {
                    const int tap = kb / a.cin_chunks, cc = kb - tap * a.cin_chunks;
                    mbar_wait(&empty_bar[stage], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                    if (a.mode == IPER_CONV_S2) {
                        // input pixel = 2*o - 1 + d : d=0 -> (parity 1, o-1), d=1 -> (parity 0, o), d=2 -> (parity 1, o)
                        const int dy = tap / 3, dx = tap - 3 * dy;
                        const int py = (dy != 1), sy = (dy == 0) ? -1 : 0;
                        const int px = (dx != 1), sx = (dx == 0) ? -1 : 0;
                        const int c0 = px * a.a_pitch + a.a_coff + cc * BLOCK_K;
                        for (int p = 0; p < NS; p++)
                            tma_load_5d(sA(stage, p), &a.mapA[p], &full_bar[stage], c0, t.px0 + sx, py, t.py0 + sy, t.pn0);
                    } else {
                        int oy, ox;
                        if (a.mode == IPER_CONV_ROW5) {
                            oy = tap - 2; ox = 0;
                        } else if (a.mode == IPER_CONV_S1) {
                            const int dy = tap / a.ksize, dx = tap - dy * a.ksize;
                            oy = dy - a.ksize / 2; ox = dx - a.ksize / 2;
                        } else {  // transposed 4x4 s2 p1, phase (py,px), tap (ta,tb): see pack order in generator.py
                            const int py = t.phase >> 1, px = t.phase & 1;
                            const int ta = tap >> 1, tb = tap & 1;
                            oy = py == 0 ? (ta == 0 ? 0 : -1) : (ta == 0 ? 1 : 0);
                            ox = px == 0 ? (tb == 0 ? 0 : -1) : (tb == 0 ? 1 : 0);
                        }
                        const int c0 = a.a_coff + cc * BLOCK_K;
                        for (int p = 0; p < NS; p++)
                            tma_load_4d(sA(stage, p), &a.mapA[p], &full_bar[stage], c0, t.px0 + ox, t.py0 + oy, t.pn0);
                    }
                    for (int p = 0; p < NS; p++)
                        tma_load_2d(sB(stage, p), &a.mapB[p], &full_bar[stage], kb * BLOCK_K, brow);
                    if (++stage == C::STAGES) { stage = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        constexpr uint32_t idesc = umma_idesc_f16(BLOCK_M, BN);
        int stage = 0; uint32_t ph = 0; int it = 0;
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, it++) {
            const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
            mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < a.num_k; kb++) {
                mbar_wait(&full_bar[stage], ph);
                tc_fence_after();
                if (elect_one()) {
                    uint32_t first = (kb == 0) ? 0u : 1u;
                    // (A plane, B plane): small cross terms first, hi*hi last
                    constexpr int NPAIR = (NS == 2) ? 3 : 1;
                    const int pa[3] = {NS == 2 ? 1 : 0, 0, 0};
                    const int pb[3] = {0, NS == 2 ? 1 : 0, 0};
#pragma unroll
                    for (int q = 0; q < NPAIR; q++) {
                        const uint32_t abase = smem_u32(sA(stage, pa[q])), bbase = smem_u32(sB(stage, pb[q]));
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; k++) {
                            umma_f16(d_tmem, umma_desc_sw128(abase + k * 32), umma_desc_sw128(bbase + k * 32), idesc,
                                     first);
                            first = 1u;
                        }
                    }
                    umma_commit(&empty_bar[stage]);                        // ring slot reusable once MMAs retire
                    if (kb == a.num_k - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator ready for the epilogue
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; ph ^= 1; }
            }
        }
    } else {
        // =========================== epilogue (warps 2..5) ===========================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // accumulator row = pixel index inside the patch
        const int tx = row % a.tw, ty = (row / a.tw) % a.th, tni = row / (a.tw * a.th);
        int it = 0;
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, it++) {
            const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
            const TileCoord t = decode_tile(a, tile);
            mbar_wait(&tmem_full_bar[acc], acc_ph);
            tc_fence_after();
            const int n = t.pn0 + tni, y = t.py0 + ty, xx = t.px0 + tx;
            const bool valid = (n < a.N) && (y < a.Ho) && (xx < a.Wo);
            int oy = y, ox = xx;
            if (a.mode == IPER_CONVT_4S2) { oy = 2 * y + (t.phase >> 1); ox = 2 * xx + (t.phase & 1); }
            const size_t opix = ((size_t)n * a.oH + oy) * a.oW + ox;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;

            if (a.epi == IPER_EPI_HEADS) {
                if constexpr (BN == 32) {
                    __shared__ float s_ex[BLOCK_M * 21];          // D[row][dx*4+o], 20 used columns (+1 pad)
                    uint32_t r[32];
                    tmem_ld32(taddr, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 20; i++) s_ex[row * 21 + i] = __uint_as_float(r[i]);
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const int xo = t.px0 + row;
                    if (row >= 2 && row < BLOCK_M - 2 && xo >= 0 && xo < a.Wo && n < a.N && y < a.Ho) {
                        float o4[4];
#pragma unroll
                        for (int o = 0; o < 4; o++) {
                            float acc4 = 0.f;
#pragma unroll
                            for (int dx = 0; dx < 5; dx++) acc4 += s_ex[(row + dx - 2) * 21 + dx * 4 + o];
                            o4[o] = acc4;
                        }
                        const size_t hw = (size_t)a.oH * a.oW, p = (size_t)y * a.oW + xo;
                        const float m = 1.f / (1.f + expf(-o4[3]));
                        if (a.mask) a.mask[(size_t)n * hw + p] = m;
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const float v = tanhf(o4[c]);
                            if (a.img) a.img[((size_t)n * 3 + c) * hw + p] = v;
                            if (a.pred) {
                                const float bgv = a.bg[(size_t)n * a.bg_batch_stride + c * hw + p];
                                a.pred[((size_t)n * 3 + c) * hw + p] = m * bgv + (1.f - m) * v;   // imitator.py:393
                            }
                        }
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");   // s_ex is reused by the next tile
                }
            } else if (a.epi == IPER_EPI_SPADE) {
                constexpr int CB = BN / 2;      // channels per tile: columns [0,CB) gamma, [CB,2CB) beta
#pragma unroll 1
                for (int j = 0; j < CB / 32; j++) {
                    uint32_t rg[32], rb[32];
                    tmem_ld32(taddr + j * 32, rg);
                    tmem_ld32(taddr + CB + j * 32, rb);
                    tmem_ld_wait();
                    if (valid) {
                        const int c0 = t.n_tile * CB + j * 32;
                        float xv[32], o[32];
                        load_planes32(a.x, a.x_planes, a.x_plane_stride, opix * a.x_pitch + a.x_coff + c0, xv);
                        const float* mr = a.mean_rstd + ((size_t)n * a.spade_C + c0) * 2;
                        const float* bgm = a.bias + t.n_tile * BN + j * 32;
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const float gamma = __uint_as_float(rg[i]) + __ldg(bgm + i);
                            const float beta = __uint_as_float(rb[i]) + __ldg(bgm + CB + i);
                            const float nrm = (xv[i] - __ldg(mr + 2 * i)) * __ldg(mr + 2 * i + 1);
                            o[i] = nrm * (1.f + gamma) + beta;                // attlwb_spade_resunet.py:92
                        }
                        store_planes32(a, opix * a.out_pitch + a.out_coff + c0, o);
                    }
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < BN / 32; j++) {
                    uint32_t r[32];
                    tmem_ld32(taddr + j * 32, r);
                    tmem_ld_wait();
                    if (valid) {
                        const int c0 = t.n_tile * BN + j * 32;
                        float o[32];
#pragma unroll
                        for (int i = 0; i < 32; i++) o[i] = __uint_as_float(r[i]);
                        if (a.bias) {
#pragma unroll
                            for (int i = 0; i < 32; i++) o[i] += __ldg(a.bias + c0 + i);
                        }
                        if (a.x) {   // residual (ResidualBlock: x + main(x), attlwb_spade_resunet.py:25)
                            float xv[32];
                            load_planes32(a.x, a.x_planes, a.x_plane_stride, opix * a.x_pitch + a.x_coff + c0, xv);
#pragma unroll
                            for (int i = 0; i < 32; i++) o[i] = xv[i] + o[i];
                        }
                        if (a.relu) {
#pragma unroll
                            for (int i = 0; i < 32; i++) o[i] = fmaxf(o[i], 0.f);
                        }
                        if (a.epi == IPER_EPI_F32) {
                            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) +
                                                                    opix * a.out_pitch + a.out_coff + c0);
#pragma unroll
                            for (int i = 0; i < 8; i++) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                        } else {
                            store_planes32(a, opix * a.out_pitch + a.out_coff + c0, o);
                        }
                    }
                }
            }
            // accumulator drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &qres) ==
                cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

static int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box) {
    PFN_cuTensorMapEncodeTiled_v12000 fn = get_encode_fn();
    IPER_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IPER_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    return 0;
}

static int floor_pow2(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }

template <int BN, int NS>
static int launch_gemm(const GemmArgs& g, int max_ctas, cudaStream_t stream) {
    using C = Cfg<BN, NS>;
    static bool attr_set = false;
    if (!attr_set) {
        IPER_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::SMEM_BYTES));
        attr_set = true;
    }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        IPER_CHECK_CUDA(cudaGetDevice(&dev));
        IPER_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int grid = g.total_tiles < num_sms ? g.total_tiles : num_sms;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    conv_gemm_kernel<BN, NS><<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(g);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace iper

using namespace iper;

extern "C" int iper_conv_gemm(const iper_conv_gemm_desc* d, iper_stream_t stream) {
    IPER_REQUIRE(d != nullptr, "iper_conv_gemm: null descriptor");
    IPER_REQUIRE(d->a && d->w, "iper_conv_gemm: null operand");
    IPER_REQUIRE(d->a_planes == d->w_planes && (d->a_planes == 1 || d->a_planes == 2),
                 "iper_conv_gemm: a_planes (%d) and w_planes (%d) must both be 1 or both be 2", d->a_planes, d->w_planes);
    IPER_REQUIRE(d->Cin > 0 && d->Cin % BLOCK_K == 0, "iper_conv_gemm: Cin=%d must be a multiple of 64", d->Cin);
    IPER_REQUIRE(d->a_pitch % 8 == 0 && d->a_coff % 8 == 0 && d->a_coff + d->Cin <= d->a_pitch,
                 "iper_conv_gemm: bad channel window (pitch %d, offset %d, Cin %d)", d->a_pitch, d->a_coff, d->Cin);
    IPER_REQUIRE(((uintptr_t)d->a & 15) == 0 && ((uintptr_t)d->w & 15) == 0, "iper_conv_gemm: operands must be 16-byte aligned");
    IPER_REQUIRE(d->block_n == 32 || d->block_n == 64 || d->block_n == 128 || d->block_n == 256,
                 "iper_conv_gemm: block_n=%d not in {32,64,128,256}", d->block_n);
    IPER_REQUIRE(d->rows > 0 && d->rows % d->block_n == 0, "iper_conv_gemm: rows=%d not a multiple of block_n=%d", d->rows,
                 d->block_n);
    IPER_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0, "iper_conv_gemm: empty input");

    GemmArgs g = {};
    g.mode = d->mode; g.ksize = d->ksize;
    int taps;
    if (d->mode == IPER_CONV_S1) {
        IPER_REQUIRE(d->ksize == 1 || d->ksize == 3 || d->ksize == 5, "iper_conv_gemm: stride-1 ksize must be 1, 3 or 5");
        taps = d->ksize * d->ksize; g.Ho = d->H; g.Wo = d->W; g.oH = d->H; g.oW = d->W; g.phases = 1;
    } else if (d->mode == IPER_CONV_S2) {
        IPER_REQUIRE(d->H % 2 == 0 && d->W % 2 == 0, "iper_conv_gemm: stride-2 needs even H, W");
        taps = 9; g.ksize = 3; g.Ho = d->H / 2; g.Wo = d->W / 2; g.oH = g.Ho; g.oW = g.Wo; g.phases = 1;
    } else if (d->mode == IPER_CONVT_4S2) {
        taps = 4; g.Ho = d->H; g.Wo = d->W; g.oH = 2 * d->H; g.oW = 2 * d->W; g.phases = 4;
    } else if (d->mode == IPER_CONV_ROW5) {
        taps = 5; g.ksize = 5; g.Ho = d->H; g.Wo = d->W; g.oH = d->H; g.oW = d->W; g.phases = 1;
    } else {
        IPER_REQUIRE(false, "iper_conv_gemm: unknown mode %d", d->mode);
    }
    g.N = d->N;
    g.tw = floor_pow2(g.Wo < 16 ? g.Wo : 16);
    g.th = floor_pow2(g.Ho < BLOCK_M / g.tw ? g.Ho : BLOCK_M / g.tw);
    g.tn = BLOCK_M / (g.tw * g.th);
    g.tiles_x = (g.Wo + g.tw - 1) / g.tw;
    if (d->mode == IPER_CONV_ROW5) {       // 128-pixel row segments, 124 outputs each (2-pixel halo either side)
        g.tw = BLOCK_M; g.th = 1; g.tn = 1;
        g.tiles_x = (g.Wo + (BLOCK_M - 4) - 1) / (BLOCK_M - 4);
    }
    g.tiles_y = (g.Ho + g.th - 1) / g.th;
    g.tiles_nb = (g.N + g.tn - 1) / g.tn;
    g.m_tiles = g.tiles_x * g.tiles_y * g.tiles_nb;
    g.n_tiles = d->rows / d->block_n;
    g.total_tiles = g.m_tiles * g.n_tiles * g.phases;
    g.cin_chunks = d->Cin / BLOCK_K;
    g.num_k = taps * g.cin_chunks;
    g.a_coff = d->a_coff; g.a_pitch = d->a_pitch;
    g.rows = d->rows;
    g.epi = d->epi; g.relu = d->relu; g.bias = d->bias;
    g.out = d->out; g.out_planes = d->out_planes; g.out_plane_stride = d->out_plane_stride;
    g.out_pitch = d->out_pitch; g.out_coff = d->out_coff;
    g.x = reinterpret_cast<const __half*>(d->x); g.x_planes = d->x_planes; g.x_plane_stride = d->x_plane_stride;
    g.x_pitch = d->x_pitch; g.x_coff = d->x_coff;
    g.mean_rstd = d->mean_rstd; g.spade_C = d->spade_C;
    g.bg = d->bg; g.bg_batch_stride = d->bg_batch_stride; g.img = d->img; g.mask = d->mask; g.pred = d->pred;

    // ---- epilogue-specific validation ----
    if (d->epi == IPER_EPI_HEADS) {
        IPER_REQUIRE(d->mode == IPER_CONV_ROW5 && d->block_n == 32 && d->rows == 32,
                     "iper_conv_gemm: heads epilogue needs mode IPER_CONV_ROW5 and rows = block_n = 32");
        IPER_REQUIRE(d->mask || d->img || d->pred, "iper_conv_gemm: heads epilogue without outputs");
        IPER_REQUIRE(!d->pred || d->bg, "iper_conv_gemm: pred needs bg");
    } else {
        IPER_REQUIRE(d->block_n >= 64 && d->mode != IPER_CONV_ROW5, "iper_conv_gemm: block_n=32 / ROW5 are reserved for the heads epilogue");
        IPER_REQUIRE(d->out != nullptr, "iper_conv_gemm: null output");
        IPER_REQUIRE(d->out_pitch % 8 == 0 && d->out_coff % 8 == 0, "iper_conv_gemm: output channel window must be 8-aligned");
        if (d->epi == IPER_EPI_SPADE) {
            IPER_REQUIRE(d->x && d->mean_rstd && d->bias && d->spade_C * 2 == d->rows,
                         "iper_conv_gemm: SPADE epilogue needs x, mean_rstd, bias and rows == 2*C");
        }
        if (d->x) IPER_REQUIRE(d->x_pitch % 8 == 0 && d->x_coff % 8 == 0, "iper_conv_gemm: x channel window must be 8-aligned");
    }

    // ---- tensor maps ----
    const size_t esz = 2;
    for (int p = 0; p < d->a_planes; p++) {
        const __half* base = reinterpret_cast<const __half*>(d->a) + (size_t)p * d->a_plane_stride;
        if (d->mode == IPER_CONV_S2) {
            cuuint64_t dims[5] = {(cuuint64_t)2 * d->a_pitch, (cuuint64_t)d->W / 2, 2, (cuuint64_t)d->H / 2, (cuuint64_t)d->N};
            cuuint64_t str[4] = {(cuuint64_t)2 * d->a_pitch * esz, (cuuint64_t)d->W * d->a_pitch * esz,
                                 (cuuint64_t)2 * d->W * d->a_pitch * esz, (cuuint64_t)d->H * d->W * d->a_pitch * esz};
            cuuint32_t box[5] = {BLOCK_K, (cuuint32_t)g.tw, 1, (cuuint32_t)g.th, (cuuint32_t)g.tn};
            if (int rc = encode_map(&g.mapA[p], base, 5, dims, str, box)) return rc;
        } else {
            cuuint64_t dims[4] = {(cuuint64_t)d->a_pitch, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
            cuuint64_t str[3] = {(cuuint64_t)d->a_pitch * esz, (cuuint64_t)d->W * d->a_pitch * esz,
                                 (cuuint64_t)d->H * d->W * d->a_pitch * esz};
            cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)g.tw, (cuuint32_t)g.th, (cuuint32_t)g.tn};
            if (int rc = encode_map(&g.mapA[p], base, 4, dims, str, box)) return rc;
        }
        const __half* wb = reinterpret_cast<const __half*>(d->w) + (size_t)p * d->w_plane_stride;
        const cuuint64_t ktot = (cuuint64_t)taps * d->Cin;
        cuuint64_t wdims[2] = {ktot, (cuuint64_t)d->rows * g.phases};
        cuuint64_t wstr[1] = {ktot * esz};
        cuuint32_t wbox[2] = {BLOCK_K, (cuuint32_t)d->block_n};
        if (int rc = encode_map(&g.mapB[p], wb, 2, wdims, wstr, wbox)) return rc;
    }
    if (d->a_planes == 1) { g.mapA[1] = g.mapA[0]; g.mapB[1] = g.mapB[0]; }

    cudaStream_t s = (cudaStream_t)stream;
    const int ns = d->a_planes;
#define IPER_DISPATCH(BNV)                                                     \
    return ns == 2 ? launch_gemm<BNV, 2>(g, d->max_ctas, s) : launch_gemm<BNV, 1>(g, d->max_ctas, s)
    switch (d->block_n) {
        case 32: IPER_DISPATCH(32);
        case 64: IPER_DISPATCH(64);
        case 128: IPER_DISPATCH(128);
        default: IPER_DISPATCH(256);
    }
#undef IPER_DISPATCH
}
