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
// Three kernels share the tile decode, the tap arithmetic and the epilogue:
//   conv_halo_pair_kernel   (default wherever a layer qualifies) CTA pairs + vertical halo + tap program, 8 epilogue warps
//   conv_gemm_pair_kernel   CTA pairs (cta_group::2), one TMA box per tap: stride-2 convs, wide transposed convs
//   conv_gemm_kernel        one CTA per tile: small maps, N = 32/64 tiles, the non-halo layers of the fp16+e4m3 mode
// conv_gemm_kernel — one persistent CTA per SM, warp-specialised:
//   warp 0   : TMA producer (one lane)            smem ring of STAGES x {A planes, B planes}
//   warp 1   : TMEM allocator + MMA issuer (one lane issues tcgen05.mma, tcgen05.commit frees ring slots)
//   warps 2-5: epilogue — tcgen05.ld the fp32 accumulator (thread = pixel row), fused bias / ReLU / residual /
//              SPADE(instance-norm) / heads+composite, stores; double-buffered accumulators overlap tile i's
//              epilogue with tile i+1's MMAs.
// Operand modes (template NS = planes format of A and of the packed weights):
//   1  fp16                : one MMA per K sub-step
//   2  split fp16 (hi+lo)  : lo*hi + hi*lo + hi*hi into one fp32 accumulator (~22-bit operands, 3 MMAs)
//   3  fp16 + fp8 cross    : hi*hi in kind::f16 into accumulator D1 (pass 0 over K); the two small cross terms lo*hi and hi*lo need
//                            only ~4 bits, so they run in kind::f8f6f4 (e4m3, 2x rate) into a second accumulator D2
//                            with power-of-two pre-scaling (pass 1 over K — the MMA kinds are never interleaved);
//                            the epilogue forms D1 + D2 * cross_scale (2 MMA-equivalents).
#include <cstdlib>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "iper_b200.h"

namespace iper {

constexpr int GEMM_THREADS = 192;
constexpr int BLOCK_M = 128;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 196 * 1024;   // ring buffer budget; keeps one CTA per SM (TMEM is per-CTA 512 cols)

// Tap program of the halo kernel (conv_halo_pair_kernel).  An A LOAD fetches the tile's pixels plus `halo` extra image
// rows for one horizontal tap offset; every ENTRY is one tap's MMA group: it reads the 128-pixel view of that box that
// starts `a_row_off` smem rows down (a vertical tap = a 1024-byte-aligned offset into the same box), multiplies with the
// weight tile at (K column b_k + chunk*64, row b_row + n_tile*BN) and accumulates into accumulator block `acc`.
constexpr int HALO_MAX_LOADS = 3, HALO_MAX_ENTRIES = 16;
// nblk > 1 (transposed conv, fused N): ONE MMA group of N = nblk * BN feeds the accumulator blocks [acc, acc + nblk) —
// several output phases read the same view — and CTA r of the pair stages nblk/2 whole BN-row weight boxes
// {fb_row[r][i], fb_k[r][i]} (its half of the concatenated N) instead of half a box.
struct HaloEntry { int a_row_off, b_row, b_k, acc, nblk; int fb_row[2][2], fb_k[2][2]; };
struct HaloSched {
    int n_loads, acc_blocks, box_rows;          // box_rows = (th + halo) * tw
    int na, nb, a_plane_bytes, tmem_cols;       // ring depths, bytes of one A box of a slot, TMEM columns to allocate
    int k8;                                     // format 3: e4m3 channels per step (128: 128-byte rows, 64: 64-byte rows)
    int cat;                                    // heads, split fp16: weights concatenated along N ([w_hi ; w_lo]), see kernel
    int blk_phase[4];                           // transposed conv: output phase stored in accumulator block i
    int ox[HALO_MAX_LOADS], oy[HALO_MAX_LOADS], first[HALO_MAX_LOADS], count[HALO_MAX_LOADS];
    HaloEntry e[HALO_MAX_ENTRIES];
};

struct alignas(64) GemmArgs {
    CUtensorMap mapA[3];
    CUtensorMap mapB[3];
    CUtensorMap mapBf[2];    // halo kernel, fused-N entries: whole BN-row weight boxes of fp16 planes 0 / 1
    int mode, ksize;
    int N, Ho, Wo;           // grid the M tiles walk (conv: output grid; convT: input grid)
    int oH, oW;              // stored output spatial dims
    int tw, th, tn;          // patch = tw x th x tn pixels = 128
    int tiles_x, tiles_y, tiles_nb, m_tiles, m_groups, n_tiles, phases, total_tiles;
    int cin_chunks, num_k;
    int a_coff, a_pitch;
    int rows;
    int epi, relu;
    const float* bias;
    void* out; int out_planes; long long out_plane_stride; int out_pitch, out_coff;
    const __half* x; int x_planes; long long x_plane_stride; int x_pitch, x_coff;
    const float* mean_rstd; int spade_C;
    const float* bg; long long bg_batch_stride; float* img; float* mask; float* pred;
    float cross_scale;
    double* stats_ws;        // optional (N, Cout, 2) fp64 sums of the stored output (instance-norm statistics)
    const float* w_scale_inv;   // optional device scalar 1/s: the packed weights hold w * s (s a power of two), see iper_b200.h
    HaloSched hs;            // conv_halo_pair_kernel only
};

// BN = GEMM-N tile (output channels), NS = operand format (1, 2, 3), TM = 128-pixel M tiles per CTA that share one
// weight tile.  The kernel is bound by operand delivery (L2 -> smem, ~42 B/clk/SM chip-wide), so bytes per MMA are
// what matters: TM = 2 halves the weight traffic per output.  TM = 2 stages hold 32 K-elements (64-byte rows,
// SWIZZLE_64B) so that three stages still fit; TM = 1 stages hold 64 (128-byte rows, SWIZZLE_128B).
template <int BN, int NS, int TM>
struct Cfg {
    static_assert(NS != 3 || TM == 1, "the fp16+fp8 mode keeps one M tile per CTA (TMEM holds D1 and D2)");
    static constexpr int BK = (TM == 2) ? 32 : 64;
    static constexpr int ROW16 = BK * 2;                              // bytes per fp16 operand row
    static constexpr int A_TILE = BLOCK_M * ROW16;                    // one fp16 A tile
    static constexpr int B_TILE = BN * ROW16;
    static constexpr int PL = (NS == 2) ? 2 : 1;                      // fp16 planes resident per stage
    // NS=3 walks K twice per tile (pass 0: fp16 tiles -> D1, pass 1: the four e4m3 tiles -> D2) so that the two MMA
    // kinds are never interleaved; either pass fills a stage with A_TILE + B_TILE bytes
    static constexpr int STAGE_BYTES = PL * (TM * A_TILE + B_TILE);
    static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) < MAX_STAGES ? (SMEM_BUDGET / STAGE_BYTES) : MAX_STAGES;
    static constexpr int ACC_COLS = (NS == 3 ? 2 : 1) * TM * BN;      // TM accumulators (NS=3: [D1 | D2])
    static constexpr int NACC = (2 * ACC_COLS <= 512) ? 2 : 1;        // double-buffer the accumulators when TMEM allows
    static constexpr int TMEM_COLS = (NACC * ACC_COLS) < 32 ? 32 : (NACC * ACC_COLS);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;    // + alignment slack
    static constexpr int B_BASE = PL * TM * A_TILE;
    static_assert(STAGES >= 2, "need at least a double-buffered ring");
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns must be a power of two <= 512");
    static_assert(SMEM_BYTES >= 116 * 1024, "ring must be large enough that only one CTA fits per SM (TMEM allocation)");
};

struct TileCoord {
    int phase, n_tile, pn0, py0, px0;
};
// tile -> (phase, M group, N tile); M tile `t` of the group.  A group index past the last M tile yields pn0 = N:
// its loads are fully out of bounds (zero fill) and its epilogue rows are all invalid.
template <int TM>
IPER_DEVINL TileCoord decode_tile(const GemmArgs& a, int tile, int t) {
    TileCoord c;
    c.n_tile = tile % a.n_tiles;
    int r = tile / a.n_tiles;
    const int m = (r % a.m_groups) * TM + t;
    c.phase = r / a.m_groups;
    if (m >= a.m_tiles) { c.px0 = 0; c.py0 = 0; c.pn0 = a.N; return c; }
    c.px0 = (a.mode == IPER_CONV_ROW5) ? (m % a.tiles_x) * (BLOCK_M - 4) - 2 : (m % a.tiles_x) * a.tw;
    const int r2 = m / a.tiles_x;
    c.py0 = (r2 % a.tiles_y) * a.th;
    c.pn0 = (r2 / a.tiles_y) * a.tn;
    return c;
}

// store / load 32 consecutive channels of one pixel in the planes format `fmt` (common.cuh)
IPER_DEVINL void store_planes32(const GemmArgs& a, size_t elem_off, const float (&v)[32]) {
    __half* base = reinterpret_cast<__half*>(a.out);
#pragma unroll
    for (int g = 0; g < 4; g++) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; j++) t[j] = v[8 * g + j];
        store_planes8(base, a.out_planes, a.out_plane_stride, elem_off + 8 * g, t);
    }
}
IPER_DEVINL void load_planes32(const __half* x, int fmt, long long plane_stride, size_t elem_off, float (&v)[32]) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
        float t[8];
        load_planes8(x, fmt, plane_stride, elem_off + 8 * g, t);
#pragma unroll
        for (int j = 0; j < 8; j++) v[8 * g + j] = t[j];
    }
}

// Epilogue of one 128-pixel tile: thread = accumulator row (pixel); `taddr` = TMEM address of the tile's D1 columns
// for this warp's lane quarter.  Shared by the single-CTA and the CTA-pair kernels.
// `nsplit` warps share a TMEM lane quarter (1: the 4-warp epilogue of conv_gemm[_pair]_kernel, 2: the 8-warp epilogue of
// the halo kernel); warp `hsel` of them takes the 32-column chunks j with j % nsplit == hsel.
template <int BN, int NS>
IPER_DEVINL void epilogue_tile(const GemmArgs& a, const TileCoord& t, uint32_t taddr, int row, int lane, int tx, int ty,
                               int tni, int hsel = 0, int nsplit = 1, double* pend = nullptr) {
    // weights are packed as w * s with s a power of two that lifts small weights (and their lo plane) out of fp16's
    // subnormal range; the accumulator is scaled back here — exact, so in-range layers are bit-identical to s = 1
    const float wsc = a.w_scale_inv ? __ldg(a.w_scale_inv) : 1.f;
    // accumulator chunk: 32 fp32 columns of D1 (+ the matching columns of the fp8 cross-term accumulator D2)
    auto ld_acc = [&](uint32_t taddr_col, uint32_t (&r)[32]) {
        tmem_ld32(taddr_col, r);
        if constexpr (NS == 3) {
            uint32_t r2[32];
            tmem_ld32(taddr_col + BN, r2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++)
                r[i] = __float_as_uint(fmaf(__uint_as_float(r2[i]), a.cross_scale, __uint_as_float(r[i])));
        } else {
            tmem_ld_wait();
            if (wsc != 1.f) {
#pragma unroll
                for (int i = 0; i < 32; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) * wsc);
            }
        }
    };
            const int n = t.pn0 + tni, y = t.py0 + ty, xx = t.px0 + tx;
            const bool valid = (n < a.N) && (y < a.Ho) && (xx < a.Wo);
            int oy = y, ox = xx;
            if (a.mode == IPER_CONVT_4S2) { oy = 2 * y + (t.phase >> 1); ox = 2 * xx + (t.phase & 1); }
            const size_t opix = ((size_t)n * a.oH + oy) * a.oW + ox;

            if (a.epi == IPER_EPI_HEADS) {
                if constexpr (BN == 32) {
                    __shared__ float s_ex[BLOCK_M * 21];          // D[row][dx*4+o], 20 used columns (+1 pad)
                    const int nthr = 128 * nsplit;
                    // horizontal taps: output x of a tw-pixel row segment needs D of x-2 .. x+2 (same image row = same
                    // segment of s_ex rows), so the two outermost pixels either side are recomputed by the neighbour tile.
                    // With two warps per pixel quarter, warp 0 writes img/pred channels 0-1, warp 1 channel 2 and the mask.
                    const int xo = t.px0 + tx;
                    const bool active = tx >= 2 && tx < a.tw - 2 && xo >= 0 && xo < a.Wo && n < a.N && y < a.Ho;
                    const size_t hw = (size_t)a.oH * a.oW, p = (size_t)y * a.oW + xo;
                    const int c_lo = (nsplit == 1) ? 0 : (hsel == 0 ? 0 : 2), c_hi = (nsplit == 1) ? 3 : (hsel == 0 ? 2 : 3);
                    float bgv[3] = {0.f, 0.f, 0.f};     // background pixels for the composite: in flight during the exchange
                    if (active && a.pred) {
#pragma unroll
                        for (int c = 0; c < 3; c++)
                            if (c >= c_lo && c < c_hi) bgv[c] = __ldg(a.bg + (size_t)n * a.bg_batch_stride + c * hw + p);
                    }
                    if (hsel == 0) {
                        uint32_t r[32];
                        ld_acc(taddr, r);
                        if (a.hs.cat) {          // N-concatenated split weights: columns [32,64) hold the hi*lo term
                            uint32_t r2[32];
                            tmem_ld32(taddr + 32, r2);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 20; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]) * wsc);
                        }
#pragma unroll
                        for (int i = 0; i < 20; i++) s_ex[row * 21 + i] = __uint_as_float(r[i]);
                    }
                    asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
                    if (active) {
                        auto head = [&](int o) {
                            float acc4 = 0.f;
#pragma unroll
                            for (int dx = 0; dx < 5; dx++) acc4 += s_ex[(row + dx - 2) * 21 + dx * 4 + o];
                            return acc4;
                        };
                        const float m = 1.f / (1.f + expf(-head(3)));
                        if (a.mask && hsel == nsplit - 1) a.mask[(size_t)n * hw + p] = m;
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            if (c < c_lo || c >= c_hi) continue;
                            const float v = tanhf(head(c));
                            if (a.img) a.img[((size_t)n * 3 + c) * hw + p] = v;
                            if (a.pred) a.pred[((size_t)n * 3 + c) * hw + p] = m * bgv[c] + (1.f - m) * v;   // imitator.py:393
                        }
                    }
                    asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");   // s_ex is reused by the next tile
                }
            } else if (a.epi == IPER_EPI_SPADE) {
                constexpr int CB = BN / 2;      // channels per tile: columns [0,CB) gamma, [CB,2CB) beta
#pragma unroll 1
                for (int j = hsel; j < CB / 32; j += nsplit) {
                    uint32_t rg[32], rb[32];
                    ld_acc(taddr + j * 32, rg);
                    ld_acc(taddr + CB + j * 32, rb);
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
                const bool warp_uniform_n = (a.tw * a.th) % 32 == 0;      // all 32 rows of a warp lie in one image
#pragma unroll 1
                for (int j = hsel; j < BN / 32; j += nsplit) {
                    uint32_t r[32];
                    ld_acc(taddr + j * 32, r);
                    const int c0 = t.n_tile * BN + j * 32;
                    float o[32];
#pragma unroll
                    for (int i = 0; i < 32; i++) o[i] = __uint_as_float(r[i]);
                    if (valid) {
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
                    if (a.stats_ws != nullptr) {
                        // fused instance-norm statistics of the value just stored (invalid rows contribute 0)
                        if (warp_uniform_n) {
                            // transpose-reduce over the warp's 32 pixels, then one fp64 atomic per channel (lane = channel)
                            float sq[32];
#pragma unroll
                            for (int i = 0; i < 32; i++) { o[i] = valid ? o[i] : 0.f; sq[i] = o[i] * o[i]; }
                            const float s1 = warp_transpose_sum32(o, lane), s2 = warp_transpose_sum32(sq, lane);
                            const int nw = __shfl_sync(0xffffffffu, n, 0);
                            if (pend) {         // this warp's running sums of the current image (flushed by the caller)
                                pend[(c0 + lane) * 2] += (double)s1;
                                pend[(c0 + lane) * 2 + 1] += (double)s2;
                            } else if (nw < a.N) {
                                atomicAdd(a.stats_ws + ((size_t)nw * a.rows + c0 + lane) * 2, (double)s1);
                                atomicAdd(a.stats_ws + ((size_t)nw * a.rows + c0 + lane) * 2 + 1, (double)s2);
                            }
                        } else if (valid) {     // tiny maps (a warp spans several images): plain per-value atomics
                            for (int i = 0; i < 32; i++) {
                                atomicAdd(a.stats_ws + ((size_t)n * a.rows + c0 + i) * 2, (double)o[i]);
                                atomicAdd(a.stats_ws + ((size_t)n * a.rows + c0 + i) * 2 + 1, (double)o[i] * (double)o[i]);
                            }
                        }
                    }
                }
            }
}

template <int BN, int NS, int TM>
__global__ void __launch_bounds__(GEMM_THREADS, 1) conv_gemm_kernel(const __grid_constant__ GemmArgs a) {
    using C = Cfg<BN, NS, TM>;
    constexpr int BK = C::BK;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    // per-epilogue-warp running instance-norm sums of the image being processed (single-N-tile layers): a CTA walks a
    // CONTIGUOUS range of tiles, so almost all of them belong to one image and the fp64 atomics on the (N, C, 2) workspace are
    // issued once per image per warp instead of once per tile (the K = 64 stem GEMM has 512 tiles per image on 128 addresses)
    __shared__ double s_pend[4][BN][2];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // contiguous tile range of this CTA
    const int tile_lo = (int)((long long)a.total_tiles * blockIdx.x / gridDim.x);
    const int tile_hi = (int)((long long)a.total_tiles * (blockIdx.x + 1) / gridDim.x);
    // 1024-byte aligned ring buffer (swizzle atoms are 1024 B / 512 B)
    const uint32_t ring = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* ring_ptr = smem_dyn + (ring - smem_u32(smem_dyn));
    // operand tiles of a stage.  p = 0: fp16 hi; NS=2: p = 1 fp16 lo; NS=3 (pass 1, TM = 1): p = 1 / 2 = e4m3 a8 / l8
    // (w8 / wl8), half-size tiles reusing the space of the pass-0 fp16 tile.
    auto sA = [&](int stage, int t, int p) -> uint8_t* {
        const int off = (NS == 3) ? (p == 2 ? C::A_TILE / 2 : 0) : (p * TM + t) * C::A_TILE;
        return ring_ptr + stage * C::STAGE_BYTES + off;
    };
    auto sB = [&](int stage, int p) -> uint8_t* {
        const int off = (NS == 3) ? (p == 2 ? C::B_TILE / 2 : 0) : p * C::B_TILE;
        return ring_ptr + stage * C::STAGE_BYTES + C::B_BASE + off;
    };
    auto desc16 = [](uint32_t saddr) -> uint64_t { return BK == 64 ? umma_desc_sw128(saddr) : umma_desc_sw64(saddr); };

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < C::NACC; i++) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4); }
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
    constexpr int NPASS = (NS == 3) ? 2 : 1;

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            for (int tile = tile_lo; tile < tile_hi; tile++) {
                TileCoord tc[TM];
#pragma unroll
                for (int t = 0; t < TM; t++) tc[t] = decode_tile<TM>(a, tile, t);
                const int brow = tc[0].phase * a.rows + tc[0].n_tile * BN;
                for (int pass = 0; pass < NPASS; pass++)
                for (int kb = 0; kb < a.num_k; kb++) {
                    // operand planes loaded into this stage: NS=1: [0,1)  NS=2: [0,2)  NS=3: pass 0 -> [0,1), pass 1 -> [1,3)
                    const int p_lo = (NS == 3 && pass == 1) ? 1 : 0;
                    const int p_hi = (NS == 3) ? (pass == 0 ? 1 : 3) : NS;
                    const int tap = kb / a.cin_chunks, cc = kb - tap * a.cin_chunks;
                    mbar_wait(&empty_bar[stage], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
#pragma unroll
                    for (int t = 0; t < TM; t++) {
                        if (a.mode == IPER_CONV_S2) {
                            // input pixel = 2*o - 1 + d : d=0 -> (parity 1, o-1), d=1 -> (parity 0, o), d=2 -> (parity 1, o)
                            const int dy = tap / 3, dx = tap - 3 * dy;
                            const int py = (dy != 1), sy = (dy == 0) ? -1 : 0;
                            const int px = (dx != 1), sx = (dx == 0) ? -1 : 0;
                            const int c0 = px * a.a_pitch + a.a_coff + cc * BK;
                            for (int p = p_lo; p < p_hi; p++)
                                tma_load_5d(sA(stage, t, p), &a.mapA[p], &full_bar[stage], c0, tc[t].px0 + sx, py,
                                            tc[t].py0 + sy, tc[t].pn0);
                        } else {
                            int oy, ox;
                            if (a.mode == IPER_CONV_ROW5) {
                                oy = tap - 2; ox = 0;
                            } else if (a.mode == IPER_CONV_S1) {
                                const int dy = tap / a.ksize, dx = tap - dy * a.ksize;
                                oy = dy - a.ksize / 2; ox = dx - a.ksize / 2;
                            } else {  // transposed 4x4 s2 p1, phase (py,px), tap (ta,tb): see pack_convT_weight (ops.py)
                                const int py = tc[t].phase >> 1, px = tc[t].phase & 1;
                                const int ta = tap >> 1, tb = tap & 1;
                                oy = py == 0 ? (ta == 0 ? 0 : -1) : (ta == 0 ? 1 : 0);
                                ox = px == 0 ? (tb == 0 ? 0 : -1) : (tb == 0 ? 1 : 0);
                            }
                            const int c0 = a.a_coff + cc * BK;
                            for (int p = p_lo; p < p_hi; p++)
                                tma_load_4d(sA(stage, t, p), &a.mapA[p], &full_bar[stage], c0, tc[t].px0 + ox,
                                            tc[t].py0 + oy, tc[t].pn0);
                        }
                    }
                    for (int p = p_lo; p < p_hi; p++)
                        tma_load_2d(sB(stage, p), &a.mapB[p], &full_bar[stage], kb * BK, brow);
                    if (++stage == C::STAGES) { stage = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        constexpr uint32_t idesc = umma_idesc_f16(BLOCK_M, BN);
        constexpr uint32_t idesc8 = umma_idesc_e4m3(BLOCK_M, BN);
        int stage = 0; uint32_t ph = 0; int it = 0;
        for (int tile = tile_lo; tile < tile_hi; tile++, it++) {
            const int acc = it % C::NACC; const uint32_t acc_ph = (it / C::NACC) & 1;
            mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
            for (int pass = 0; pass < NPASS; pass++)
            for (int kb = 0; kb < a.num_k; kb++) {
                mbar_wait(&full_bar[stage], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t fresh = (kb == 0) ? 0u : 1u;
                    if constexpr (NS == 3) {
                        if (pass == 0) {            // D1 += hi * w_hi (kind::f16)
                            const uint32_t a16 = smem_u32(sA(stage, 0, 0)), b16 = smem_u32(sB(stage, 0));
#pragma unroll
                            for (int k = 0; k < BK / 16; k++)
                                umma_f16(d_tmem, desc16(a16 + k * 32), desc16(b16 + k * 32), idesc, (k == 0) ? fresh : 1u);
                        } else {                    // D2 += l8 * w8 + a8 * wl8 (kind::f8f6f4, K = 32 per instruction)
                            const uint32_t a8 = smem_u32(sA(stage, 0, 1)), l8 = smem_u32(sA(stage, 0, 2));
                            const uint32_t w8 = smem_u32(sB(stage, 1)), wl8 = smem_u32(sB(stage, 2));
#pragma unroll
                            for (int k = 0; k < BK / 32; k++) {
                                umma_f8(d_tmem + BN, umma_desc_sw64(l8 + k * 32), umma_desc_sw64(w8 + k * 32), idesc8,
                                        (k == 0) ? fresh : 1u);
                                umma_f8(d_tmem + BN, umma_desc_sw64(a8 + k * 32), umma_desc_sw64(wl8 + k * 32), idesc8, 1u);
                            }
                        }
                    } else {
                        // (A plane, B plane): small cross terms first, hi*hi last
                        constexpr int NPAIR = (NS == 2) ? 3 : 1;
                        const int pa[3] = {NS == 2 ? 1 : 0, 0, 0};
                        const int pb[3] = {0, NS == 2 ? 1 : 0, 0};
#pragma unroll
                        for (int t = 0; t < TM; t++) {
                            uint32_t first = fresh;
#pragma unroll
                            for (int q = 0; q < NPAIR; q++) {
                                const uint32_t abase = smem_u32(sA(stage, t, pa[q])), bbase = smem_u32(sB(stage, pb[q]));
#pragma unroll
                                for (int k = 0; k < BK / 16; k++) {
                                    umma_f16(d_tmem + t * BN, desc16(abase + k * 32), desc16(bbase + k * 32), idesc, first);
                                    first = 1u;
                                }
                            }
                        }
                    }
                    umma_commit(&empty_bar[stage]);                        // ring slot reusable once MMAs retire
                    if (pass == NPASS - 1 && kb == a.num_k - 1) umma_commit(&tmem_full_bar[acc]);  // accumulators ready
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
        // pending statistics: only when the layer has one N tile (channel = accumulator column) and a warp's 32 rows lie in
        // one image, so that the warp's sums belong to exactly one (image, channel) set at a time
        double* pend = (a.stats_ws != nullptr && a.n_tiles == 1 && (a.tw * a.th) % 32 == 0) ? &s_pend[q][0][0] : nullptr;
        int pend_n = -1;
        auto flush = [&]() {
            if (pend_n >= 0 && pend_n < a.N)
                for (int c = lane; c < BN; c += 32) {
                    atomicAdd(a.stats_ws + ((size_t)pend_n * a.rows + c) * 2, pend[c * 2]);
                    atomicAdd(a.stats_ws + ((size_t)pend_n * a.rows + c) * 2 + 1, pend[c * 2 + 1]);
                }
            for (int c = lane; c < BN; c += 32) { pend[c * 2] = 0.0; pend[c * 2 + 1] = 0.0; }
            __syncwarp();
        };
        if (pend) flush();
        for (int tile = tile_lo; tile < tile_hi; tile++, it++) {
            const int acc = it % C::NACC; const uint32_t acc_ph = (it / C::NACC) & 1;
            mbar_wait(&tmem_full_bar[acc], acc_ph);
            tc_fence_after();
#pragma unroll 1
            for (int tm = 0; tm < TM; tm++) {
            const TileCoord t = decode_tile<TM>(a, tile, tm);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::ACC_COLS + tm * BN;
            if (pend) {
                const int nw = t.pn0 + (q * 32) / (a.tw * a.th);        // image of this warp's rows
                if (nw != pend_n) { flush(); pend_n = nw; }
            }
            epilogue_tile<BN, NS>(a, t, taddr, row, lane, tx, ty, tni, 0, 1, pend);
            }  // tm
            // accumulators drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (pend) flush();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2).  A 2-CTA cluster computes two 128-pixel M tiles against one N tile: every CTA
// stages its own A tile and HALF of the weight tile (N/2 rows); the leader CTA issues tcgen05.mma.cta_group::2
// (M = 256, N = BN) that reads A and B from both CTAs' shared memory and writes each CTA's own TMEM.  Per CTA a K step
// is A + B/2 bytes (64 KB instead of 96 KB for BN = 256, split fp16), so the ring is 3 deep instead of 2 and every SM
// ingests 1.5x fewer operand bytes per MMA.  Barriers: TMA of both CTAs completes on the LEADER's full barrier; the
// leader's tcgen05.commit multicasts to both CTAs' empty / tmem_full barriers; both epilogues arrive on the leader's
// tmem_empty barrier.  NS in {1, 2}.
// ------------------------------------------------------------------------------------------------------------
template <int BN, int NS>
struct Cfg2 {
    static_assert(NS == 1 || NS == 2, "CTA pairs support fp16 and split fp16");
    static constexpr int BK = 64;
    static constexpr int A_TILE = BLOCK_M * 128;
    static constexpr int B_HALF = (BN / 2) * 128;
    static constexpr int STAGE_BYTES = NS * (A_TILE + B_HALF);          // per CTA
    static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) < MAX_STAGES ? (SMEM_BUDGET / STAGE_BYTES) : MAX_STAGES;
    static constexpr int ACC_COLS = BN;
    static constexpr int NACC = 2;
    static constexpr int TMEM_COLS = (NACC * ACC_COLS) < 32 ? 32 : (NACC * ACC_COLS);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
    static_assert(SMEM_BYTES >= 116 * 1024, "one CTA per SM");
};

template <int BN, int NS>
__global__ void __launch_bounds__(GEMM_THREADS, 1) conv_gemm_pair_kernel(const __grid_constant__ GemmArgs a) {
    using C = Cfg2<BN, NS>;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];     // used in the leader CTA (count 2 + tx bytes of both CTAs)
    __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];    // per CTA, signalled by the leader's multicast commit
    __shared__ __align__(8) uint64_t tmem_full_bar[2];         // per CTA, multicast commit
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];        // used in the leader CTA (8 epilogue warps of the pair)
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                   // 0 = leader
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const uint32_t ring = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* ring_ptr = smem_dyn + (ring - smem_u32(smem_dyn));
    auto sA = [&](int stage, int p) -> uint8_t* { return ring_ptr + stage * C::STAGE_BYTES + p * C::A_TILE; };
    auto sB = [&](int stage, int p) -> uint8_t* { return ring_ptr + stage * C::STAGE_BYTES + NS * C::A_TILE + p * C::B_HALF; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::STAGES; i++) { mbar_init(&full_bar[i], 2); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 8); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        for (int p = 0; p < NS; p++) { tma_prefetch_desc(&a.mapA[p]); tma_prefetch_desc(&a.mapB[p]); }
    }
    if (warp == 1) tmem_alloc_2sm(&tmem_base_slot, C::TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();            // barriers of BOTH CTAs are initialised before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    // work unit = (phase, n tile, pair of consecutive M tiles); this CTA owns M tile 2*pair + rank
    const int m_pairs = (a.m_tiles + 1) >> 1;
    const int units = m_pairs * a.n_tiles * a.phases;
    auto unit_coord = [&](int u) -> TileCoord {
        TileCoord c;
        c.n_tile = u % a.n_tiles;
        const int r = u / a.n_tiles;
        const int m = (r % m_pairs) * 2 + (int)rank;
        c.phase = r / m_pairs;
        if (m >= a.m_tiles) { c.px0 = 0; c.py0 = 0; c.pn0 = a.N; return c; }
        c.px0 = (m % a.tiles_x) * a.tw;
        const int r2 = m / a.tiles_x;
        c.py0 = (r2 % a.tiles_y) * a.th;
        c.pn0 = (r2 / a.tiles_y) * a.tn;
        return c;
    };

    if (warp == 0) {
        // =========================== TMA producer (both CTAs) ===========================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            for (int u = cluster_id; u < units; u += num_clusters) {
                const TileCoord t = unit_coord(u);
                const int brow = t.phase * a.rows + t.n_tile * BN + (int)rank * (BN / 2);     // this CTA's half of the N tile
                for (int kb = 0; kb < a.num_k; kb++) {
                    const int tap = kb / a.cin_chunks, cc = kb - tap * a.cin_chunks;
                    mbar_wait(&empty_bar[stage], ph ^ 1);
                    if (a.mode == IPER_CONV_S2) {
                        const int dy = tap / 3, dx = tap - 3 * dy;
                        const int py = (dy != 1), sy = (dy == 0) ? -1 : 0;
                        const int px = (dx != 1), sx = (dx == 0) ? -1 : 0;
                        const int c0 = px * a.a_pitch + a.a_coff + cc * C::BK;
                        for (int p = 0; p < NS; p++)
                            tma_load_5d_2sm(sA(stage, p), &a.mapA[p], &full_bar[stage], c0, t.px0 + sx, py, t.py0 + sy, t.pn0);
                    } else {
                        int oy, ox;
                        if (a.mode == IPER_CONV_S1) {
                            const int dy = tap / a.ksize, dx = tap - dy * a.ksize;
                            oy = dy - a.ksize / 2; ox = dx - a.ksize / 2;
                        } else {
                            const int py = t.phase >> 1, px = t.phase & 1;
                            const int ta = tap >> 1, tb = tap & 1;
                            oy = py == 0 ? (ta == 0 ? 0 : -1) : (ta == 0 ? 1 : 0);
                            ox = px == 0 ? (tb == 0 ? 0 : -1) : (tb == 0 ? 1 : 0);
                        }
                        const int c0 = a.a_coff + cc * C::BK;
                        for (int p = 0; p < NS; p++)
                            tma_load_4d_2sm(sA(stage, p), &a.mapA[p], &full_bar[stage], c0, t.px0 + ox, t.py0 + oy, t.pn0);
                    }
                    for (int p = 0; p < NS; p++)
                        tma_load_2d_2sm(sB(stage, p), &a.mapB[p], &full_bar[stage], kb * C::BK, brow);
                    // the leader's full barrier collects the bytes of both CTAs and one arrival from each producer
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
                    else mbar_arrive_remote(&full_bar[stage], 0);
                    if (++stage == C::STAGES) { stage = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer (leader CTA only) ===========================
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(2 * BLOCK_M, BN);
            int stage = 0; uint32_t ph = 0; int it = 0;
            for (int u = cluster_id; u < units; u += num_clusters, it++) {
                const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
                for (int kb = 0; kb < a.num_k; kb++) {
                    mbar_wait(&full_bar[stage], ph);
                    tc_fence_after();
                    if (elect_one()) {
                        uint32_t first = (kb == 0) ? 0u : 1u;
                        constexpr int NPAIR = (NS == 2) ? 3 : 1;
                        const int pa[3] = {NS == 2 ? 1 : 0, 0, 0};
                        const int pb[3] = {0, NS == 2 ? 1 : 0, 0};
#pragma unroll
                        for (int q = 0; q < NPAIR; q++) {
                            const uint32_t abase = smem_u32(sA(stage, pa[q])), bbase = smem_u32(sB(stage, pb[q]));
#pragma unroll
                            for (int k = 0; k < C::BK / 16; k++) {
                                umma_f16_2sm(d_tmem, umma_desc_sw128(abase + k * 32), umma_desc_sw128(bbase + k * 32), idesc, first);
                                first = 1u;
                            }
                        }
                        umma_commit_2sm(&empty_bar[stage], 0x3);                       // free the slot in both CTAs
                        if (kb == a.num_k - 1) umma_commit_2sm(&tmem_full_bar[acc], 0x3);   // wake both epilogues
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; ph ^= 1; }
                }
            }
        }
    } else {
        // =========================== epilogue (warps 2..5, both CTAs) ===========================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int tx = row % a.tw, ty = (row / a.tw) % a.th, tni = row / (a.tw * a.th);
        int it = 0;
        for (int u = cluster_id; u < units; u += num_clusters, it++) {
            const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
            mbar_wait(&tmem_full_bar[acc], acc_ph);
            tc_fence_after();
            const TileCoord t = unit_coord(u);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::ACC_COLS;
            epilogue_tile<BN, NS>(a, t, taddr, row, lane, tx, ty, tni);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&tmem_empty_bar[acc]);
                else mbar_arrive_remote(&tmem_empty_bar[acc], 0);
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();            // neither CTA may exit (or free TMEM) while the pair still has work in flight
    if (warp == 1) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// Halo variant of the CTA-pair kernel: vertical taps share one operand load.
// The chip delivers ~6300 B/clk from L2 to shared memory (42 B/clk/SM), and a split-fp16 MMA group of the plain
// implicit GEMM needs exactly that (a 3x3 / 256-channel layer re-reads every activation tile 9 times), so the plain
// kernels sit at the delivery cap.  Here a tile is tw x th = 16 x 8 (heads: 32 x 4) pixels and ONE TMA box per
// horizontal tap brings th + halo image rows; the vertical taps are views into that box at offsets of tw smem rows
// (tw * 128 B, a multiple of the 1024-byte swizzle atom, so the UMMA descriptor is just advanced).  Operand bytes per
// MMA drop by 3*th/(th+2) on A (2.4x for 3x3; 2.5x for the 5-tap heads); the weights still stream through their own ring.
// The transposed 4x4/s2 convolution runs all four output phases from the same three boxes into four accumulator blocks
// (16 (phase, tap) MMA groups), 3.5x fewer activation bytes than four separate phase GEMMs.
//   warp 0: A producer   warp 1: MMA issuer (leader CTA)   warp 2: weight producer   warps 3-10: epilogue — two warps
//   per TMEM lane quarter, each taking every other 32-column chunk (the epilogue, not the MMA, paces the big layers)
// ------------------------------------------------------------------------------------------------------------
constexpr int HALO_THREADS = 352;          // warps: 0 A producer, 1 MMA issuer, 2 weight producer, 3..10 epilogue
constexpr int HALO_MAX_NA = 4, HALO_MAX_NB = 8;
constexpr int HALO_SMEM_BUDGET = 212 * 1024;

// K walk of the halo kernel.  A ring slot always holds two operand boxes (A: box_rows x 128 B each; B: BN/2 rows each):
//   NS = 2             one pass, 64 channels per step: boxes = fp16 hi, fp16 lo
//   NS = 1, NS = 3 pass 0   128 channels per step (64 for an odd tail): boxes = hi[c0 .. +64), hi[c0+64 .. +128)  -> D1 (kind::f16)
//   NS = 3, pass 1     k8 (128 or 64) channels per step: boxes = e4m3 a8, l8                                  -> D2 (kind::f8f6f4)
// Every role walks the same (pass, step, load, entry) sequence.
struct HaloWalk {
    int npass, k8;
    IPER_DEVINL int steps(int pass, int NS, int cin) const {
        if (NS == 2) return cin / 64;
        return pass == 0 ? (cin + 127) / 128 : cin / k8;
    }
};

template <int BN, int NS>
__global__ void __launch_bounds__(HALO_THREADS, 1) conv_halo_pair_kernel(const __grid_constant__ GemmArgs a) {
    // one weight box of a slot: this CTA's half of a weight tile; the BN = 64 build reserves 2*BN rows because the fused-N
    // transposed conv stages up to two whole 64-row boxes per plane (its half of an N = 256 MMA)
    constexpr int B_PLANE = (BN == 64 ? 2 * BN : BN / 2) * 128;
    // N-concatenated split weights (BN = 32 heads, NS = 2, hs.cat): with N = 32 an MMA is paced by reading its 128-row A
    // operand from shared memory, so hi*hi and hi*lo share ONE read of a_hi: B = [w_hi ; w_lo] (N = 64, CTA r stages all
    // 32 rows of plane r = region X) and only lo*hi runs as a separate N = 32 MMA (region Y = this CTA's half of w_hi).
    // Two MMAs per K step instead of three; D holds [hi*hi + lo*hi | hi*lo] and the epilogue adds the halves.
    constexpr bool CAT_OK = (BN == 32 && NS == 2);
    constexpr int B_SLOT = (CAT_OK ? 3 : 2) * B_PLANE;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t a_full[HALO_MAX_NA];      // leader: 2 arrivals + bytes of both CTAs
    __shared__ __align__(8) uint64_t a_empty[HALO_MAX_NA];     // per CTA, multicast commit
    __shared__ __align__(8) uint64_t b_full[HALO_MAX_NB];
    __shared__ __align__(8) uint64_t b_empty[HALO_MAX_NB];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];        // NS=3: [0] = D1 drained, [1] = D2 drained
    __shared__ uint32_t tmem_base_slot;

    const HaloSched& hs = a.hs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const uint32_t ring = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* ring_ptr = smem_dyn + (ring - smem_u32(smem_dyn));
    const int a_slot = 2 * hs.a_plane_bytes;
    uint8_t* b_ring = ring_ptr + hs.na * a_slot;
    auto sA = [&](int slot, int p) -> uint8_t* { return ring_ptr + slot * a_slot + p * hs.a_plane_bytes; };
    auto sB = [&](int slot, int p) -> uint8_t* { return b_ring + slot * B_SLOT + p * B_PLANE; };
    const int cin = a.cin_chunks * 64;
    HaloWalk walk;
    walk.npass = (NS == 3) ? 2 : 1; walk.k8 = hs.k8;

    if (threadIdx.x == 0) {
        for (int i = 0; i < hs.na; i++) { mbar_init(&a_full[i], 2); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < hs.nb; i++) { mbar_init(&b_full[i], 2); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 16); }
        fence_barrier_init();
    }
    if ((warp == 0 || warp == 2) && lane == 0) {
        for (int p = 0; p < NS; p++) tma_prefetch_desc(warp == 0 ? &a.mapA[p] : &a.mapB[p]);
    }
    if (warp == 1) tmem_alloc_2sm(&tmem_base_slot, (uint32_t)hs.tmem_cols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const bool cat = CAT_OK && hs.cat != 0;
    const int acc_cols = hs.acc_blocks * BN * (cat ? 2 : 1);

    // work unit = (n tile, pair of consecutive M tiles); this CTA owns M tile 2*pair + rank (all phases of it)
    const int m_pairs = (a.m_tiles + 1) >> 1;
    const int units = m_pairs * a.n_tiles;
    auto unit_coord = [&](int u) -> TileCoord {
        TileCoord c;
        c.phase = 0;
        c.n_tile = u % a.n_tiles;
        const int m = (u / a.n_tiles) * 2 + (int)rank;
        if (m >= a.m_tiles) { c.px0 = 0; c.py0 = 0; c.pn0 = a.N; return c; }
        c.px0 = (a.mode == IPER_CONV_ROW5) ? (m % a.tiles_x) * (a.tw - 4) - 2 : (m % a.tiles_x) * a.tw;
        const int r2 = m / a.tiles_x;
        c.py0 = (r2 % a.tiles_y) * a.th;
        c.pn0 = r2 / a.tiles_y;
        return c;
    };

    if (warp == 0) {
        // =========================== activation producer (both CTAs) ===========================
        if (lane == 0) {
            int slot = 0; uint32_t ph = 0;
            for (int u = cluster_id; u < units; u += num_clusters) {
                const TileCoord t = unit_coord(u);
                for (int pass = 0; pass < walk.npass; pass++) {
                    const int nsteps = walk.steps(pass, NS, cin);
                    for (int cc = 0; cc < nsteps; cc++)
                        for (int l = 0; l < hs.n_loads; l++) {
                            mbar_wait(&a_empty[slot], ph ^ 1);
                            const int x = t.px0 + hs.ox[l], y = t.py0 + hs.oy[l];
                            uint32_t bytes;
                            if constexpr (NS == 2) {
                                for (int p = 0; p < 2; p++)
                                    tma_load_4d_2sm(sA(slot, p), &a.mapA[p], &a_full[slot], a.a_coff + cc * 64, x, y, t.pn0);
                                bytes = 2 * hs.a_plane_bytes;
                            } else if (pass == 0) {
                                const int nbox = (cin - cc * 128 >= 128) ? 2 : 1;
                                for (int h = 0; h < nbox; h++)
                                    tma_load_4d_2sm(sA(slot, h), &a.mapA[0], &a_full[slot], a.a_coff + cc * 128 + h * 64, x, y, t.pn0);
                                bytes = nbox * hs.a_plane_bytes;
                            } else {
                                for (int p = 0; p < 2; p++)
                                    tma_load_4d_2sm(sA(slot, p), &a.mapA[1 + p], &a_full[slot], a.a_coff + cc * walk.k8, x, y, t.pn0);
                                bytes = 2 * hs.box_rows * walk.k8;
                            }
                            if (rank == 0) mbar_arrive_expect_tx(&a_full[slot], 2u * bytes);
                            else mbar_arrive_remote(&a_full[slot], 0);
                            if (++slot == hs.na) { slot = 0; ph ^= 1; }
                        }
                }
            }
        }
    } else if (warp == 2) {
        // =========================== weight producer (both CTAs) ===========================
        if (lane == 0) {
            int slot = 0; uint32_t ph = 0;
            for (int u = cluster_id; u < units; u += num_clusters) {
                const int brow = (u % a.n_tiles) * BN + (int)rank * (BN / 2);
                for (int pass = 0; pass < walk.npass; pass++) {
                    const int nsteps = walk.steps(pass, NS, cin);
                    for (int cc = 0; cc < nsteps; cc++)
                        for (int l = 0; l < hs.n_loads; l++)
                            for (int i = hs.first[l]; i < hs.first[l] + hs.count[l]; i++) {
                                mbar_wait(&b_empty[slot], ph ^ 1);
                                const int r = hs.e[i].b_row + brow;
                                uint32_t bytes;
                                constexpr uint32_t HALF_BOX = (BN / 2) * 128, FULL_BOX = BN * 128;
                                if (hs.e[i].nblk > 1) {
                                    // fused N: this CTA's nblk/2 whole boxes, consecutive in every plane of the slot
                                    const HaloEntry& e = hs.e[i];
                                    const int nbx = e.nblk >> 1, nt = (u % a.n_tiles) * BN;
                                    if constexpr (NS == 2) {
                                        for (int p = 0; p < 2; p++)
                                            for (int bx = 0; bx < nbx; bx++)
                                                tma_load_2d_2sm(sB(slot, p) + bx * FULL_BOX, &a.mapBf[p], &b_full[slot],
                                                                e.fb_k[rank][bx] + cc * 64, e.fb_row[rank][bx] + nt);
                                        bytes = 2 * nbx * FULL_BOX;
                                    } else {
                                        const int nbox = (cin - cc * 128 >= 128) ? 2 : 1;
                                        for (int h = 0; h < nbox; h++)
                                            for (int bx = 0; bx < nbx; bx++)
                                                tma_load_2d_2sm(sB(slot, h) + bx * FULL_BOX, &a.mapBf[0], &b_full[slot],
                                                                e.fb_k[rank][bx] + cc * 128 + h * 64, e.fb_row[rank][bx] + nt);
                                        bytes = nbox * nbx * FULL_BOX;
                                    }
                                } else if (cat) {
                                    // X = plane `rank`, all BN rows (two boxes of BN/2 rows); Y = this CTA's half of w_hi
                                    const int r0 = hs.e[i].b_row + (u % a.n_tiles) * BN, k = hs.e[i].b_k + cc * 64;
                                    tma_load_2d_2sm(sB(slot, 0), &a.mapB[rank], &b_full[slot], k, r0);
                                    tma_load_2d_2sm(sB(slot, 1), &a.mapB[rank], &b_full[slot], k, r0 + BN / 2);
                                    tma_load_2d_2sm(sB(slot, 2), &a.mapB[0], &b_full[slot], k, r);
                                    bytes = 3 * HALF_BOX;
                                } else if constexpr (NS == 2) {
                                    for (int p = 0; p < 2; p++)
                                        tma_load_2d_2sm(sB(slot, p), &a.mapB[p], &b_full[slot], hs.e[i].b_k + cc * 64, r);
                                    bytes = 2 * HALF_BOX;
                                } else if (pass == 0) {
                                    const int nbox = (cin - cc * 128 >= 128) ? 2 : 1;
                                    for (int h = 0; h < nbox; h++)
                                        tma_load_2d_2sm(sB(slot, h), &a.mapB[0], &b_full[slot], hs.e[i].b_k + cc * 128 + h * 64, r);
                                    bytes = nbox * HALF_BOX;
                                } else {
                                    for (int p = 0; p < 2; p++)
                                        tma_load_2d_2sm(sB(slot, p), &a.mapB[1 + p], &b_full[slot], hs.e[i].b_k + cc * walk.k8, r);
                                    bytes = 2 * (BN / 2) * walk.k8;
                                }
                                if (rank == 0) mbar_arrive_expect_tx(&b_full[slot], 2u * bytes);
                                else mbar_arrive_remote(&b_full[slot], 0);
                                if (++slot == hs.nb) { slot = 0; ph ^= 1; }
                            }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer (leader CTA only) ===========================
        if (rank == 0) {
            constexpr uint32_t idesc8 = umma_idesc_e4m3(2 * BLOCK_M, BN);
            int aslot = 0, bslot = 0; uint32_t aph = 0, bph = 0; int it = 0;
            for (int u = cluster_id; u < units; u += num_clusters, it++) {
                // NS != 3: two accumulator sets, one per unit parity.  NS == 3: one set [D1 | D2], D1 is handed back early.
                const int acc = (NS == 3) ? 0 : (it & 1);
                const uint32_t acc_ph = (NS == 3) ? (it & 1) : ((it >> 1) & 1);
                const uint32_t d_tmem = tmem_base + acc * acc_cols;
                for (int pass = 0; pass < walk.npass; pass++) {
                    mbar_wait(&tmem_empty_bar[NS == 3 ? pass : acc], acc_ph ^ 1);
                    tc_fence_after();
                    uint32_t touched = 0;            // accumulator blocks already written in this pass
                    const uint32_t d_pass = d_tmem + pass * acc_cols;
                    const int nsteps = walk.steps(pass, NS, cin);
                    for (int cc = 0; cc < nsteps; cc++)
                        for (int l = 0; l < hs.n_loads; l++) {
                            mbar_wait(&a_full[aslot], aph);
                            tc_fence_after();
                            const bool last_load = (pass == walk.npass - 1) && (cc == nsteps - 1) && (l == hs.n_loads - 1);
                            for (int i = hs.first[l]; i < hs.first[l] + hs.count[l]; i++) {
                                mbar_wait(&b_full[bslot], bph);
                                tc_fence_after();
                                const HaloEntry& e = hs.e[i];
                                const bool last_entry = (i == hs.first[l] + hs.count[l] - 1);
                                // fused-N entries (nblk > 1) issue N = nblk * BN; all their blocks are in the same state
                                const uint32_t idesc = umma_idesc_f16(2 * BLOCK_M, (e.nblk > 1 ? e.nblk : 1) * BN);
                                if (elect_one()) {
                                    uint32_t first = (touched >> e.acc) & 1u;
                                    const uint32_t d = d_pass + e.acc * BN * (cat ? 2 : 1);
                                    if (cat) {
                                        constexpr uint32_t idesc_cat = umma_idesc_f16(2 * BLOCK_M, 2 * BN);
                                        const uint32_t ahi = smem_u32(sA(aslot, 0)) + (uint32_t)e.a_row_off * 128u;
                                        const uint32_t alo = smem_u32(sA(aslot, 1)) + (uint32_t)e.a_row_off * 128u;
                                        const uint32_t bx = smem_u32(sB(bslot, 0)), by = smem_u32(sB(bslot, 2));
#pragma unroll
                                        for (int k = 0; k < 4; k++) {       // [hi*hi | hi*lo] in one read of a_hi
                                            umma_f16_2sm(d, umma_desc_sw128(ahi + k * 32), umma_desc_sw128(bx + k * 32), idesc_cat, first);
                                            first = 1u;
                                        }
#pragma unroll
                                        for (int k = 0; k < 4; k++)         // lo*hi into the first BN columns
                                            umma_f16_2sm(d, umma_desc_sw128(alo + k * 32), umma_desc_sw128(by + k * 32), idesc, 1u);
                                    } else if constexpr (NS == 2) {
                                        const int pa[3] = {1, 0, 0};      // lo*hi, hi*lo, hi*hi
                                        const int pb[3] = {0, 1, 0};
#pragma unroll
                                        for (int q = 0; q < 3; q++) {
                                            const uint32_t abase = smem_u32(sA(aslot, pa[q])) + (uint32_t)e.a_row_off * 128u;
                                            const uint32_t bbase = smem_u32(sB(bslot, pb[q]));
#pragma unroll
                                            for (int k = 0; k < 4; k++) {
                                                umma_f16_2sm(d, umma_desc_sw128(abase + k * 32), umma_desc_sw128(bbase + k * 32), idesc, first);
                                                first = 1u;
                                            }
                                        }
                                    } else if (pass == 0) {          // D1 += hi * w_hi over one or two 64-channel boxes
                                        const int nbox = (cin - cc * 128 >= 128) ? 2 : 1;
                                        for (int h = 0; h < nbox; h++) {
                                            const uint32_t abase = smem_u32(sA(aslot, h)) + (uint32_t)e.a_row_off * 128u;
                                            const uint32_t bbase = smem_u32(sB(bslot, h));
#pragma unroll
                                            for (int k = 0; k < 4; k++) {
                                                umma_f16_2sm(d, umma_desc_sw128(abase + k * 32), umma_desc_sw128(bbase + k * 32), idesc, first);
                                                first = 1u;
                                            }
                                        }
                                    } else {                         // D2 += l8 * w8 + a8 * wl8 (e4m3, K = 32 per instruction)
                                        const uint32_t off = (uint32_t)(e.a_row_off * walk.k8);
                                        const uint32_t a8 = smem_u32(sA(aslot, 0)) + off, l8 = smem_u32(sA(aslot, 1)) + off;
                                        const uint32_t w8 = smem_u32(sB(bslot, 0)), wl8 = smem_u32(sB(bslot, 1));
                                        if (walk.k8 == 128) {
#pragma unroll
                                            for (int k = 0; k < 4; k++) {
                                                umma_f8_2sm(d, umma_desc_sw128(l8 + k * 32), umma_desc_sw128(w8 + k * 32), idesc8, first);
                                                umma_f8_2sm(d, umma_desc_sw128(a8 + k * 32), umma_desc_sw128(wl8 + k * 32), idesc8, 1u);
                                                first = 1u;
                                            }
                                        } else {
#pragma unroll
                                            for (int k = 0; k < 2; k++) {
                                                umma_f8_2sm(d, umma_desc_sw64(l8 + k * 32), umma_desc_sw64(w8 + k * 32), idesc8, first);
                                                umma_f8_2sm(d, umma_desc_sw64(a8 + k * 32), umma_desc_sw64(wl8 + k * 32), idesc8, 1u);
                                                first = 1u;
                                            }
                                        }
                                    }
                                    umma_commit_2sm(&b_empty[bslot], 0x3);
                                    if (last_entry) umma_commit_2sm(&a_empty[aslot], 0x3);
                                    if (last_entry && last_load) umma_commit_2sm(&tmem_full_bar[acc], 0x3);
                                }
                                __syncwarp();
                                touched |= ((1u << (e.nblk > 1 ? e.nblk : 1)) - 1u) << e.acc;
                                if (++bslot == hs.nb) { bslot = 0; bph ^= 1; }
                            }
                            if (++aslot == hs.na) { aslot = 0; aph ^= 1; }
                        }
                }
            }
        }
    } else {
        // =========================== epilogue (warps 3..10, both CTAs) ===========================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int hsel = (warp - 3) >> 2;       // which of the two warps of the quarter
        const int row = q * 32 + lane;
        const int tx = row % a.tw, ty = (row / a.tw) % a.th;
        auto release = [&](uint64_t* bar) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(bar);
                else mbar_arrive_remote(bar, 0);
            }
        };
        int it = 0;
        for (int u = cluster_id; u < units; u += num_clusters, it++) {
            const int acc = (NS == 3) ? 0 : (it & 1);
            const uint32_t acc_ph = (NS == 3) ? (it & 1) : ((it >> 1) & 1);
            mbar_wait(&tmem_full_bar[acc], acc_ph);
            tc_fence_after();
            TileCoord t = unit_coord(u);
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + acc * acc_cols;
            uint32_t src = lane_base;
            if constexpr (NS == 3) {
                // fold the fp16 accumulator into the e4m3 one (D2 <- D1 + D2 * cross_scale) so that D1 goes back to the
                // MMA warp at once: the next unit's fp16 pass overlaps the stores below
                for (int j = hsel; j < acc_cols / 32; j += 2) {
                    uint32_t r1[32], r2[32];
                    tmem_ld32(lane_base + j * 32, r1);
                    tmem_ld32(lane_base + acc_cols + j * 32, r2);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i++)
                        r2[i] = __float_as_uint(fmaf(__uint_as_float(r2[i]), a.cross_scale, __uint_as_float(r1[i])));
                    tmem_st32(lane_base + acc_cols + j * 32, r2);
                }
                tmem_st_wait();
                asm volatile("bar.sync 2, 256;" ::: "memory");      // every chunk of D2 holds the folded sum
                release(&tmem_empty_bar[0]);
                src = lane_base + acc_cols;
            }
            for (int blk = 0; blk < hs.acc_blocks; blk++) {
                t.phase = hs.blk_phase[blk];    // transposed conv: the output phase this accumulator block holds
                epilogue_tile<BN, (NS == 3 ? 1 : NS)>(a, t, src + blk * BN * (cat ? 2 : 1), row, lane, tx, ty, 0, hsel, 2);
            }
            release(&tmem_empty_bar[NS == 3 ? 1 : acc]);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2sm(tmem_base, (uint32_t)hs.tmem_cols);
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

// swizzle = bytes of the inner box row: fp16 BK=64 -> 128, fp16 BK=32 -> 64, e4m3 BK=64 -> 64
static int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, bool u8, int row_bytes) {
    PFN_cuTensorMapEncodeTiled_v12000 fn = get_encode_fn();
    IPER_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IPER_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, u8 %d)", (int)r, rank, (int)u8);
    return 0;
}

static int floor_pow2(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }

// Per-device launch state.  Function attributes (opt-in shared memory) and the SM count belong to a DEVICE, and a process
// may drive several (one thread per GPU, or set_device between calls), so nothing here is cached per process: the slot of
// the current device is looked up on every launch.  Slots only ever grow towards the same value, so racing threads that
// both set an attribute are harmless.
constexpr int MAX_DEVICES = 64;
struct DeviceSlot { int dev, sms; };
static int current_device(DeviceSlot& slot) {
    static int sms[MAX_DEVICES] = {};
    int dev = 0;
    IPER_CHECK_CUDA(cudaGetDevice(&dev));
    IPER_REQUIRE(dev >= 0 && dev < MAX_DEVICES, "device ordinal %d out of range", dev);
    if (!sms[dev]) {
        int n = 0;
        IPER_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        sms[dev] = n;
    }
    slot.dev = dev; slot.sms = sms[dev];
    return 0;
}
// opt-in dynamic shared memory of `func` on the current device, raised at most once per (kernel, device, size)
template <typename F>
static int ensure_smem(F func, int (&have)[MAX_DEVICES], int dev, int bytes) {
    if (bytes > have[dev]) {
        IPER_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        have[dev] = bytes;
    }
    return 0;
}
// process-wide experiment switches, read once (not on every launch)
static bool env_flag_default_on(const char* name) {
    const char* v = getenv(name);
    return !(v && atoi(v) == 0);
}
static bool convt_fuse_n_enabled() { static const bool on = env_flag_default_on("IPER_CONVT_FUSE_N"); return on; }
static bool heads_cat_enabled() { static const bool on = env_flag_default_on("IPER_HEADS_CAT"); return on; }

template <int BN, int NS, int TM>
static int launch_gemm(const GemmArgs& g, int max_ctas, cudaStream_t stream) {
    using C = Cfg<BN, NS, TM>;
    static int have[MAX_DEVICES] = {};
    DeviceSlot ds;
    if (int rc = current_device(ds)) return rc;
    if (int rc = ensure_smem(conv_gemm_kernel<BN, NS, TM>, have, ds.dev, C::SMEM_BYTES)) return rc;
    const int num_sms = ds.sms;
    int grid = g.total_tiles < num_sms ? g.total_tiles : num_sms;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    conv_gemm_kernel<BN, NS, TM><<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(g);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int BN, int NS>
static int launch_gemm_pair(const GemmArgs& g, int max_ctas, cudaStream_t stream) {
    using C = Cfg2<BN, NS>;
    static int have[MAX_DEVICES] = {};
    DeviceSlot ds;
    if (int rc = current_device(ds)) return rc;
    if (int rc = ensure_smem(conv_gemm_pair_kernel<BN, NS>, have, ds.dev, C::SMEM_BYTES)) return rc;
    const int num_sms = ds.sms;
    const int units = ((g.m_tiles + 1) / 2) * g.n_tiles * g.phases;
    int clusters = num_sms / 2;
    if (max_ctas > 0 && clusters > max_ctas / 2) clusters = max_ctas / 2 > 0 ? max_ctas / 2 : 1;
    if (clusters > units) clusters = units;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    IPER_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_pair_kernel<BN, NS>, g));
    return 0;
}

template <int BN, int NS>
static int launch_gemm_halo(const GemmArgs& g, int max_ctas, cudaStream_t stream) {
    const int a_slot = 2 * g.hs.a_plane_bytes, b_slot = ((BN == 32 && NS == 2) ? 3 : 2) * ((BN == 64 ? 2 * BN : BN / 2) * 128);
    int smem = g.hs.na * a_slot + g.hs.nb * b_slot + 1024;
    if (smem < 120 * 1024) smem = 120 * 1024;            // one CTA per SM (TMEM allocation)
    static int have[MAX_DEVICES] = {};
    DeviceSlot ds;
    if (int rc = current_device(ds)) return rc;
    if (int rc = ensure_smem(conv_halo_pair_kernel<BN, NS>, have, ds.dev, smem)) return rc;
    const int num_sms = ds.sms;
    const int units = ((g.m_tiles + 1) / 2) * g.n_tiles;
    int clusters = num_sms / 2;
    if (max_ctas > 0 && clusters > max_ctas / 2) clusters = max_ctas / 2 > 0 ? max_ctas / 2 : 1;
    if (clusters > units) clusters = units;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(HALO_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    IPER_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_pair_kernel<BN, NS>, g));
    return 0;
}

// transposed 4x4 s2 p1: spatial offset (oy, ox) of tap (ta, tb) of output phase (py, px) — same algebra as conv_gemm_kernel
static void convt_tap_offset(int phase, int tap, int& oy, int& ox) {
    const int py = phase >> 1, px = phase & 1, ta = tap >> 1, tb = tap & 1;
    oy = py == 0 ? (ta == 0 ? 0 : -1) : (ta == 0 ? 1 : 0);
    ox = px == 0 ? (tb == 0 ? 0 : -1) : (tb == 0 ? 1 : 0);
}

// tap program of the halo kernel for one layer (see HaloSched).  fuse_n (transposed conv, formats 1/2): the output phases
// that read the same view share ONE MMA group (N = 2 or 4 blocks) — the accumulator blocks are ordered cyclically
// [(0,0), (0,1), (1,1), (1,0)]; runs of 2 / 4 adjacent blocks that start on a multiple of their length are fused.
static void build_halo_sched(HaloSched& hs, int mode, int Cin, int rows, int tw, int th, bool fuse_n) {
    hs = HaloSched{};
    for (int i = 0; i < 4; i++) hs.blk_phase[i] = i;
    int ne = 0;
    auto simple = [](int a_row_off, int b_row, int b_k, int acc) {
        HaloEntry e = {};
        e.a_row_off = a_row_off; e.b_row = b_row; e.b_k = b_k; e.acc = acc; e.nblk = 1;
        return e;
    };
    if (mode == IPER_CONV_S1) {                 // 3x3: one box per horizontal tap, three vertical views
        hs.n_loads = 3; hs.acc_blocks = 1; hs.box_rows = (th + 2) * tw;
        for (int dx = 0; dx < 3; dx++) {
            hs.ox[dx] = dx - 1; hs.oy[dx] = -1; hs.first[dx] = ne; hs.count[dx] = 3;
            for (int dy = 0; dy < 3; dy++) hs.e[ne++] = simple(dy * tw, 0, (dy * 3 + dx) * Cin, 0);
        }
    } else if (mode == IPER_CONV_ROW5) {        // heads: one box, five vertical views
        hs.n_loads = 1; hs.acc_blocks = 1; hs.box_rows = (th + 4) * tw;
        hs.ox[0] = 0; hs.oy[0] = -2; hs.first[0] = 0; hs.count[0] = 5;
        for (int dy = 0; dy < 5; dy++) hs.e[ne++] = simple(dy * tw, 0, dy * Cin, 0);
    } else if (!fuse_n) {                       // transposed 4x4 s2: one entry per (phase, tap), block = phase
        hs.n_loads = 3; hs.acc_blocks = 4; hs.box_rows = (th + 2) * tw;
        for (int l = 0; l < 3; l++) {
            hs.ox[l] = l - 1; hs.oy[l] = -1; hs.first[l] = ne;
            for (int phase = 0; phase < 4; phase++)
                for (int tap = 0; tap < 4; tap++) {
                    int oy, ox;
                    convt_tap_offset(phase, tap, oy, ox);
                    if (ox != l - 1) continue;
                    hs.e[ne++] = simple((oy + 1) * tw, phase * rows, tap * Cin, phase);
                }
            hs.count[l] = ne - hs.first[l];
        }
    } else {                                    // transposed 4x4 s2, one entry per VIEW (dy, dx) and run of adjacent blocks
        hs.n_loads = 3; hs.acc_blocks = 4; hs.box_rows = (th + 2) * tw;
        const int order[4] = {0, 1, 3, 2};      // accumulator block -> phase
        for (int i = 0; i < 4; i++) hs.blk_phase[i] = order[i];
        const int load_ox[3] = {0, -1, 1};      // the centre view comes first: it initialises all four blocks at once
        for (int l = 0; l < 3; l++) {
            hs.ox[l] = load_ox[l]; hs.oy[l] = -1; hs.first[l] = ne;
            const int dys[3] = {0, -1, 1};
            for (int di = 0; di < 3; di++) {
                const int dy = dys[di];
                // blocks whose phase has a tap at this view, with that tap
                int tap_of_blk[4];
                for (int b = 0; b < 4; b++) {
                    tap_of_blk[b] = -1;
                    for (int tap = 0; tap < 4; tap++) {
                        int oy, ox;
                        convt_tap_offset(order[b], tap, oy, ox);
                        if (oy == dy && ox == load_ox[l]) tap_of_blk[b] = tap;
                    }
                }
                for (int b = 0; b < 4;) {           // maximal runs of adjacent blocks of length 4, 2 or 1
                    if (tap_of_blk[b] < 0) { b++; continue; }
                    int run = 1;
                    while (b + run < 4 && tap_of_blk[b + run] >= 0) run++;
                    // N must split evenly over the CTA pair, and a group's first TMEM column stays a multiple of its N
                    run = (run == 4 && b == 0) ? 4 : ((run >= 2 && b % 2 == 0) ? 2 : 1);
                    HaloEntry e = {};
                    e.a_row_off = (dy + 1) * tw; e.acc = b; e.nblk = run;
                    e.b_row = order[b] * rows; e.b_k = tap_of_blk[b] * Cin;
                    if (run > 1)
                        for (int r = 0; r < 2; r++)
                            for (int i = 0; i < run / 2; i++) {
                                const int blk = b + r * (run / 2) + i;
                                e.fb_row[r][i] = order[blk] * rows; e.fb_k[r][i] = tap_of_blk[blk] * Cin;
                            }
                    hs.e[ne++] = e;
                    b += run;
                }
            }
            hs.count[l] = ne - hs.first[l];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Stem on the tensor cores without an im2col round trip: Conv2d(CIN <= 7 -> 64, 3x3, s2, p1) + bias + ReLU (+ instance-norm
// statistics) from the fp32 NCHW image (attlwb_spade_resunet.py:268-271).  K = 9*CIN <= 63 is padded to ONE 64-channel
// K step, so a tile is a single [128 px x 64] A operand that no tensor map can describe (9 taps x 6 channels of 9 different
// pixels): four BUILDER warps stage the tile's (33 x 17 x CIN) input patch in shared memory and write the A operand — split
// into fp16 hi / lo — straight into the 128-byte-swizzled K-major layout UMMA reads (16-byte chunk c of row r lives at
// r*128 + ((c ^ (r & 7)) << 4)), make it visible to the async proxy (fence.proxy.async) and arrive on the stage's mbarrier.
// Warp 4 issues the three split-fp16 MMA groups (N = 64) into a double-buffered TMEM accumulator; warps 5-8 run the shared
// epilogue (bias, ReLU, planes store, per-warp running statistics).  The 16 KB weight tile is loaded once per CTA by TMA.
// Per frame this reads the 6.3 MB image once and writes the 33.6 MB output; the im2col + GEMM form (iper_stem_im2col) moved
// another 67 MB through HBM and the CUDA-core stem was shared-memory-LSU bound.
// ------------------------------------------------------------------------------------------------------------
constexpr int STEM_TC_THREADS = 288;      // warps 0-3 builders, 4 MMA issuer, 5-8 epilogue
constexpr int STEM_PW = 40, STEM_PH = 17;  // staged patch: 17 rows x 40 columns (33 needed, start shifted to a 16-byte boundary)
struct alignas(64) StemTcArgs {
    CUtensorMap mapIn;                     // fp32 image as (W, H, CIN, N), box (40, 17, CIN, 1), no swizzle, zero fill
};

template <int NS, int CIN>
__global__ void __launch_bounds__(STEM_TC_THREADS, 1) conv_stem_tc_kernel(const __grid_constant__ GemmArgs a,
                                                                          const __grid_constant__ StemTcArgs sa) {
    constexpr int BN = 64, A_PLANE = BLOCK_M * 128, B_PLANE = BN * 128;
    constexpr int PATCH_FLOATS = CIN * STEM_PH * STEM_PW, PATCH_STRIDE = ((PATCH_FLOATS * 4 + 1023) / 1024) * 1024;
    constexpr int KMAX = 9 * CIN;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], tmem_full_bar[2], tmem_empty_bar[2], patch_bar[2], w_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ double s_pend[4][BN][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* base = smem_dyn + (ring - smem_u32(smem_dyn));
    uint8_t* sB = base;                                    // [NS planes][64 rows x 128 B]
    auto sA = [&](int stage, int p) -> uint8_t* { return base + NS * B_PLANE + (stage * NS + p) * A_PLANE; };
    auto sP = [&](int buf) -> float* { return reinterpret_cast<float*>(base + NS * B_PLANE + 2 * NS * A_PLANE + buf * PATCH_STRIDE); };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&full_bar[i], 128); mbar_init(&empty_bar[i], 1); mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4);
            mbar_init(&patch_bar[i], 1);
        }
        mbar_init(&w_bar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&sa.mapIn);
    }
    if (warp == 4) tmem_alloc(&tmem_base_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const int tile_lo = (int)((long long)a.total_tiles * blockIdx.x / gridDim.x);
    const int tile_hi = (int)((long long)a.total_tiles * (blockIdx.x + 1) / gridDim.x);
    auto coord = [&](int tile) -> TileCoord {
        TileCoord c; c.phase = 0; c.n_tile = 0;
        c.px0 = (tile % a.tiles_x) * 16; c.py0 = ((tile / a.tiles_x) % a.tiles_y) * 8; c.pn0 = tile / (a.tiles_x * a.tiles_y);
        return c;
    };

    if (warp < 4) {
        // =========================== builders: image patch (TMA) -> swizzled split-fp16 A operand ===========================
        const int r = threadIdx.x;                        // A row = pixel of the tile
        const int tx = r & 15, ty = r >> 4;
        auto fetch = [&](int tile, int buf) {             // one thread: the tile's input patch, zero-filled outside the image
            const TileCoord t = coord(tile);
            mbar_arrive_expect_tx(&patch_bar[buf], PATCH_FLOATS * 4);
            // columns start at 2*px0 - 4 (a multiple of 4 floats: TMA's innermost coordinate must stay 16-byte aligned);
            // the tap (ky, kx) of output column tx is patch column 2*tx + kx + 3
            tma_load_4d(sP(buf), &sa.mapIn, &patch_bar[buf], 2 * t.px0 - 4, 2 * t.py0 - 1, 0, t.pn0);
        };
        if (r == 0 && tile_lo < tile_hi) fetch(tile_lo, 0);
        int stage = 0; uint32_t ph = 0; int it = 0;
        for (int tile = tile_lo; tile < tile_hi; tile++, it++) {
            const int buf = it & 1;
            asm volatile("bar.sync 1, 128;" ::: "memory");             // every builder is done with the other patch buffer
            if (r == 0 && tile + 1 < tile_hi) fetch(tile + 1, buf ^ 1);
            mbar_wait(&patch_bar[buf], (it >> 1) & 1);
            mbar_wait(&empty_bar[stage], ph ^ 1);                       // the MMAs that read this A stage have retired
            const float* patch = sP(buf) + (2 * ty) * STEM_PW + 2 * tx + 3;
            uint8_t* ahi = sA(stage, 0) + r * 128;
            uint8_t* alo = (NS == 2) ? sA(stage, 1) + r * 128 : nullptr;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float v[2];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int k = c * 8 + 2 * j + e;                // compile-time after unrolling
                        v[e] = (k < KMAX) ? patch[((k % CIN) * STEM_PH + (k / CIN) / 3) * STEM_PW + (k / CIN) % 3] : 0.f;
                    }
                    const __half h0 = __float2half_rn(v[0]), h1 = __float2half_rn(v[1]);
                    hi[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                    lo[j] = (uint32_t)__half_as_ushort(__float2half_rn(v[0] - __half2float(h0))) |
                            ((uint32_t)__half_as_ushort(__float2half_rn(v[1] - __half2float(h1))) << 16);
                }
                const int sw = (c ^ (r & 7)) << 4;                       // SWIZZLE_128B: chunk index xor row-in-atom
                *reinterpret_cast<uint4*>(ahi + sw) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                if (NS == 2) *reinterpret_cast<uint4*>(alo + sw) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            fence_proxy_async();                                        // generic-proxy stores -> visible to tcgen05.mma
            mbar_arrive(&full_bar[stage]);
            if (++stage == 2) { stage = 0; ph ^= 1; }
        }
    } else if (warp == 4) {
        // =========================== weights (TMA, once) + MMA issuer ===========================
        if (lane == 0) {
            mbar_arrive_expect_tx(&w_bar, NS * B_PLANE);
            for (int p = 0; p < NS; p++) tma_load_2d(sB + p * B_PLANE, &a.mapB[p], &w_bar, 0, 0);
        }
        mbar_wait(&w_bar, 0);
        constexpr uint32_t idesc = umma_idesc_f16(BLOCK_M, BN);
        int stage = 0; uint32_t ph = 0; int it = 0;
        for (int tile = tile_lo; tile < tile_hi; tile++, it++) {
            const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
            mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);
            mbar_wait(&full_bar[stage], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + acc * BN;
                constexpr int NPAIR = (NS == 2) ? 3 : 1;
                const int pa[3] = {NS == 2 ? 1 : 0, 0, 0}, pb[3] = {0, NS == 2 ? 1 : 0, 0};      // lo*hi, hi*lo, hi*hi
                uint32_t first = 0u;
#pragma unroll
                for (int q = 0; q < NPAIR; q++) {
                    const uint32_t ab = smem_u32(sA(stage, pa[q])), bb = smem_u32(sB + pb[q] * B_PLANE);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        umma_f16(d, umma_desc_sw128(ab + k * 32), umma_desc_sw128(bb + k * 32), idesc, first);
                        first = 1u;
                    }
                }
                umma_commit(&empty_bar[stage]);
                umma_commit(&tmem_full_bar[acc]);
            }
            __syncwarp();
            if (++stage == 2) { stage = 0; ph ^= 1; }
        }
    } else {
        // =========================== epilogue (warps 5..8) ===========================
        const int q = warp & 3, row = q * 32 + lane;
        const int tx = row % 16, ty = row / 16;
        double* pend = a.stats_ws ? &s_pend[q][0][0] : nullptr;
        int pend_n = -1, it = 0;
        auto flush = [&]() {
            if (pend_n >= 0 && pend_n < a.N)
                for (int c = lane; c < BN; c += 32) {
                    atomicAdd(a.stats_ws + ((size_t)pend_n * a.rows + c) * 2, pend[c * 2]);
                    atomicAdd(a.stats_ws + ((size_t)pend_n * a.rows + c) * 2 + 1, pend[c * 2 + 1]);
                }
            for (int c = lane; c < BN; c += 32) { pend[c * 2] = 0.0; pend[c * 2 + 1] = 0.0; }
            __syncwarp();
        };
        if (pend) flush();
        for (int tile = tile_lo; tile < tile_hi; tile++, it++) {
            const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
            mbar_wait(&tmem_full_bar[acc], acc_ph);
            tc_fence_after();
            const TileCoord t = coord(tile);
            if (pend && t.pn0 != pend_n) { flush(); pend_n = t.pn0; }
            epilogue_tile<BN, NS>(a, t, tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN, row, lane, tx, ty, 0, 0, 1, pend);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (pend) flush();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 2 * BN);
}

}  // namespace iper

using namespace iper;

// Host-only: the tap program the halo kernel would run for a layer, flattened for inspection (tests/test_abi_cpu.py).
// out[0] = number of A loads, out[1] = accumulator blocks, out[2] = box rows, out[3..6] = phase held by block 0..3; then per
// load 4 ints {ox, oy, first, count}; then per entry 13 ints {a_row_off, b_row, b_k, acc, nblk, fb_row[2][2], fb_k[2][2]}.
// fuse_n selects the fused-N form of the transposed conv.  Returns the number of ints written, or -1.
extern "C" int iper_conv_halo_plan(int mode, int Cin, int rows, int fuse_n, int32_t* out, int capacity) {
    IPER_REQUIRE(out != nullptr && capacity >= 7 + 4 * HALO_MAX_LOADS + 13 * HALO_MAX_ENTRIES, "iper_conv_halo_plan: buffer too small");
    IPER_REQUIRE(mode == IPER_CONV_S1 || mode == IPER_CONVT_4S2 || mode == IPER_CONV_ROW5, "iper_conv_halo_plan: mode %d has no halo form", mode);
    const int tw = mode == IPER_CONV_ROW5 ? 32 : 16, th = BLOCK_M / tw;
    HaloSched hs;
    build_halo_sched(hs, mode, Cin, rows, tw, th, fuse_n != 0 && mode == IPER_CONVT_4S2);
    int n = 0;
    out[n++] = hs.n_loads; out[n++] = hs.acc_blocks; out[n++] = hs.box_rows;
    for (int i = 0; i < 4; i++) out[n++] = hs.blk_phase[i];
    int ne = 0;
    for (int l = 0; l < hs.n_loads; l++) {
        out[n++] = hs.ox[l]; out[n++] = hs.oy[l]; out[n++] = hs.first[l]; out[n++] = hs.count[l];
        ne = hs.first[l] + hs.count[l];
    }
    for (int i = 0; i < ne; i++) {
        const HaloEntry& e = hs.e[i];
        out[n++] = e.a_row_off; out[n++] = e.b_row; out[n++] = e.b_k; out[n++] = e.acc; out[n++] = e.nblk;
        for (int r = 0; r < 2; r++) for (int k = 0; k < 2; k++) out[n++] = e.fb_row[r][k];
        for (int r = 0; r < 2; r++) for (int k = 0; k < 2; k++) out[n++] = e.fb_k[r][k];
    }
    return n;
}

extern "C" int iper_conv_stem_tc(const float* in_nchw, int N, int Cin, int H, int W, const void* w_packed, int w_planes,
                                 long long w_plane_stride, const float* w_scale_inv, const float* bias, void* out, int out_planes,
                                 long long out_plane_stride, int out_pitch, int out_coff, double* stats_ws, iper_stream_t stream) {
    IPER_REQUIRE(in_nchw && w_packed && out, "iper_conv_stem_tc: null pointer");
    IPER_REQUIRE(Cin == 6, "iper_conv_stem_tc: built for the 6-channel encoders of the generator (Cin=%d): use iper_stem_im2col + "
                 "iper_conv_gemm for other channel counts", Cin);
    IPER_REQUIRE(N > 0 && H % 2 == 0 && W % 4 == 0 && H >= 2 && W >= 4, "iper_conv_stem_tc: needs even H and W %% 4 == 0 (got %dx%d)", H, W);
    IPER_REQUIRE(w_planes == 1 || w_planes == 2, "iper_conv_stem_tc: weight format %d not in {1,2}", w_planes);
    IPER_REQUIRE(out_pitch % 8 == 0 && out_coff % 8 == 0 && out_coff + 64 <= out_pitch, "iper_conv_stem_tc: bad output channel window");
    IPER_REQUIRE(((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)in_nchw & 15) == 0, "iper_conv_stem_tc: operands must be 16-byte aligned");
    GemmArgs g = {};
    g.mode = IPER_CONV_S1; g.ksize = 1;
    g.N = N; g.Ho = H / 2; g.Wo = W / 2; g.oH = g.Ho; g.oW = g.Wo;
    g.tw = 16; g.th = 8; g.tn = 1;
    g.tiles_x = (g.Wo + 15) / 16; g.tiles_y = (g.Ho + 7) / 8; g.tiles_nb = N;
    g.m_tiles = g.tiles_x * g.tiles_y * N; g.m_groups = g.m_tiles; g.n_tiles = 1; g.phases = 1; g.total_tiles = g.m_tiles;
    g.rows = 64; g.epi = IPER_EPI_PLANES; g.relu = 1; g.bias = bias;
    g.out = out; g.out_planes = out_planes; g.out_plane_stride = out_plane_stride; g.out_pitch = out_pitch; g.out_coff = out_coff;
    g.stats_ws = stats_ws; g.w_scale_inv = w_scale_inv;
    for (int p = 0; p < w_planes; p++) {
        cuuint64_t wdims[2] = {64, 64};
        cuuint64_t wstr[1] = {64 * 2};
        cuuint32_t wbox[2] = {64, 64};
        if (int rc = encode_map(&g.mapB[p], reinterpret_cast<const __half*>(w_packed) + (size_t)p * w_plane_stride, 2, wdims, wstr, wbox, false, 128)) return rc;
    }
    for (int p = w_planes; p < 3; p++) g.mapB[p] = g.mapB[0];
    for (int p = 0; p < 3; p++) g.mapA[p] = g.mapB[0];
    g.mapBf[0] = g.mapB[0]; g.mapBf[1] = g.mapB[0];
    StemTcArgs sa;
    {   // fp32 image (W, H, C, N), plain (unswizzled) boxes of 40 x 17 x Cin floats, zero fill outside = the conv's padding
        PFN_cuTensorMapEncodeTiled_v12000 fn = get_encode_fn();
        IPER_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Cin, (cuuint64_t)N};
        cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)Cin * H * W * 4};
        cuuint32_t box[4] = {(cuuint32_t)STEM_PW, (cuuint32_t)STEM_PH, (cuuint32_t)Cin, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = fn(&sa.mapIn, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in_nchw), dims, str, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPER_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32 image) failed with CUresult %d", (int)r);
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (stats_ws) IPER_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * (size_t)N * 64, s));
    DeviceSlot ds;
    if (int rc = current_device(ds)) return rc;
    const int patch_stride = ((6 * STEM_PH * STEM_PW * 4 + 1023) / 1024) * 1024;
    const int smem = w_planes * (64 * 128) + 2 * w_planes * (BLOCK_M * 128) + 2 * patch_stride + 1024;
    const int grid = g.total_tiles < ds.sms ? g.total_tiles : ds.sms;
    if (w_planes == 2) {
        static int have[MAX_DEVICES] = {};
        if (int rc = ensure_smem(conv_stem_tc_kernel<2, 6>, have, ds.dev, smem)) return rc;
        conv_stem_tc_kernel<2, 6><<<grid, STEM_TC_THREADS, smem, s>>>(g, sa);
    } else {
        static int have[MAX_DEVICES] = {};
        if (int rc = ensure_smem(conv_stem_tc_kernel<1, 6>, have, ds.dev, smem)) return rc;
        conv_stem_tc_kernel<1, 6><<<grid, STEM_TC_THREADS, smem, s>>>(g, sa);
    }
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_conv_gemm(const iper_conv_gemm_desc* d, iper_stream_t stream) {
    IPER_REQUIRE(d != nullptr, "iper_conv_gemm: null descriptor");
    IPER_REQUIRE(d->a && d->w, "iper_conv_gemm: null operand");
    // a_planes == w_planes, or split-fp16 activations with single-plane weights: a single-pass layer of a mixed-precision
    // plan reads only the hi plane of its input (1 MMA per K step) — the tensor stays usable by split-precision consumers
    const bool hi_only = d->a_planes == 2 && d->w_planes == 1;
    IPER_REQUIRE((d->a_planes == d->w_planes || hi_only) && d->a_planes >= 1 && d->a_planes <= 3,
                 "iper_conv_gemm: a_planes (%d) and w_planes (%d) must be the same format 1, 2 or 3 (or 2 with 1)", d->a_planes, d->w_planes);
    IPER_REQUIRE(d->a_planes != 3 || (d->w8 && d->wl8 && d->cross_scale > 0.f && d->a_pitch % 16 == 0 && d->a_coff % 16 == 0),
                 "iper_conv_gemm: format 3 needs w8, wl8, cross_scale and a 16-aligned channel window");
    IPER_REQUIRE(d->Cin > 0 && d->Cin % 64 == 0, "iper_conv_gemm: Cin=%d must be a multiple of 64", d->Cin);
    // tiles_m = 2: two M tiles per CTA share one weight tile (32-channel K stages).  Measured SLOWER on B200 than one
    // tile with 64-channel stages (64-byte TMA rows deliver fewer bytes per request), so auto = 1; kept selectable.
    const int fmt = hi_only ? 1 : d->a_planes;
    const int TMv = (d->tiles_m == 2 && fmt != 3) ? 2 : 1;
    // cta_pair: 2-CTA clusters (cta_group::2) — each CTA stages its own A tile and half of the weight tile
    const bool halo = d->cta_pair == 2;
    const bool pair = d->cta_pair != 0;
    IPER_REQUIRE(d->cta_pair >= 0 && d->cta_pair <= 2, "iper_conv_gemm: cta_pair=%d not in {0,1,2}", d->cta_pair);
    if (halo) {
        const bool ok_mode = (d->mode == IPER_CONV_S1 && d->ksize == 3 && d->block_n >= 64) ||
                             (d->mode == IPER_CONVT_4S2 && d->block_n == 64) ||
                             (d->mode == IPER_CONV_ROW5 && d->block_n == 32);
        IPER_REQUIRE(TMv == 1 && ok_mode,
                     "iper_conv_gemm: cta_pair=2 (halo) supports 3x3 stride-1 (block_n >= 64), transposed (block_n 64) or "
                     "ROW5 heads (block_n 32) layers");
    } else if (pair) {
        IPER_REQUIRE(fmt != 3 && TMv == 1 && (d->block_n == 128 || d->block_n == 256) && d->epi != IPER_EPI_HEADS &&
                     d->mode != IPER_CONV_ROW5,
                     "iper_conv_gemm: cta_pair supports formats 1/2, tiles_m = 1, block_n 128/256, no heads epilogue");
    }
    IPER_REQUIRE(!(fmt == 3 && TMv == 2), "iper_conv_gemm: format 3 supports tiles_m = 1 only");
    const int BKv = TMv == 2 ? 32 : 64;
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
    if (halo) {                            // 16 x 8 pixel tiles (heads: 32 x 4 with a 4-pixel horizontal overlap), one image each
        g.tw = d->mode == IPER_CONV_ROW5 ? 32 : 16; g.th = BLOCK_M / g.tw; g.tn = 1;
        IPER_REQUIRE(g.Wo >= g.tw && g.Ho >= g.th, "iper_conv_gemm: cta_pair=2 (halo) needs a map of at least %dx%d pixels", g.tw, g.th);
        g.tiles_x = d->mode == IPER_CONV_ROW5 ? (g.Wo + (g.tw - 4) - 1) / (g.tw - 4) : (g.Wo + g.tw - 1) / g.tw;
    }
    g.tiles_y = (g.Ho + g.th - 1) / g.th;
    g.tiles_nb = (g.N + g.tn - 1) / g.tn;
    g.m_tiles = g.tiles_x * g.tiles_y * g.tiles_nb;
    g.m_groups = (g.m_tiles + TMv - 1) / TMv;
    g.n_tiles = d->rows / d->block_n;
    g.total_tiles = g.m_groups * g.n_tiles * g.phases;
    if (halo) {
        // transposed conv in formats 1/2: phases that read the same view share one MMA group (IPER_CONVT_FUSE_N=0 disables)
        const bool fuse_n = d->mode == IPER_CONVT_4S2 && fmt != 3 && convt_fuse_n_enabled();
        build_halo_sched(g.hs, d->mode, d->Cin, d->rows, g.tw, g.th, fuse_n);
        g.hs.a_plane_bytes = g.hs.box_rows * 128;
        g.hs.k8 = (d->Cin % 128 == 0) ? 128 : 64;
        // heads in split fp16: N-concatenated weights (IPER_HEADS_CAT=0 keeps the three-MMA form for comparison)
        g.hs.cat = (d->mode == IPER_CONV_ROW5 && d->block_n == 32 && fmt == 2 && heads_cat_enabled()) ? 1 : 0;
        const int b_plane = (d->block_n == 64 ? 2 * d->block_n : d->block_n / 2) * 128;       // as B_PLANE in the kernel
        const int a_slot = 2 * g.hs.a_plane_bytes, b_slot = (d->block_n == 32 && fmt == 2 ? 3 : 2) * b_plane;
        g.hs.na = (3 * a_slot + 6 * b_slot <= HALO_SMEM_BUDGET) ? 3 : 2;
        g.hs.nb = (HALO_SMEM_BUDGET - g.hs.na * a_slot) / b_slot;
        if (g.hs.nb > HALO_MAX_NB) g.hs.nb = HALO_MAX_NB;
        IPER_REQUIRE(g.hs.nb >= 3, "iper_conv_gemm: halo rings do not fit shared memory");
        int cols = 32;
        while (cols < 2 * g.hs.acc_blocks * d->block_n * (g.hs.cat ? 2 : 1)) cols *= 2;
        IPER_REQUIRE(cols <= 512, "iper_conv_gemm: halo accumulators exceed TMEM");
        g.hs.tmem_cols = cols;
    }
    g.cin_chunks = d->Cin / BKv;
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
    g.cross_scale = d->cross_scale;
    g.stats_ws = d->stats_ws;
    g.w_scale_inv = d->w_scale_inv;
    IPER_REQUIRE(!(d->w_scale_inv && fmt == 3), "iper_conv_gemm: w_scale_inv is not supported with format 3");

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
    const int nmaps = fmt;                       // 1: hi | 2: hi, lo | 3: hi, a8, l8
    const cuuint64_t ktot = (cuuint64_t)taps * d->Cin;
    for (int p = 0; p < nmaps; p++) {
        const bool u8 = (fmt == 3 && p > 0);
        const size_t esz = u8 ? 1 : 2;
        const int kbox = (u8 && halo) ? g.hs.k8 : BKv;          // K elements per box row
        const int row_bytes = (int)(kbox * esz);
        const uint8_t* abase8 = reinterpret_cast<const uint8_t*>(d->a);
        const void* base;
        if (!u8) base = abase8 + (size_t)p * d->a_plane_stride * 2;                       // fp16 plane p
        else base = abase8 + (size_t)d->a_plane_stride * 2 + (size_t)(p - 1) * d->a_plane_stride;   // a8 / l8
        if (d->mode == IPER_CONV_S2) {
            cuuint64_t dims[5] = {(cuuint64_t)2 * d->a_pitch, (cuuint64_t)d->W / 2, 2, (cuuint64_t)d->H / 2, (cuuint64_t)d->N};
            cuuint64_t str[4] = {(cuuint64_t)2 * d->a_pitch * esz, (cuuint64_t)d->W * d->a_pitch * esz,
                                 (cuuint64_t)2 * d->W * d->a_pitch * esz, (cuuint64_t)d->H * d->W * d->a_pitch * esz};
            cuuint32_t box[5] = {(cuuint32_t)BKv, (cuuint32_t)g.tw, 1, (cuuint32_t)g.th, (cuuint32_t)g.tn};
            if (int rc = encode_map(&g.mapA[p], base, 5, dims, str, box, u8, row_bytes)) return rc;
        } else {
            cuuint64_t dims[4] = {(cuuint64_t)d->a_pitch, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
            cuuint64_t str[3] = {(cuuint64_t)d->a_pitch * esz, (cuuint64_t)d->W * d->a_pitch * esz,
                                 (cuuint64_t)d->H * d->W * d->a_pitch * esz};
            cuuint32_t box[4] = {(cuuint32_t)kbox, (cuuint32_t)g.tw, (cuuint32_t)(halo ? g.hs.box_rows / g.tw : g.th), (cuuint32_t)g.tn};
            if (int rc = encode_map(&g.mapA[p], base, 4, dims, str, box, u8, row_bytes)) return rc;
        }
        const void* wb;
        if (!u8) wb = reinterpret_cast<const __half*>(d->w) + (size_t)p * d->w_plane_stride;
        else wb = (p == 1) ? d->w8 : d->wl8;
        cuuint64_t wdims[2] = {ktot, (cuuint64_t)d->rows * g.phases};
        cuuint64_t wstr[1] = {ktot * esz};
        cuuint32_t wbox[2] = {(cuuint32_t)kbox, (cuuint32_t)(pair ? d->block_n / 2 : d->block_n)};
        if (int rc = encode_map(&g.mapB[p], wb, 2, wdims, wstr, wbox, u8, row_bytes)) return rc;
        if (halo && !u8 && p < 2) {             // whole BN-row boxes for the fused-N entries
            cuuint32_t fbox[2] = {(cuuint32_t)kbox, (cuuint32_t)d->block_n};
            if (int rc = encode_map(&g.mapBf[p], wb, 2, wdims, wstr, fbox, false, row_bytes)) return rc;
        }
    }
    for (int p = nmaps; p < 3; p++) { g.mapA[p] = g.mapA[0]; g.mapB[p] = g.mapB[0]; }
    if (!halo) { g.mapBf[0] = g.mapB[0]; g.mapBf[1] = g.mapB[0]; }
    else if (nmaps < 2 || fmt == 3) g.mapBf[1] = g.mapBf[0];

    cudaStream_t s = (cudaStream_t)stream;
    if (d->stats_ws) {
        IPER_REQUIRE(d->epi == IPER_EPI_PLANES && d->mode != IPER_CONVT_4S2, "iper_conv_gemm: fused statistics need the planes epilogue of a (strided) conv");
        IPER_CHECK_CUDA(cudaMemsetAsync(d->stats_ws, 0, sizeof(double) * 2 * (size_t)d->N * d->rows, s));
    }
    if (halo) {
        switch (d->block_n) {
#define IPER_HALO_CASE(BNV)                                                                                           \
    case BNV:                                                                                                        \
        return fmt == 3 ? launch_gemm_halo<BNV, 3>(g, d->max_ctas, s)                                                \
                        : (fmt == 2 ? launch_gemm_halo<BNV, 2>(g, d->max_ctas, s) : launch_gemm_halo<BNV, 1>(g, d->max_ctas, s));
            IPER_HALO_CASE(32) IPER_HALO_CASE(64) IPER_HALO_CASE(128)
            default: return fmt == 3 ? launch_gemm_halo<256, 3>(g, d->max_ctas, s)
                                     : (fmt == 2 ? launch_gemm_halo<256, 2>(g, d->max_ctas, s) : launch_gemm_halo<256, 1>(g, d->max_ctas, s));
#undef IPER_HALO_CASE
        }
    }
    if (pair) {
        if (d->block_n == 256) return fmt == 2 ? launch_gemm_pair<256, 2>(g, d->max_ctas, s) : launch_gemm_pair<256, 1>(g, d->max_ctas, s);
        return fmt == 2 ? launch_gemm_pair<128, 2>(g, d->max_ctas, s) : launch_gemm_pair<128, 1>(g, d->max_ctas, s);
    }
#define IPER_DISPATCH(BNV)                                                                                           \
    do {                                                                                                             \
        if (fmt == 3) return launch_gemm<BNV, 3, 1>(g, d->max_ctas, s);                                              \
        if (fmt == 2) return TMv == 2 ? launch_gemm<BNV, 2, 2>(g, d->max_ctas, s) : launch_gemm<BNV, 2, 1>(g, d->max_ctas, s); \
        return TMv == 2 ? launch_gemm<BNV, 1, 2>(g, d->max_ctas, s) : launch_gemm<BNV, 1, 1>(g, d->max_ctas, s);       \
    } while (0)
    switch (d->block_n) {
        case 32: IPER_DISPATCH(32);
        case 64: IPER_DISPATCH(64);
        case 128: IPER_DISPATCH(128);
        default: IPER_DISPATCH(256);
    }
#undef IPER_DISPATCH
}
