// Training-step kernels (BASELINE.json configs[4]): bf16 tcgen05 kernels for the stride-1 "same"
// convolutions of the step (1x1 attention projections, 3x3 ResidualBlock / SPADE / skip / VGG convs, 5x5 heads, 7x7 BGNet
// ends) — forward, data gradient and weight gradient — plus the bias gradient and the fused Adam + weight-repack pass, behind
// ipercore_b200/train.py.  The reference trains these layers through cuDNN and torch.optim.Adam
// (iPERCore/tools/trainers/lwg_trainer.py:326-352, 699-833 on attlwb_spade_resunet.py:14-25, 80-93, 208-252, 316-357, 605-613).
//
// Everything is NHWC bf16 (torch channels_last), so no tensor is ever transposed for a kernel:
//   forward / dgrad  implicit GEMM  D[128 px, BN co] += A[128 px, 64 ci] B[BN co, 64 ci]^T  per (tap, 64-channel chunk); A is one
//                    TMA box per tap (OOB zero fill = padding), both operands K-major.  dgrad is the same kernel on dY with the
//                    180-degree-rotated, in/out-transposed weights.  Epilogue: bias, residual add, ReLU, bf16 store.
//   wgrad            dW[co, tap, ci] = sum_px dY[px, co] X[px + tap, ci] contracts over PIXELS.  In NHWC a pixel is a 128-byte row
//                    of 64 channels, i.e. both operands are MN-major: the TMA box (64 ch, 16 px, 8 rows) lands as 128 K-rows of
//                    128 bytes with the 128-byte swizzle, which is exactly tcgen05's canonical MN-major SWIZZLE_128B layout
//                    ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units with SBO = 1024 (8 pixels) and LBO = the distance between
//                    64-channel groups (one box = 16 KB).  Instruction descriptor bits 15/16 select MN-major A and B.  The tap
//                    shift is a TMA coordinate in both x and y (the first version of this file read NCHW, where the horizontal
//                    shift is a 2-byte offset of the innermost dimension, which TMA cannot address, and paid two transposes and
//                    three shifted copies per layer).  One CTA per (128-channel row tile, BN-channel column tile, tap, K split);
//                    fp32 atomics combine the splits straight into the caller's gradient buffer (any element strides, so the
//                    flat fp32 gradient of the parameter itself can be the target).
//   bias grad        column sums of dY, fp32 atomics into the gradient buffer.
//   Adam + repack    one pass over the flat fp32 parameter / gradient / moment buffers (torch.optim.Adam's update, bias
//                    correction from a device-side step counter so the pass can live in a CUDA graph) that also writes the bf16
//                    K-major forward and dgrad packings of every convolution weight the kernels above consume.
// One CTA per output tile, 192 threads: warp 0 TMA producer, warp 1 MMA issuer (single thread), warps 2-5 epilogue.
#include <algorithm>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "iper_b200.h"

namespace iper {

constexpr int TR_THREADS = 192;
constexpr int TR_BOX = 128 * 128;            // one activation box: 128 pixels x 64 channels x 2 bytes

struct alignas(64) TrainArgs {
    CUtensorMap mapA, mapB;
    int N, H, W, Cin, Cout, ks, pad;     // conv geometry (forward: Cin -> Cout)
    int tiles_x, tiles_y, n_tiles;
    int steps;                           // forward: ks*ks * Cin/64
    const float* bias; int relu;
    const __nv_bfloat16* add;            // residual (same shape and pitch as out) or null
    __nv_bfloat16* out; int out_pitch;
    // wgrad: mapA = the tensor whose channels are the 128 accumulator rows, mapB = the BN accumulator columns
    int x_rows;                          // 0: rows = dY channels (co), cols = X channels (ci);  1: rows = ci, cols = co
    int rowsC, colsC, col_tiles, splitk, kblocks, taps;
    float* dW; long long s_row, s_col, s_tap;   // element strides of dW for (row channel, column channel, tap)
};

// [4,6) D fmt = f32, [7,10) A fmt, [10,13) B fmt (1 = bf16), [15] A major, [16] B major (1 = MN-major), [17,23) N>>3, [24,29) M>>4
IPER_DEVINL constexpr uint32_t umma_idesc_bf16(int M, int N, int mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major operand, SWIZZLE_128B: K rows of 128 bytes (64 MN elements), 8-row groups every 1024 bytes (SBO), 64-element MN groups
// every `lbo` bytes.
IPER_DEVINL uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// MODE 0: forward / dgrad (K-major operands, stage = A box + BN weight rows); MODE 1: wgrad (MN-major operands, stage = two row
// boxes + BN/64 column boxes)
template <int BN, int MODE>
struct TrainCfg {
    static constexpr int A_BYTES = MODE == 0 ? TR_BOX : 2 * TR_BOX;
    static constexpr int B_BYTES = MODE == 0 ? BN * 128 : (BN / 64) * TR_BOX;
    static constexpr int STAGE = A_BYTES + B_BYTES;
    // forward / dgrad: <= 96 KB of ring so that TWO CTAs share an SM — at batch 1 the tiles are short (9-49 K steps) and one CTA's
    // prologue / epilogue hides behind the other's main loop; wgrad keeps the whole SM (64 KB stages)
    static constexpr int BUDGET = MODE == 0 ? 96 * 1024 : 200 * 1024;
    static constexpr int STAGES = BUDGET / STAGE > 8 ? 8 : BUDGET / STAGE;
    static constexpr int SMEM = STAGES * STAGE + 1024;
};

template <int BN, int MODE>
__global__ void __launch_bounds__(TR_THREADS, MODE == 0 ? 2 : 1) train_gemm_kernel(const __grid_constant__ TrainArgs a) {
    using Cfg = TrainCfg<BN, MODE>;
    constexpr int STAGES = Cfg::STAGES, STAGE = Cfg::STAGE;
    constexpr int TCOLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* ring_ptr = smem_dyn + (ring - smem_u32(smem_dyn));
    auto sA = [&](int s) { return ring_ptr + s * STAGE; };
    auto sB = [&](int s) { return ring_ptr + s * STAGE + Cfg::A_BYTES; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&a.mapA); tma_prefetch_desc(&a.mapB); }
    if (warp == 1) tmem_alloc(&tmem_slot, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    // ---- work item decode ----
    int steps, k0 = 0;
    int n_tile = 0, px0 = 0, py0 = 0, pn = 0;            // forward
    int row0 = 0, col0 = 0, tap = 0;                     // wgrad
    if (MODE == 0) {
        n_tile = blockIdx.x % a.n_tiles;
        const int m = blockIdx.x / a.n_tiles;
        px0 = (m % a.tiles_x) * 16; py0 = ((m / a.tiles_x) % a.tiles_y) * 8; pn = m / (a.tiles_x * a.tiles_y);
        steps = a.steps;
    } else {
        int u = blockIdx.x;
        const int split = u % a.splitk; u /= a.splitk;
        tap = u % a.taps; u /= a.taps;
        col0 = (u % a.col_tiles) * BN; row0 = (u / a.col_tiles) * 128;
        const int per = (a.kblocks + a.splitk - 1) / a.splitk;
        k0 = split * per;
        steps = min(per, a.kblocks - k0);
        if (steps < 0) steps = 0;
    }
    const int cin_chunks = a.Cin / 64;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < steps; i++) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], STAGE);
                if (MODE == 0) {
                    const int t = i / cin_chunks, cc = i - t * cin_chunks;
                    tma_load_4d(sA(s), &a.mapA, &full_bar[s], cc * 64, px0 + t % a.ks - a.pad, py0 + t / a.ks - a.pad, pn);
                    tma_load_2d(sB(s), &a.mapB, &full_bar[s], i * 64, n_tile * BN);
                } else {
                    const int j = k0 + i;
                    const int x0 = (j % a.tiles_x) * 16, y0 = ((j / a.tiles_x) % a.tiles_y) * 8, n = j / (a.tiles_x * a.tiles_y);
                    const int dx = tap % a.ks - a.pad, dy = tap / a.ks - a.pad;      // X is read at the output pixel + tap offset
                    const int rx = a.x_rows ? dx : 0, ry = a.x_rows ? dy : 0, cx = a.x_rows ? 0 : dx, cy = a.x_rows ? 0 : dy;
#pragma unroll
                    for (int g = 0; g < 2; g++)        // channels beyond the tensor are zero-filled by TMA (64-channel tensors)
                        tma_load_4d(sA(s) + g * TR_BOX, &a.mapA, &full_bar[s], row0 + g * 64, x0 + rx, y0 + ry, n);
#pragma unroll
                    for (int g = 0; g < BN / 64; g++)
                        tma_load_4d(sB(s) + g * TR_BOX, &a.mapB, &full_bar[s], col0 + g * 64, x0 + cx, y0 + cy, n);
                }
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, BN, MODE);
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < steps; i++) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t ab = smem_u32(sA(s)), bb = smem_u32(sB(s));
                if (MODE == 0) {
#pragma unroll
                    for (int k = 0; k < 4; k++)        // 64 channels = 4 x K16, 32 bytes along the swizzled row
                        umma_f16(tmem, umma_desc_sw128(ab + k * 32), umma_desc_sw128(bb + k * 32), idesc, (i > 0 || k > 0) ? 1u : 0u);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; k++)        // 128 pixels = 8 x K16, 16 rows of 128 bytes each
                        umma_f16(tmem, umma_desc_mn_sw128(ab + k * 2048, TR_BOX), umma_desc_mn_sw128(bb + k * 2048, TR_BOX), idesc,
                                 (i > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
                if (i == steps - 1) umma_commit(&acc_bar);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (steps > 0) {
        const int q = warp & 3, row = q * 32 + lane;
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
        if (MODE == 0) {
            const int x = px0 + row % 16, y = py0 + row / 16;
            const bool valid = x < a.W && y < a.H && pn < a.N;
            const size_t off = (((size_t)pn * a.H + y) * a.W + x) * a.out_pitch + n_tile * BN;
#pragma unroll 1
            for (int j = 0; j < BN / 32; j++) {
                uint32_t r[32];
                tmem_ld32(taddr + j * 32, r);
                tmem_ld_wait();
                if (valid) {
                    uint4 res[4];
                    if (a.add) {
                        const uint4* ap = reinterpret_cast<const uint4*>(a.add + off + j * 32);
#pragma unroll
                        for (int i = 0; i < 4; i++) res[i] = __ldg(ap + i);
                    }
                    const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(res);
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        float v0 = __uint_as_float(r[2 * i]), v1 = __uint_as_float(r[2 * i + 1]);
                        if (a.bias) { v0 += __ldg(a.bias + n_tile * BN + j * 32 + 2 * i); v1 += __ldg(a.bias + n_tile * BN + j * 32 + 2 * i + 1); }
                        if (a.add) { const float2 f = __bfloat1622float2(rb[i]); v0 += f.x; v1 += f.y; }
                        if (a.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                        pk[i] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(a.out + off + j * 32);
#pragma unroll
                    for (int i = 0; i < 4; i++) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                }
            }
        } else {
            const int rr = row0 + row;
            float* base = a.dW + (long long)rr * a.s_row + (long long)tap * a.s_tap;
#pragma unroll 1
            for (int j = 0; j < BN / 32; j++) {
                uint32_t r[32];
                tmem_ld32(taddr + j * 32, r);
                tmem_ld_wait();
                if (rr < a.rowsC) {
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        const int cc = col0 + j * 32 + i;
                        if (cc < a.colsC) atomicAdd(base + (long long)cc * a.s_col, __uint_as_float(r[i]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TCOLS);
}

// db[c] += sum over pixels of dy[px, c] (NHWC bf16, `pitch` channels per pixel, the first C of them reduced); 8 channels per thread
__global__ void __launch_bounds__(256) bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, long long pixels, int C, int pitch,
                                                        float* __restrict__ db) {
    __shared__ float red[256][9];
    const int groups = (C + 7) / 8, lanes = 256 / groups;       // groups <= 256
    const int cg = threadIdx.x % groups, pl = threadIdx.x / groups;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (pl < lanes) {
        for (long long p = (long long)blockIdx.x * lanes + pl; p < pixels; p += (long long)gridDim.x * lanes) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(dy + p * pitch + cg * 8));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int i = 0; i < 4; i++) { const float2 f = __bfloat1622float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) red[threadIdx.x][i] = acc[i];
    __syncthreads();
    if (threadIdx.x < groups * 8) {
        const int g = threadIdx.x / 8, i = threadIdx.x % 8;
        float s = 0.f;
        for (int l = 0; l < lanes; l++) s += red[l * groups + g][i];
        if (g * 8 + i < C) atomicAdd(db + g * 8 + i, s);
    }
}

// torch.optim.Adam (no weight decay, no amsgrad) over flat fp32 buffers + bf16 repacking of the convolution weights
__global__ void __launch_bounds__(256) adam_pack_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, const iper_adam_seg* __restrict__ segs,
                                                        const int2* __restrict__ chunks, int chunk_elems, float lr, float b1, float b2,
                                                        float eps, float grad_scale, const float* __restrict__ step_ptr, int update,
                                                        __nv_bfloat16* __restrict__ pack_fwd, __nv_bfloat16* __restrict__ pack_dgrad) {
    const int2 ch = chunks[blockIdx.x];
    const iper_adam_seg sg = segs[ch.x];
    float inv_bc1 = 1.f, inv_sqrt_bc2 = 1.f;
    if (update) {
        const float t = *step_ptr;
        inv_bc1 = 1.f / (1.f - powf(b1, t));
        inv_sqrt_bc2 = rsqrtf(1.f - powf(b2, t));
    }
    const long long end = min((long long)ch.y + chunk_elems, sg.numel);
    for (long long e = (long long)ch.y + threadIdx.x; e < end; e += 256) {
        const long long idx = sg.offset + e;
        float pp = p[idx];
        int tap = 0, ci = 0, co = 0;
        long long gidx = idx;
        if (sg.taps > 0) {                           // convolution weight (co, ci, ky, kx); its gradient is stored (co, tap, ci)
            tap = (int)(e % sg.taps);
            const long long cc = e / sg.taps;
            ci = (int)(cc % sg.ci); co = (int)(cc / sg.ci);
            gidx = sg.offset + ((long long)co * sg.taps + tap) * sg.ci + ci;
        }
        if (update) {
            const float gr = g[gidx] * grad_scale;
            const float mm = b1 * m[idx] + (1.f - b1) * gr;
            const float vv = b2 * v[idx] + (1.f - b2) * gr * gr;
            m[idx] = mm; v[idx] = vv;
            pp -= lr * inv_bc1 * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
            p[idx] = pp;
        }
        if (sg.taps > 0) {                           // (co, ci, ky, kx) -> forward (co_pad, tap, ci_pad) and dgrad (ci_pad, taps-1-tap, co_pad)
            const __nv_bfloat16 h = __float2bfloat16(pp);
            pack_fwd[sg.fwd_offset + ((long long)co * sg.taps + tap) * sg.ci_pad + ci] = h;
            if (sg.dgrad_offset >= 0)
                pack_dgrad[sg.dgrad_offset + ((long long)ci * sg.taps + (sg.taps - 1 - tap)) * sg.co_pad + co] = h;
        }
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 train_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}
static int bf16_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
    auto fn = train_encode_fn();
    IPER_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IPER_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (bf16, rank %d) failed with CUresult %d", rank, (int)r);
    return 0;
}
// NHWC activation (C, W, H, N), box = 64 channels x 16 x 8 pixels of one image
static int nhwc_map(CUtensorMap* m, const void* base, int N, int H, int W, int C) {
    cuuint64_t d[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t s[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t b[4] = {64, 16, 8, 1};
    return bf16_map(m, base, 4, d, s, b);
}

template <int BN, int MODE>
static int launch_train(const TrainArgs& t, int grid, cudaStream_t st) {
    constexpr int SMEM = TrainCfg<BN, MODE>::SMEM;
    static int have[64] = {};
    int dev = 0;
    IPER_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && have[dev] < SMEM) {
        IPER_CHECK_CUDA(cudaFuncSetAttribute(train_gemm_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        have[dev] = SMEM;
    }
    train_gemm_kernel<BN, MODE><<<grid, TR_THREADS, SMEM, st>>>(t);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace iper

using namespace iper;

extern "C" int iper_conv_bf16(const void* x_nhwc, int N, int H, int W, int Cin, const void* w_packed, int Cout, int ksize,
                              const float* bias, int relu, const void* add_nhwc, void* out_nhwc, iper_stream_t stream) {
    IPER_REQUIRE(x_nhwc && w_packed && out_nhwc, "iper_conv_bf16: null pointer");
    IPER_REQUIRE(ksize >= 1 && ksize <= 7 && (ksize & 1), "iper_conv_bf16: kernel size %d (odd, <= 7: stride-1 'same' convolutions)", ksize);
    IPER_REQUIRE(N > 0 && H >= 8 && W >= 16 && Cin % 64 == 0 && Cout % 64 == 0, "iper_conv_bf16: needs H >= 8, W >= 16, Cin %% 64 == 0, Cout %% 64 == 0 "
                 "(got %dx%d, %d -> %d)", H, W, Cin, Cout);
    IPER_REQUIRE(((uintptr_t)x_nhwc & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)out_nhwc & 15) == 0 &&
                 ((uintptr_t)add_nhwc & 15) == 0, "iper_conv_bf16: 16-byte alignment");
    TrainArgs t = {};
    t.N = N; t.H = H; t.W = W; t.Cin = Cin; t.Cout = Cout; t.ks = ksize; t.pad = ksize / 2;
    t.tiles_x = (W + 15) / 16; t.tiles_y = (H + 7) / 8;
    const int BN = Cout % 128 == 0 ? 128 : 64;
    t.n_tiles = Cout / BN; t.steps = ksize * ksize * (Cin / 64);
    t.bias = bias; t.relu = relu; t.add = reinterpret_cast<const __nv_bfloat16*>(add_nhwc);
    t.out = reinterpret_cast<__nv_bfloat16*>(out_nhwc); t.out_pitch = Cout;
    if (int rc = nhwc_map(&t.mapA, x_nhwc, N, H, W, Cin)) return rc;
    cuuint64_t bd[2] = {(cuuint64_t)ksize * ksize * Cin, (cuuint64_t)Cout};
    cuuint64_t bs[1] = {(cuuint64_t)ksize * ksize * Cin * 2};
    cuuint32_t bb[2] = {64, (cuuint32_t)BN};
    if (int rc = bf16_map(&t.mapB, w_packed, 2, bd, bs, bb)) return rc;
    const int grid = t.tiles_x * t.tiles_y * N * t.n_tiles;
    return BN == 128 ? launch_train<128, 0>(t, grid, (cudaStream_t)stream) : launch_train<64, 0>(t, grid, (cudaStream_t)stream);
}

extern "C" int iper_conv_wgrad_bf16(const void* x_nhwc, const void* dy_nhwc, int N, int H, int W, int Cin, int Cout, int ksize,
                                    float* dW, long long stride_co, long long stride_ci, long long stride_tap, int co_valid,
                                    int ci_valid, iper_stream_t stream) {
    IPER_REQUIRE(x_nhwc && dy_nhwc && dW, "iper_conv_wgrad_bf16: null pointer");
    IPER_REQUIRE(ksize >= 1 && ksize <= 7 && (ksize & 1), "iper_conv_wgrad_bf16: kernel size %d (odd, <= 7)", ksize);
    IPER_REQUIRE(N > 0 && H > 0 && W > 0 && Cin % 64 == 0 && Cout % 64 == 0,
                 "iper_conv_wgrad_bf16: needs Cin %% 64 == 0, Cout %% 64 == 0 (got %dx%d, %d -> %d)", H, W, Cin, Cout);
    IPER_REQUIRE(((uintptr_t)x_nhwc & 15) == 0 && ((uintptr_t)dy_nhwc & 15) == 0, "iper_conv_wgrad_bf16: 16-byte alignment");
    IPER_REQUIRE(co_valid > 0 && co_valid <= Cout && ci_valid > 0 && ci_valid <= Cin, "iper_conv_wgrad_bf16: valid channel counts out of range");
    cudaStream_t st = (cudaStream_t)stream;
    TrainArgs t = {};
    t.N = N; t.H = H; t.W = W; t.Cin = Cin; t.Cout = Cout; t.ks = ksize; t.pad = ksize / 2; t.taps = ksize * ksize; t.dW = dW;
    t.tiles_x = (W + 15) / 16; t.tiles_y = (H + 7) / 8;
    t.kblocks = t.tiles_x * t.tiles_y * N;
    // accumulator rows = TMEM lanes = the 32 threads of a warp in the epilogue: put the channel index with the SMALLER dW stride
    // there, so one warp-wide atomic instruction touches consecutive floats (one 128-byte line for stride 1) instead of 32 lines
    t.x_rows = stride_ci <= stride_co ? 1 : 0;
    const int rowsP = t.x_rows ? Cin : Cout, colsP = t.x_rows ? Cout : Cin;          // padded (tensor) channel counts
    t.rowsC = t.x_rows ? ci_valid : co_valid; t.colsC = t.x_rows ? co_valid : ci_valid;
    t.s_row = t.x_rows ? stride_ci : stride_co; t.s_col = t.x_rows ? stride_co : stride_ci; t.s_tap = stride_tap;
    const int BN = colsP % 128 == 0 ? 128 : 64;
    const int row_tiles = (rowsP + 127) / 128;
    t.col_tiles = colsP / BN;
    const int items = row_tiles * t.col_tiles * t.taps;
    int sk = (148 * 2 + items - 1) / items;                // fill the machine about twice over
    if (sk > t.kblocks) sk = t.kblocks;
    if (sk < 1) sk = 1;
    t.splitk = sk;
    if (int rc = nhwc_map(&t.mapA, t.x_rows ? x_nhwc : dy_nhwc, N, H, W, rowsP)) return rc;
    if (int rc = nhwc_map(&t.mapB, t.x_rows ? dy_nhwc : x_nhwc, N, H, W, colsP)) return rc;
    const int grid = items * sk;
    return BN == 128 ? launch_train<128, 1>(t, grid, st) : launch_train<64, 1>(t, grid, st);
}

extern "C" int iper_bias_grad_bf16(const void* dy_nhwc, long long pixels, int C, int pitch, float* db, iper_stream_t stream) {
    IPER_REQUIRE(dy_nhwc && db, "iper_bias_grad_bf16: null pointer");
    IPER_REQUIRE(pixels > 0 && C > 0 && C <= pitch && pitch % 8 == 0 && (C + 7) / 8 <= 256 && ((uintptr_t)dy_nhwc & 15) == 0,
                 "iper_bias_grad_bf16: needs C <= pitch, pitch %% 8 == 0, C <= 2048, 16-byte alignment (got C %d pitch %d)", C, pitch);
    const int lanes = 256 / ((C + 7) / 8);
    long long blocks = (pixels + (long long)lanes * 32 - 1) / ((long long)lanes * 32);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    bias_grad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy_nhwc), pixels, C, pitch, db);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_adam_pack(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const iper_adam_seg* segs_dev,
                              const int* chunks_dev, int n_chunks, int chunk_elems, float lr, float beta1, float beta2, float eps,
                              float grad_scale, const float* step_dev, int update, void* pack_fwd, void* pack_dgrad,
                              iper_stream_t stream) {
    IPER_REQUIRE(params && segs_dev && chunks_dev && n_chunks > 0 && chunk_elems > 0, "iper_adam_pack: null pointer / empty chunk table");
    IPER_REQUIRE(!update || (grads && exp_avg && exp_avg_sq && step_dev), "iper_adam_pack: update needs gradients, both moments and the step counter");
    adam_pack_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, segs_dev,
                                                                reinterpret_cast<const int2*>(chunks_dev), chunk_elems, lr, beta1, beta2,
                                                                eps, grad_scale, step_dev, update,
                                                                reinterpret_cast<__nv_bfloat16*>(pack_fwd),
                                                                reinterpret_cast<__nv_bfloat16*>(pack_dgrad));
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}
