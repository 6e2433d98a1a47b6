// Training-step kernels (BASELINE.json configs[4]): bf16 tcgen05 kernels for the 3x3 / stride-1
// convolutions of the generator — forward, data gradient and weight gradient — behind ipercore_b200/train.py's
// autograd.Function.  These are the layers that carry ~60 % of the generator's FLOPs (ResidualBlocks, SPADE MLPs, skip
// convolutions: attlwb_spade_resunet.py:14-25, 80-93, 316-357); the reference trains them through cuDNN
// (iPERCore/tools/trainers/lwg_trainer.py:699-833).
//
//   forward / dgrad  implicit GEMM over NHWC bf16 (torch channels_last): D[128 px, BN co] += A[128 px, 64 ci] B[BN co, 64 ci]^T
//                    per (tap, 64-channel chunk), A by one TMA box per tap (OOB zero fill = padding) — dgrad is the same
//                    kernel on the 180-degree-rotated, in/out-transposed weights.
//   wgrad            dW[co, tap, ci] = sum_{n,y,x} dY[n,co,y,x] X[n,ci,y+dy,x+dx]: the contraction runs over PIXELS, so both
//                    operands are read from NCHW tensors, where a row segment of 64 pixels of one channel is a K-major
//                    128-byte row: A = 128 channels of dY, B = BN channels of X shifted by the tap.  The vertical shift is a
//                    TMA coordinate (zero fill = padding); the horizontal one cannot be — the innermost TMA coordinate must
//                    stay 16-byte aligned (a +-1 pixel start faults) — so a small pre-pass writes the three x-shifted copies
//                    [X(x-1) | X | X(x+1)] into a caller-owned workspace and the tap picks its copy through a 5th tensor
//                    dimension.  One CTA per (row tile, column tile, tap, K split); fp32 atomics combine the splits.
// One CTA per output tile, 192 threads: warp 0 TMA producer, warp 1 MMA issuer (single thread), warps 2-5 epilogue.
#include <algorithm>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "iper_b200.h"

namespace iper {

constexpr int TR_THREADS = 192, TR_STAGES = 4;

struct alignas(64) TrainArgs {
    CUtensorMap mapA, mapB;
    int mode;                    // 0 forward / dgrad (NHWC), 1 wgrad (NCHW)
    int N, H, W, Cin, Cout;      // conv geometry (forward: Cin -> Cout)
    int tiles_x, tiles_y, n_tiles;
    int steps;                   // forward: 9 * Cin/64
    const float* bias; int relu;
    __nv_bfloat16* out; int out_pitch;
    // wgrad
    int a_shift;                 // 0: A = dY (rows = co), B = X shifted (cols = ci);  1: A = X shifted (rows = ci), B = dY (cols = co)
    int row_tiles, col_tiles, splitk, ksteps, rc;     // rc = W / 64 row chunks
    float* dW;                   // (Cout, 9, Cin) fp32, accumulated with atomics
};

IPER_DEVINL constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN>
__global__ void __launch_bounds__(TR_THREADS, 1) train_gemm_kernel(const __grid_constant__ TrainArgs a) {
    constexpr int A_TILE = 128 * 128, B_TILE = BN * 128, STAGE = A_TILE + B_TILE;
    constexpr int TCOLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[TR_STAGES], empty_bar[TR_STAGES], acc_bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* ring_ptr = smem_dyn + (ring - smem_u32(smem_dyn));
    auto sA = [&](int s) { return ring_ptr + s * STAGE; };
    auto sB = [&](int s) { return ring_ptr + s * STAGE + A_TILE; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < TR_STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
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
    if (a.mode == 0) {
        n_tile = blockIdx.x % a.n_tiles;
        const int m = blockIdx.x / a.n_tiles;
        px0 = (m % a.tiles_x) * 16; py0 = ((m / a.tiles_x) % a.tiles_y) * 8; pn = m / (a.tiles_x * a.tiles_y);
        steps = a.steps;
    } else {
        int u = blockIdx.x;
        const int split = u % a.splitk; u /= a.splitk;
        tap = u % 9; u /= 9;
        col0 = (u % a.col_tiles) * BN; row0 = (u / a.col_tiles) * 128;
        const int per = (a.ksteps + a.splitk - 1) / a.splitk;
        k0 = split * per;
        steps = min(per, a.ksteps - k0);
        if (steps < 0) steps = 0;
    }
    const int cin_chunks = a.Cin / 64;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < steps; i++) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], STAGE);
                if (a.mode == 0) {
                    const int t = i / cin_chunks, cc = i - t * cin_chunks;
                    tma_load_4d(sA(s), &a.mapA, &full_bar[s], cc * 64, px0 + t % 3 - 1, py0 + t / 3 - 1, pn);
                    tma_load_2d(sB(s), &a.mapB, &full_bar[s], i * 64, n_tile * BN);
                } else {
                    const int j = k0 + i;
                    const int xc = j % a.rc, y = (j / a.rc) % a.H, n = j / (a.rc * a.H);
                    const int dxi = tap % 3, dy = tap / 3 - 1;       // dxi selects the x-shifted copy of X (5th dimension)
                    if (a.a_shift == 0) {
                        tma_load_5d(sA(s), &a.mapA, &full_bar[s], xc * 64, y, row0, n, 0);
                        tma_load_5d(sB(s), &a.mapB, &full_bar[s], xc * 64, y + dy, col0, n, dxi);
                    } else {
                        tma_load_5d(sA(s), &a.mapA, &full_bar[s], xc * 64, y + dy, row0, n, dxi);
                        tma_load_5d(sB(s), &a.mapB, &full_bar[s], xc * 64, y, col0, n, 0);
                    }
                }
                if (++s == TR_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < steps; i++) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t ab = smem_u32(sA(s)), bb = smem_u32(sB(s));
#pragma unroll
                for (int k = 0; k < 4; k++)
                    umma_f16(tmem, umma_desc_sw128(ab + k * 32), umma_desc_sw128(bb + k * 32), idesc, (i > 0 || k > 0) ? 1u : 0u);
                umma_commit(&empty_bar[s]);
                if (i == steps - 1) umma_commit(&acc_bar);
            }
            __syncwarp();
            if (++s == TR_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (steps > 0) {
        const int q = warp & 3, row = q * 32 + lane;
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
        if (a.mode == 0) {
            const int x = px0 + row % 16, y = py0 + row / 16;
            const bool valid = x < a.W && y < a.H && pn < a.N;
            __nv_bfloat16* o = a.out + (((size_t)pn * a.H + y) * a.W + x) * a.out_pitch + n_tile * BN;
#pragma unroll 1
            for (int j = 0; j < BN / 32; j++) {
                uint32_t r[32];
                tmem_ld32(taddr + j * 32, r);
                tmem_ld_wait();
                if (valid) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        float v0 = __uint_as_float(r[2 * i]), v1 = __uint_as_float(r[2 * i + 1]);
                        if (a.bias) { v0 += __ldg(a.bias + n_tile * BN + j * 32 + 2 * i); v1 += __ldg(a.bias + n_tile * BN + j * 32 + 2 * i + 1); }
                        if (a.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                        pk[i] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(o + j * 32);
#pragma unroll
                    for (int i = 0; i < 4; i++) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                }
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < BN / 32; j++) {
                uint32_t r[32];
                tmem_ld32(taddr + j * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const int rr = row0 + row, cc = col0 + j * 32 + i;
                    const int co = a.a_shift == 0 ? rr : cc, ci = a.a_shift == 0 ? cc : rr;
                    if (co < a.Cout && ci < a.Cin) atomicAdd(a.dW + ((size_t)co * 9 + tap) * a.Cin + ci, __uint_as_float(r[i]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TCOLS);
}

// [X(x-1) | X | X(x+1)] with zero fill at the row ends, NCHW bf16: ws (3, N*C*H*W)
__global__ void shift3_kernel(const __nv_bfloat16* __restrict__ x, size_t total, int W, __nv_bfloat16* __restrict__ ws) {
    const __nv_bfloat16 zero = __float2bfloat16(0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % W);
        ws[i] = xx > 0 ? x[i - 1] : zero;
        ws[total + i] = x[i];
        ws[2 * total + i] = xx < W - 1 ? x[i + 1] : zero;
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

template <int BN>
static int launch_train(const TrainArgs& t, int grid, cudaStream_t st) {
    constexpr int SMEM = TR_STAGES * (128 * 128 + BN * 128) + 1024;
    static int have[64] = {};
    int dev = 0;
    IPER_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && have[dev] < SMEM) {
        IPER_CHECK_CUDA(cudaFuncSetAttribute(train_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        have[dev] = SMEM;
    }
    train_gemm_kernel<BN><<<grid, TR_THREADS, SMEM, st>>>(t);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace iper

using namespace iper;

extern "C" int iper_conv3x3_bf16(const void* x_nhwc, int N, int H, int W, int Cin, const void* w_packed, int Cout,
                                 const float* bias, int relu, void* out_nhwc, iper_stream_t stream) {
    IPER_REQUIRE(x_nhwc && w_packed && out_nhwc, "iper_conv3x3_bf16: null pointer");
    IPER_REQUIRE(N > 0 && H >= 8 && W >= 16 && Cin % 64 == 0 && Cout % 64 == 0, "iper_conv3x3_bf16: needs H >= 8, W >= 16, Cin %% 64 == 0, Cout %% 64 == 0 "
                 "(got %dx%d, %d -> %d)", H, W, Cin, Cout);
    IPER_REQUIRE(((uintptr_t)x_nhwc & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)out_nhwc & 15) == 0, "iper_conv3x3_bf16: 16-byte alignment");
    TrainArgs t = {};
    t.mode = 0; t.N = N; t.H = H; t.W = W; t.Cin = Cin; t.Cout = Cout;
    t.tiles_x = (W + 15) / 16; t.tiles_y = (H + 7) / 8;
    const int BN = Cout % 128 == 0 ? 128 : 64;
    t.n_tiles = Cout / BN; t.steps = 9 * (Cin / 64);
    t.bias = bias; t.relu = relu; t.out = reinterpret_cast<__nv_bfloat16*>(out_nhwc); t.out_pitch = Cout;
    cuuint64_t ad[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t as[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t ab[4] = {64, 16, 8, 1};
    if (int rc = bf16_map(&t.mapA, x_nhwc, 4, ad, as, ab)) return rc;
    cuuint64_t bd[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
    cuuint64_t bs[1] = {(cuuint64_t)9 * Cin * 2};
    cuuint32_t bb[2] = {64, (cuuint32_t)BN};
    if (int rc = bf16_map(&t.mapB, w_packed, 2, bd, bs, bb)) return rc;
    const int grid = t.tiles_x * t.tiles_y * N * t.n_tiles;
    return BN == 128 ? launch_train<128>(t, grid, (cudaStream_t)stream) : launch_train<64>(t, grid, (cudaStream_t)stream);
}

extern "C" size_t iper_conv3x3_wgrad_workspace_bytes(int N, int H, int W, int Cin) {
    if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0) return 0;
    return (size_t)3 * N * Cin * H * W * sizeof(__nv_bfloat16);
}

extern "C" int iper_conv3x3_wgrad_bf16(const void* x_nchw, const void* dy_nchw, int N, int H, int W, int Cin, int Cout,
                                       float* dW, void* workspace, size_t workspace_bytes, iper_stream_t stream) {
    IPER_REQUIRE(x_nchw && dy_nchw && dW && workspace, "iper_conv3x3_wgrad_bf16: null pointer");
    IPER_REQUIRE(workspace_bytes >= iper_conv3x3_wgrad_workspace_bytes(N, H, W, Cin) && ((uintptr_t)workspace & 15) == 0,
                 "iper_conv3x3_wgrad_bf16: workspace of %zu bytes (16-byte aligned) needed", iper_conv3x3_wgrad_workspace_bytes(N, H, W, Cin));
    IPER_REQUIRE(N > 0 && H > 0 && W % 64 == 0 && Cin % 64 == 0 && Cout % 64 == 0,
                 "iper_conv3x3_wgrad_bf16: needs W %% 64 == 0, Cin %% 64 == 0, Cout %% 64 == 0 (got %dx%d, %d -> %d)", H, W, Cin, Cout);
    IPER_REQUIRE(Cin % 128 == 0 || Cout % 128 == 0, "iper_conv3x3_wgrad_bf16: one of Cin, Cout must be a multiple of 128");
    cudaStream_t st = (cudaStream_t)stream;
    IPER_CHECK_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)Cout * 9 * Cin, st));
    const size_t xtotal = (size_t)N * Cin * H * W;
    shift3_kernel<<<(unsigned)std::min<size_t>((xtotal + 255) / 256, 148 * 16), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x_nchw), xtotal, W, reinterpret_cast<__nv_bfloat16*>(workspace));
    IPER_CHECK_CUDA(cudaGetLastError());
    TrainArgs t = {};
    t.mode = 1; t.N = N; t.H = H; t.W = W; t.Cin = Cin; t.Cout = Cout; t.dW = dW;
    t.rc = W / 64; t.ksteps = N * H * t.rc;
    t.a_shift = (Cout % 128 == 0) ? 0 : 1;                 // the 128-row operand is the tensor whose channel count allows it
    const int rowsC = t.a_shift == 0 ? Cout : Cin, colsC = t.a_shift == 0 ? Cin : Cout;
    const int BN = colsC % 128 == 0 ? 128 : 64;
    t.row_tiles = rowsC / 128; t.col_tiles = colsC / BN;
    const int items = t.row_tiles * t.col_tiles * 9;
    int sk = (148 * 2 + items - 1) / items;                // fill the machine about twice over
    if (sk > t.ksteps) sk = t.ksteps;
    if (sk < 1) sk = 1;
    t.splitk = sk;
    // 5-D maps (W, H, C, N, copy): X lives in the workspace as three x-shifted copies, dY has a single "copy"
    const void* rows_t = t.a_shift == 0 ? dy_nchw : (const void*)workspace; const void* cols_t = t.a_shift == 0 ? (const void*)workspace : dy_nchw;
    const cuuint64_t rcopies = t.a_shift == 0 ? 1 : 3, ccopies = t.a_shift == 0 ? 3 : 1;
    cuuint64_t rd[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)rowsC, (cuuint64_t)N, rcopies};
    cuuint64_t rs[4] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)rowsC * H * W * 2, (cuuint64_t)N * rowsC * H * W * 2};
    cuuint32_t rb[5] = {64, 1, 128, 1, 1};
    if (int rc = bf16_map(&t.mapA, rows_t, 5, rd, rs, rb)) return rc;
    cuuint64_t cd[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)colsC, (cuuint64_t)N, ccopies};
    cuuint64_t cs[4] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)colsC * H * W * 2, (cuuint64_t)N * colsC * H * W * 2};
    cuuint32_t cb[5] = {64, 1, (cuuint32_t)BN, 1, 1};
    if (int rc = bf16_map(&t.mapB, cols_t, 5, cd, cs, cb)) return rc;
    const int grid = items * sk;
    return BN == 128 ? launch_train<128>(t, grid, st) : launch_train<64>(t, grid, st);
}
