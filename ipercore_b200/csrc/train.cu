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
    int N, H, W;                         // the TILE DOMAIN: forward -> output pixels (or, for a phase of a transposed convolution, the
                                         // input-resolution grid); wgrad -> the pixels of dY
    int Cin, Cout;                       // forward: contraction channels and output channels
    int stride;                          // 1 | 2: the read tensor (forward: x, wgrad: X) is sampled at stride * p + tap offset; stride 2
                                         // reads through a 5-D view (2C [x parity, c], W/2, 2 [y parity], H/2, N) of the NHWC tensor
    int ntaps; int8_t tdx[49], tdy[49];  // tap offsets
    int tiles_x, tiles_y, n_tiles;
    int steps;                           // forward: ntaps * Cin/64
    const float* bias; int relu;
    const __nv_bfloat16* add;            // residual (same addressing as out) or null
    __nv_bfloat16* out; int out_pitch;
    int Hout, Wout, osy, osx, ooy, oox;  // output pixel of tile-domain pixel (y, x) = (osy*y + ooy, osx*x + oox) in (N, Hout, Wout, out_pitch)
    // wgrad: mapA = the tensor whose channels are the 128 accumulator rows, mapB = the BN accumulator columns
    int x_rows;                          // 0: rows = dY channels (co), cols = X channels (ci);  1: rows = ci, cols = co
    int xC;                              // channel count of X (parity offset of the 5-D view)
    int rowsC, colsC, col_tiles, splitk, kblocks;
    float* dW; long long s_row, s_col, s_tap;   // element strides of dW for (row channel, column channel, tap)
};

// activation box at tile origin (x0, y0) of image n, channel chunk c0, shifted by tap offset (dx, dy); stride 2: through the 5-D view
IPER_DEVINL void load_act(void* smem, const CUtensorMap* map, uint64_t* bar, int stride, int C, int c0, int x0, int y0, int n, int dx, int dy) {
    if (stride == 1) {
        tma_load_4d(smem, map, bar, c0, x0 + dx, y0 + dy, n);
    } else {
        const int px = dx & 1, py = dy & 1;                 // 2*x0 + dx = 2*(x0 + (dx - px)/2) + px
        tma_load_5d(smem, map, bar, px * C + c0, x0 + ((dx - px) >> 1), py, y0 + ((dy - py) >> 1), n);
    }
}

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
        tap = u % a.ntaps; u /= a.ntaps;
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
                    load_act(sA(s), &a.mapA, &full_bar[s], a.stride, a.Cin, cc * 64, px0, py0, pn, a.tdx[t], a.tdy[t]);
                    tma_load_2d(sB(s), &a.mapB, &full_bar[s], i * 64, n_tile * BN);
                } else {
                    const int j = k0 + i;
                    const int x0 = (j % a.tiles_x) * 16, y0 = ((j / a.tiles_x) % a.tiles_y) * 8, n = j / (a.tiles_x * a.tiles_y);
                    const int dx = a.tdx[tap], dy = a.tdy[tap];                      // X is read at stride * (dY pixel) + tap offset
#pragma unroll
                    for (int g = 0; g < 2; g++) {      // channels beyond the tensor are zero-filled by TMA (64-channel tensors)
                        if (a.x_rows) load_act(sA(s) + g * TR_BOX, &a.mapA, &full_bar[s], a.stride, a.xC, row0 + g * 64, x0, y0, n, dx, dy);
                        else tma_load_4d(sA(s) + g * TR_BOX, &a.mapA, &full_bar[s], row0 + g * 64, x0, y0, n);
                    }
#pragma unroll
                    for (int g = 0; g < BN / 64; g++) {
                        if (a.x_rows) tma_load_4d(sB(s) + g * TR_BOX, &a.mapB, &full_bar[s], col0 + g * 64, x0, y0, n);
                        else load_act(sB(s) + g * TR_BOX, &a.mapB, &full_bar[s], a.stride, a.xC, col0 + g * 64, x0, y0, n, dx, dy);
                    }
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
            const size_t off = (((size_t)pn * a.Hout + (a.osy * y + a.ooy)) * a.Wout + (a.osx * x + a.oox)) * a.out_pitch + n_tile * BN;
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
        if (sg.taps > 0) {
            // repacking of the (co, ci, ky, kx) tensor (a transposed convolution's weight is (ci_T, co_T, ky, kx): co = ci_T, ci = co_T):
            //   plain  (co_pad rows, K = (tap, ci))                 forward of kinds 1, 2; data gradient of kind 3
            //   rev    (ci_pad rows, K = (taps-1-tap, co))          data gradient of kind 1 (rotated by 180 degrees, in/out transposed)
            //   phases (ci_pad rows, K = (tap in phase, co)) x 4    data gradient of kind 2, forward of kind 3 (stride-2 transposition)
            const int kind = sg.reserved & 0xff, ks = (sg.reserved >> 8) & 0xff, pad = (sg.reserved >> 16) & 0xff;
            const __nv_bfloat16 h = __float2bfloat16(pp);
            const long long plain = ((long long)co * sg.taps + tap) * sg.ci_pad + ci;
            if (kind == 1) {
                pack_fwd[sg.fwd_offset + plain] = h;
                if (sg.dgrad_offset >= 0)
                    pack_dgrad[sg.dgrad_offset + ((long long)ci * sg.taps + (sg.taps - 1 - tap)) * sg.co_pad + co] = h;
            } else {
                const int ky = tap / ks, kx = tap % ks, py = (ky + pad) & 1, px = (kx + pad) & 1;
                int n0 = 0;                                   // taps per axis whose phase is 0
                for (int k = 0; k < ks; k++) n0 += ((k + pad) & 1) == 0;
                const int n1 = ks - n0, nyp = py ? n1 : n0, nxp = px ? n1 : n0;
                long long off = 0;                            // blocks of the phases before (py, px) in the order (0,0) (0,1) (1,0) (1,1)
                for (int ph = 0; ph < py * 2 + px; ph++) off += (long long)sg.ci_pad * ((ph >> 1) ? n1 : n0) * ((ph & 1) ? n1 : n0) * sg.co_pad;
                const long long pidx = off + ((long long)ci * (nyp * nxp) + (ky >> 1) * nxp + (kx >> 1)) * sg.co_pad + co;
                if (kind == 2) { pack_fwd[sg.fwd_offset + plain] = h; pack_dgrad[sg.dgrad_offset + pidx] = h; }
                else { pack_fwd[sg.fwd_offset + pidx] = h; pack_dgrad[sg.dgrad_offset + plain] = h; }
            }
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

// the read tensor of a convolution: plain NHWC (stride 1) or the 5-D parity view of it (stride 2; H, W even)
static int read_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int stride) {
    if (stride == 1) return nhwc_map(m, base, N, H, W, C);
    cuuint64_t d[5] = {(cuuint64_t)2 * C, (cuuint64_t)W / 2, 2, (cuuint64_t)H / 2, (cuuint64_t)N};
    cuuint64_t st[4] = {(cuuint64_t)2 * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)2 * W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t b[5] = {64, 16, 1, 8, 1};
    return bf16_map(m, base, 5, d, st, b);
}
static int weight_map(CUtensorMap* m, const void* w, long long K, int rows, int BN) {
    cuuint64_t bd[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t bs[1] = {(cuuint64_t)K * 2};
    cuuint32_t bb[2] = {64, (cuuint32_t)BN};
    return bf16_map(m, w, 2, bd, bs, bb);
}
static int launch_fwd(TrainArgs& t, cudaStream_t st) {
    t.tiles_x = (t.W + 15) / 16; t.tiles_y = (t.H + 7) / 8;
    const int BN = t.Cout % 128 == 0 ? 128 : 64;
    t.n_tiles = t.Cout / BN; t.steps = t.ntaps * (t.Cin / 64);
    const int grid = t.tiles_x * t.tiles_y * t.N * t.n_tiles;
    return BN == 128 ? launch_train<128, 0>(t, grid, st) : launch_train<64, 0>(t, grid, st);
}

}  // namespace iper

using namespace iper;

#define IPER_A16P(p) (((uintptr_t)(p) & 15) == 0)

extern "C" int iper_conv_bf16(const void* x_nhwc, int N, int H, int W, int Cin, const void* w_packed, int Cout, int ksize, int stride,
                              int pad, const float* bias, int relu, const void* add_nhwc, void* out_nhwc, iper_stream_t stream) {
    IPER_REQUIRE(x_nhwc && w_packed && out_nhwc, "iper_conv_bf16: null pointer");
    IPER_REQUIRE(ksize >= 1 && ksize <= 7 && (stride == 1 || stride == 2) && pad >= 0 && pad < ksize,
                 "iper_conv_bf16: kernel %d stride %d pad %d unsupported (k <= 7, stride 1 | 2, pad < k)", ksize, stride, pad);
    IPER_REQUIRE(N > 0 && Cin % 64 == 0 && Cout % 64 == 0 && (stride == 1 || (H % 2 == 0 && W % 2 == 0)),
                 "iper_conv_bf16: needs Cin %% 64 == 0, Cout %% 64 == 0, even H, W for stride 2 (got %dx%d, %d -> %d)", H, W, Cin, Cout);
    const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    IPER_REQUIRE(Ho >= 1 && Wo >= 1 && H / stride >= 8 && W / stride >= 16, "iper_conv_bf16: map %dx%d too small (one 16x8 TMA box per tap)", H, W);
    IPER_REQUIRE(IPER_A16P(x_nhwc) && IPER_A16P(w_packed) && IPER_A16P(out_nhwc) && IPER_A16P(add_nhwc), "iper_conv_bf16: 16-byte alignment");
    TrainArgs t = {};
    t.N = N; t.H = Ho; t.W = Wo; t.Cin = Cin; t.Cout = Cout; t.stride = stride; t.ntaps = ksize * ksize;
    for (int k = 0; k < t.ntaps; k++) { t.tdx[k] = (int8_t)(k % ksize - pad); t.tdy[k] = (int8_t)(k / ksize - pad); }
    t.bias = bias; t.relu = relu; t.add = reinterpret_cast<const __nv_bfloat16*>(add_nhwc);
    t.out = reinterpret_cast<__nv_bfloat16*>(out_nhwc); t.out_pitch = Cout;
    t.Hout = Ho; t.Wout = Wo; t.osx = t.osy = 1;
    if (int rc = read_map(&t.mapA, x_nhwc, N, H, W, Cin, stride)) return rc;
    if (int rc = weight_map(&t.mapB, w_packed, (long long)t.ntaps * Cin, Cout, Cout % 128 == 0 ? 128 : 64)) return rc;
    return launch_fwd(t, (cudaStream_t)stream);
}

extern "C" int iper_conv_transposed_bf16(const void* x_nhwc, int N, int H, int W, int Cin, const void* w_phases, int Cout, int ksize,
                                         int pad, const float* bias, int relu, void* out_nhwc, iper_stream_t stream) {
    IPER_REQUIRE(x_nhwc && w_phases && out_nhwc, "iper_conv_transposed_bf16: null pointer");
    IPER_REQUIRE((ksize == 3 || ksize == 4) && pad == 1, "iper_conv_transposed_bf16: kernel %d pad %d unsupported (3 or 4, pad 1, stride 2: output 2H x 2W)", ksize, pad);
    IPER_REQUIRE(N > 0 && H >= 8 && W >= 16 && Cin % 64 == 0 && Cout % 64 == 0,
                 "iper_conv_transposed_bf16: needs H >= 8, W >= 16, Cin %% 64 == 0, Cout %% 64 == 0 (got %dx%d, %d -> %d)", H, W, Cin, Cout);
    IPER_REQUIRE(IPER_A16P(x_nhwc) && IPER_A16P(w_phases) && IPER_A16P(out_nhwc), "iper_conv_transposed_bf16: 16-byte alignment");
    // out[2y + py, 2x + px] = sum over taps with (ky + pad + py) even of x[y + (py + pad - ky)/2, x + (px + pad - kx)/2] w[ky, kx]:
    // four stride-1 phase convolutions over the input grid, each with its own K-major weight block (rows Cout, K = (tap in phase, ci))
    long long woff = 0;
    for (int ph = 0; ph < 4; ph++) {
        const int py = ph >> 1, px = ph & 1;
        TrainArgs t = {};
        t.N = N; t.H = H; t.W = W; t.Cin = Cin; t.Cout = Cout; t.stride = 1;
        for (int ky = 0; ky < ksize; ky++) {
            if ((ky + pad + py) & 1) continue;
            for (int kx = 0; kx < ksize; kx++) {
                if ((kx + pad + px) & 1) continue;
                t.tdx[t.ntaps] = (int8_t)((px + pad - kx) / 2); t.tdy[t.ntaps] = (int8_t)((py + pad - ky) / 2); t.ntaps++;
            }
        }
        t.bias = bias; t.relu = relu; t.out = reinterpret_cast<__nv_bfloat16*>(out_nhwc); t.out_pitch = Cout;
        t.Hout = 2 * H; t.Wout = 2 * W; t.osx = t.osy = 2; t.oox = px; t.ooy = py;
        if (int rc = nhwc_map(&t.mapA, x_nhwc, N, H, W, Cin)) return rc;
        if (int rc = weight_map(&t.mapB, reinterpret_cast<const __nv_bfloat16*>(w_phases) + woff, (long long)t.ntaps * Cin, Cout,
                                Cout % 128 == 0 ? 128 : 64)) return rc;
        if (int rc = launch_fwd(t, (cudaStream_t)stream)) return rc;
        woff += (long long)Cout * t.ntaps * Cin;
    }
    return 0;
}

extern "C" int iper_conv_wgrad_bf16(const void* x_nhwc, const void* dy_nhwc, int N, int H, int W, int Cin, int Cout, int ksize, int stride,
                                    int pad, float* dW, long long stride_co, long long stride_ci, long long stride_tap, int co_valid,
                                    int ci_valid, iper_stream_t stream) {
    IPER_REQUIRE(x_nhwc && dy_nhwc && dW, "iper_conv_wgrad_bf16: null pointer");
    IPER_REQUIRE(ksize >= 1 && ksize <= 7 && (stride == 1 || stride == 2) && pad >= 0 && pad < ksize,
                 "iper_conv_wgrad_bf16: kernel %d stride %d pad %d unsupported", ksize, stride, pad);
    IPER_REQUIRE(N > 0 && H > 0 && W > 0 && Cin % 64 == 0 && Cout % 64 == 0 && (stride == 1 || (H % 2 == 0 && W % 2 == 0)),
                 "iper_conv_wgrad_bf16: needs Cin %% 64 == 0, Cout %% 64 == 0, even H, W for stride 2 (got %dx%d, %d -> %d)", H, W, Cin, Cout);
    IPER_REQUIRE(IPER_A16P(x_nhwc) && IPER_A16P(dy_nhwc), "iper_conv_wgrad_bf16: 16-byte alignment");
    IPER_REQUIRE(co_valid > 0 && co_valid <= Cout && ci_valid > 0 && ci_valid <= Cin, "iper_conv_wgrad_bf16: valid channel counts out of range");
    const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;      // dY is (N, Ho, Wo, Cout)
    IPER_REQUIRE(Ho >= 1 && Wo >= 1 && H / stride >= 8 && W / stride >= 16, "iper_conv_wgrad_bf16: map %dx%d too small (one 16x8 TMA box per tap)", H, W);
    cudaStream_t st = (cudaStream_t)stream;
    TrainArgs t = {};
    t.N = N; t.H = Ho; t.W = Wo; t.Cin = Cin; t.Cout = Cout; t.stride = stride; t.ntaps = ksize * ksize; t.dW = dW; t.xC = Cin;
    for (int k = 0; k < t.ntaps; k++) { t.tdx[k] = (int8_t)(k % ksize - pad); t.tdy[k] = (int8_t)(k / ksize - pad); }
    t.tiles_x = (Wo + 15) / 16; t.tiles_y = (Ho + 7) / 8;
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
    const int items = row_tiles * t.col_tiles * t.ntaps;
    int sk = (148 * 2 + items - 1) / items;                // fill the machine about twice over
    if (sk > t.kblocks) sk = t.kblocks;
    if (sk < 1) sk = 1;
    t.splitk = sk;
    CUtensorMap mx, my;
    if (int rc = read_map(&mx, x_nhwc, N, H, W, Cin, stride)) return rc;
    if (int rc = nhwc_map(&my, dy_nhwc, N, Ho, Wo, Cout)) return rc;
    t.mapA = t.x_rows ? mx : my; t.mapB = t.x_rows ? my : mx;
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
