// HBM-bound kernels of the generator path (sm_100a): stem conv, instance-norm statistics, flow-guided warp +
// per-pixel source attention, layout converters, and a CUDA-core direct convolution used as the on-device
// cross-check of the tcgen05 path.  Reference line citations are in include/iper_b200.h.
#include <cstdlib>

#include "common.cuh"
#include "iper_b200.h"

namespace iper {

// ------------------------------------------------------------------------------------------------------------
// CUDA-core direct convolution on NHWC planes; one thread per (pixel, output channel), fp32 accumulate in the
// reference's own weight layout.  Slow by design: it is the checker for conv_tc.cu and the fallback for shapes
// the tensor-core kernel does not take.
// ------------------------------------------------------------------------------------------------------------
struct DirectArgs {
    const __half* a; int a_planes; long long a_plane_stride; int N, H, W, a_pitch, a_coff, Cin;
    int mode, ksize, Cout, oH, oW;
    const float* w; const float* bias; int relu, epi;
    void* out; int out_planes; long long out_plane_stride; int out_pitch, out_coff;
    const __half* x; int x_planes; long long x_plane_stride; int x_pitch, x_coff;
    const float* mean_rstd; int spade_C;
};

IPER_DEVINL float direct_dot(const DirectArgs& d, int n, int oy, int ox, int co) {
    float acc = 0.f;
    if (d.mode == IPER_CONVT_4S2) {
        // out[oy] += in[iy] * W[ci][co][ky][kx] with oy = 2*iy - 1 + ky
        for (int ky = 0; ky < 4; ky++) {
            const int ty = oy + 1 - ky;
            if (ty < 0 || (ty & 1)) continue;
            const int iy = ty >> 1;
            if (iy >= d.H) continue;
            for (int kx = 0; kx < 4; kx++) {
                const int tx = ox + 1 - kx;
                if (tx < 0 || (tx & 1)) continue;
                const int ix = tx >> 1;
                if (ix >= d.W) continue;
                const size_t base = (((size_t)n * d.H + iy) * d.W + ix) * d.a_pitch + d.a_coff;
                for (int ci = 0; ci < d.Cin; ci++)
                    acc += load_plane_val(d.a, d.a_planes, d.a_plane_stride, base + ci) *
                           d.w[(((size_t)ci * d.Cout + co) * 4 + ky) * 4 + kx];
            }
        }
    } else {
        const int stride = d.mode == IPER_CONV_S2 ? 2 : 1;
        const int k = d.mode == IPER_CONV_S2 ? 3 : d.ksize, pad = k / 2;
        for (int ky = 0; ky < k; ky++) {
            const int iy = oy * stride - pad + ky;
            if (iy < 0 || iy >= d.H) continue;
            for (int kx = 0; kx < k; kx++) {
                const int ix = ox * stride - pad + kx;
                if (ix < 0 || ix >= d.W) continue;
                const size_t base = (((size_t)n * d.H + iy) * d.W + ix) * d.a_pitch + d.a_coff;
                for (int ci = 0; ci < d.Cin; ci++)
                    acc += load_plane_val(d.a, d.a_planes, d.a_plane_stride, base + ci) *
                           d.w[(((size_t)co * d.Cin + ci) * k + ky) * k + kx];
            }
        }
    }
    return acc;
}

__global__ void conv_direct_kernel(const DirectArgs d) {
    const int Cthreads = d.epi == IPER_EPI_SPADE ? d.spade_C : d.Cout;
    const size_t total = (size_t)d.N * d.oH * d.oW * Cthreads;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cthreads);
        const size_t pix = i / Cthreads;
        const int ox = (int)(pix % d.oW), oy = (int)((pix / d.oW) % d.oH), n = (int)(pix / ((size_t)d.oW * d.oH));
        float v;
        if (d.epi == IPER_EPI_SPADE) {
            // weights (2C, Cin, k, k): rows [0,C) gamma, [C,2C) beta (reference order: mlp_gamma then mlp_beta)
            const float gamma = direct_dot(d, n, oy, ox, c) + d.bias[c];
            const float beta = direct_dot(d, n, oy, ox, d.spade_C + c) + d.bias[d.spade_C + c];
            const float xv = load_plane_val(d.x, d.x_planes, d.x_plane_stride, pix * d.x_pitch + d.x_coff + c);
            const float* mr = d.mean_rstd + ((size_t)n * d.spade_C + c) * 2;
            v = (xv - mr[0]) * mr[1] * (1.f + gamma) + beta;
        } else {
            v = direct_dot(d, n, oy, ox, c);
            if (d.bias) v += d.bias[c];
            if (d.x) v = load_plane_val(d.x, d.x_planes, d.x_plane_stride, pix * d.x_pitch + d.x_coff + c) + v;
            if (d.relu) v = fmaxf(v, 0.f);
        }
        const size_t o = pix * d.out_pitch + d.out_coff + c;
        if (d.epi == IPER_EPI_F32) reinterpret_cast<float*>(d.out)[o] = v;
        else store_plane_val(reinterpret_cast<__half*>(d.out), d.out_planes, d.out_plane_stride, o, v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Stem: Conv2d(Cin<=8 -> 64, 3x3, s2, p1)(+bias)+ReLU, NCHW fp32 image -> NHWC planes.  K = 9*Cin <= 72 is
// tensor-core hostile; 0.45 GFLOP per 512^2 frame.  Block = 32x8 output pixels; the 65x17 input halo and the whole
// filter bank live in shared memory; each thread computes 1 pixel x 8 channels per pass.  (A 2-pixels-per-thread
// variant that halves the shared-memory filter reads per FMA was measured slower: 1.44 vs 1.34 ms per 50 frames.)
// ------------------------------------------------------------------------------------------------------------
constexpr int STEM_TW = 32, STEM_TH = 8, STEM_MAXC = 8;
__host__ __device__ constexpr int stem_halo_floats(int cin) { return ((cin * (2 * STEM_TH + 1) * (2 * STEM_TW + 1) + 3) / 4) * 4; }

template <int CIN>
__global__ void __launch_bounds__(256) conv_stem_kernel(const float* __restrict__ in, int N, int H, int W,
                                                        const float* __restrict__ wgt, const float* __restrict__ bias,
                                                        int Cout, __half* __restrict__ out, int out_planes,
                                                        long long out_plane_stride, int out_pitch, int out_coff,
                                                        double* __restrict__ stats_ws) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float s_stat[2 * 128];
    constexpr int HALO_W = 2 * STEM_TW + 1, HALO_H = 2 * STEM_TH + 1;
    float* s_in = sm;                                  // [CIN][HALO_H][HALO_W]
    float* s_w = sm + stem_halo_floats(CIN);           // [9*CIN][Cout]  (ci,ky,kx major, cout fastest)
    const int Ho = H / 2, Wo = W / 2;
    const int n = blockIdx.z, oy0 = blockIdx.y * STEM_TH, ox0 = blockIdx.x * STEM_TW;
    const int tid = threadIdx.x;
    for (int i = tid; i < Cout * CIN * 9; i += 256) {
        const int co = i / (CIN * 9), r = i % (CIN * 9);        // source order (co, ci, ky, kx)
        s_w[r * Cout + co] = wgt[i];
    }
    const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
    for (int i = tid; i < CIN * HALO_H * HALO_W; i += 256) {
        const int c = i / (HALO_H * HALO_W), r = i % (HALO_H * HALO_W);
        const int iy = iy0 + r / HALO_W, ix = ix0 + r % HALO_W;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = in[(((size_t)n * CIN + c) * H + iy) * W + ix];
        s_in[i] = v;
    }
    __syncthreads();
    const int lx = tid % STEM_TW, ly = tid / STEM_TW;   // 32 x 8 output pixels per block
    const int oy = oy0 + ly, ox = ox0 + lx;
    float patch[CIN * 9];
#pragma unroll
    for (int c = 0; c < CIN; c++)
#pragma unroll
        for (int t = 0; t < 9; t++)
            patch[c * 9 + t] = s_in[(c * HALO_H + 2 * ly + t / 3) * HALO_W + 2 * lx + t % 3];
    const bool inside = (oy < Ho) && (ox < Wo);
    if (stats_ws) {
        for (int i = tid; i < 2 * Cout; i += 256) s_stat[i] = 0.f;
        __syncthreads();
    } else if (!inside) {
        return;
    }
    const size_t obase = (((size_t)n * Ho + oy) * Wo + ox) * out_pitch + out_coff;
    for (int co0 = 0; co0 < Cout; co0 += 8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
        for (int r = 0; r < CIN * 9; r++) {
            const float pv = patch[r];
            const float4 w0 = *reinterpret_cast<const float4*>(&s_w[r * Cout + co0]);
            const float4 w1 = *reinterpret_cast<const float4*>(&s_w[r * Cout + co0 + 4]);
            acc[0] += pv * w0.x; acc[1] += pv * w0.y; acc[2] += pv * w0.z; acc[3] += pv * w0.w;
            acc[4] += pv * w1.x; acc[5] += pv * w1.y; acc[6] += pv * w1.z; acc[7] += pv * w1.w;
        }
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; j++) o8[j] = fmaxf(acc[j] + (bias ? bias[co0 + j] : 0.f), 0.f);
        if (inside) store_planes8(out, out_planes, out_plane_stride, obase + co0, o8);
        if (stats_ws) {
            // fused instance-norm statistics: 8 channels x 32 pixels per warp -> exchange-and-halve (8 -> 4 -> 2 -> 1
            // values, 7 shuffles) then two plain xor steps; lane l < 8 ends with the warp's sum of channel co0 + l
            float s1[8], s2[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { s1[j] = inside ? o8[j] : 0.f; s2[j] = s1[j] * s1[j]; }
            const int lane = tid & 31;
#pragma unroll
            for (int half = 4; half >= 1; half >>= 1) {
                const bool up = (lane & half) != 0;
#pragma unroll
                for (int j = 0; j < half; j++) {
                    const float k1 = up ? s1[j + half] : s1[j], d1 = up ? s1[j] : s1[j + half];
                    const float k2 = up ? s2[j + half] : s2[j], d2 = up ? s2[j] : s2[j + half];
                    s1[j] = k1 + __shfl_xor_sync(0xffffffffu, d1, half);
                    s2[j] = k2 + __shfl_xor_sync(0xffffffffu, d2, half);
                }
            }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
                s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], o);
                s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], o);
            }
            if (lane < 8) { atomicAdd(&s_stat[2 * (co0 + lane)], s1[0]); atomicAdd(&s_stat[2 * (co0 + lane) + 1], s2[0]); }
        }
    }
    if (stats_ws) {
        __syncthreads();
        for (int i = tid; i < 2 * Cout; i += 256) atomicAdd(&stats_ws[(size_t)n * 2 * Cout + i], (double)s_stat[i]);
    }
}

// Stem as a tensor-core GEMM: explicit im2col of the 3x3 / stride-2 / pad-1 patches of a (N, CIN <= 7, H, W) fp32 image into
// NHWC planes (N, H/2, W/2, 64) with channel k = (ky*3 + kx)*CIN + ci (zeros from 9*CIN up), so that the stem becomes a
// K = 64 "1x1 convolution" on the tcgen05 path (fused bias / ReLU / instance-norm statistics epilogue) instead of 54-term
// FMA chains fed by shared-memory filter reads.  8 lanes cover one output pixel (8 channels = one 128-bit store per plane).
template <int CIN>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ in, int N, int H, int W,
                                                          __half* __restrict__ out, int out_planes, long long out_plane_stride,
                                                          int out_pitch, int out_coff) {
    const int Ho = H / 2, Wo = W / 2;
    const size_t total = (size_t)N * Ho * Wo * 8;
    constexpr int kmax = 9 * CIN;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int cg = (int)(i & 7);
        const size_t pix = i >> 3;
        const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho);
        const size_t n = pix / ((size_t)Wo * Ho);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = cg * 8 + j;
            float val = 0.f;
            if (k < kmax) {
                const int tap = k / CIN, ci = k - tap * CIN;
                const int iy = 2 * oy - 1 + tap / 3, ix = 2 * ox - 1 + tap % 3;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = __ldg(in + ((n * CIN + ci) * H + iy) * W + ix);
            }
            v[j] = val;
        }
        store_planes8(out, out_planes, out_plane_stride, pix * out_pitch + out_coff + cg * 8, v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Instance-norm statistics over an NHWC planes tensor, two launches:
//   partial: grid (splits, N), block 256; a thread owns 8 consecutive channels (one 128-bit load per plane),
//            C/8 lanes cover a pixel, the rest of the block strides over the CTA's pixel chunk; fp32 partial sums
//            per thread are folded into fp64 shared accumulators and then into a global fp64 workspace (N,C,2);
//   final  : mean / rstd as fp32 (biased variance, eps inside the sqrt) — F.instance_norm semantics.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) instnorm_partial_kernel(const __half* __restrict__ x, int x_planes,
                                                               long long x_plane_stride, int HW, int C, int x_pitch,
                                                               int x_coff, int chunk, double* __restrict__ ws) {
    __shared__ double s_acc[2 * 256];
    const int n = blockIdx.y, tid = threadIdx.x;
    const int lpp = C >> 3;                       // lanes per pixel
    const int cg = tid % lpp, pl = tid / lpp;     // channel group, pixel lane
    const int ppl = 256 / lpp;                    // pixel lanes per block
    for (int i = tid; i < 2 * C; i += 256) s_acc[i] = 0.0;
    __syncthreads();
    float sum[8], sq[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { sum[j] = 0.f; sq[j] = 0.f; }
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, HW);
    const size_t base = (size_t)n * HW * x_pitch + x_coff + cg * 8;
    if (pl < ppl) {
        for (int p = p0 + pl; p < p1; p += ppl) {
            float v[8];
            load_planes8(x, x_planes, x_plane_stride, base + (size_t)p * x_pitch, v);
#pragma unroll
            for (int j = 0; j < 8; j++) { sum[j] += v[j]; sq[j] += v[j] * v[j]; }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            atomicAdd(&s_acc[2 * (cg * 8 + j)], (double)sum[j]);
            atomicAdd(&s_acc[2 * (cg * 8 + j) + 1], (double)sq[j]);
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * C; i += 256) atomicAdd(&ws[(size_t)n * 2 * C + i], s_acc[i]);
}

__global__ void instnorm_final_kernel(const double* __restrict__ ws, int total, int HW, float eps,
                                      float* __restrict__ mean_rstd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double mean = ws[2 * i] / HW;
    double var = ws[2 * i + 1] / HW - mean * mean;   // biased variance (F.instance_norm)
    if (var < 0.0) var = 0.0;
    mean_rstd[2 * i] = (float)mean;
    mean_rstd[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// InstanceNorm2d apply (+ReLU) (+residual) on NHWC planes — BGNet blocks (bg_inpaintor.py:13-21, 33-52)
__global__ void __launch_bounds__(256) instnorm_apply_kernel(const __half* __restrict__ x, int x_planes,
                                                             long long x_plane_stride, int x_pitch, int x_coff,
                                                             const float* __restrict__ mean_rstd, int N, int HW, int C,
                                                             int relu, const __half* __restrict__ res, int res_planes,
                                                             long long res_plane_stride, int res_pitch, int res_coff,
                                                             __half* __restrict__ out, int out_planes,
                                                             long long out_plane_stride, int out_pitch, int out_coff) {
    const size_t total = (size_t)N * HW * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t pix = i / C;
        const int n = (int)(pix / HW);
        const float* mr = mean_rstd + ((size_t)n * C + c) * 2;
        float v = (load_plane_val(x, x_planes, x_plane_stride, pix * x_pitch + x_coff + c) - mr[0]) * mr[1];
        if (relu) v = fmaxf(v, 0.f);
        if (res) v = load_plane_val(res, res_planes, res_plane_stride, pix * res_pitch + res_coff + c) + v;
        store_plane_val(out, out_planes, out_plane_stride, pix * out_pitch + out_coff + c, v);
    }
}

__global__ void tanh_nhwc_to_nchw_kernel(const float* __restrict__ in, int N, int HW, int C, int pitch,
                                         float* __restrict__ out) {
    const size_t total = (size_t)N * C * HW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i % HW, c = (i / HW) % C, n = i / ((size_t)HW * C);
        out[i] = tanhf(in[(n * HW + p) * pitch + c]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Flow-guided warp + per-pixel source attention.  One warp per target pixel; lanes stride the channel dimension
// with 128-bit (4 x fp32) gathers from the precomputed [Wk x | Wv x] source maps.  Bilinear taps / zero padding
// follow F.grid_sample(align_corners=False).  C <= 256, ns <= 8.
// ------------------------------------------------------------------------------------------------------------
constexpr int ATT_MAX_NS = 8;
struct Taps { int off[4]; float wt[4]; };
IPER_DEVINL Taps bilinear_taps(float gx, float gy, int h, int w) {
    Taps t;
    const float ix = ((gx + 1.f) * w - 1.f) * 0.5f, iy = ((gy + 1.f) * h - 1.f) * 0.5f;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    // huge/NaN coordinates: every tap out of range
    const bool sane = fabsf(ix) < 1e8f && fabsf(iy) < 1e8f;
    const int x0 = sane ? (int)fx0 : -10, y0 = sane ? (int)fy0 : -10, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix, wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
    const int xs[4] = {x0, x1, x0, x1}, ys[4] = {y0, y0, y1, y1};
    const float ws[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool ok = xs[i] >= 0 && xs[i] < w && ys[i] >= 0 && ys[i] < h;
        t.off[i] = ok ? ys[i] * w + xs[i] : -1;
        t.wt[i] = ok ? ws[i] : 0.f;
    }
    return t;
}

// Source-side maps per source pixel (fp32, pitch 2C+64): [ K'' = (Wq^T Wk) x_s | V' = Wv x_s | k0 = (Wk^T bq) . x_s | pad ].
// With q = Wq x_t + bq and K_s = warp(Wk x_s) + bk:  K_s . q = warp(K'')_s . x_t + warp(k0)_s + (bk . q), and the last
// term is the same for every source s, so it cancels in softmax_s — the per-frame q projection disappears.
// WIDE = 1: the 16 128-bit gathers of a source (4 corners x {K'' lo, K'' hi, V' lo, V' hi}) are issued before any of them
// is consumed (out-of-range corners read pixel 0 with weight 0), two CTAs per SM; WIDE = 0: corner by corner, three CTAs;
// WIDE = 2: corner by corner compiled for four CTAs per SM (64 registers, ~20 spilled values).
template <int C, int NSMAX, int WIDE>
__global__ void __launch_bounds__(256, WIDE == 1 ? 2 : (WIDE == 2 ? 4 : 3)) warp_attention_kernel(const __half* __restrict__ xt, int xt_planes,
                                                             long long xt_plane_stride, int xt_pitch, int xt_coff,
                                                             const float* __restrict__ kv,
                                                             const float* __restrict__ bias_v,
                                                             const float* __restrict__ T, int B, int ns, int h, int w,
                                                             __half* __restrict__ out, int out_planes,
                                                             long long out_plane_stride, int out_pitch, int out_coff) {
    // a thread owns 8 consecutive channels (two 128-bit loads per map); LPP lanes cover one pixel
    constexpr int LPP = C / 8, PPW = 32 / LPP, KVP = 2 * C + 64;
    const int lane = threadIdx.x & 31, cg = lane % LPP, sub = lane / LPP;
    const size_t hw = (size_t)h * w, total = (size_t)B * hw;
    const float inv_sqrt = 1.f / sqrtf((float)C);
    const float4* bvp = reinterpret_cast<const float4*>(bias_v + cg * 8);   // re-read per source (L1 hit): 8 registers saved
    const size_t warps = (size_t)gridDim.x * 8;
    for (size_t wi = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5); wi * PPW < total; wi += warps) {
        const size_t pix = wi * PPW + sub;
        const bool live = pix < total;
        const size_t pc = live ? pix : total - 1;           // keep the whole warp in the shuffles
        const size_t b = pc / hw, p = pc % hw;
        // the flows of every source first: a pixel none of whose taps lands inside a source map (background, most of a
        // frame) has logits 0 and value bias_v for every source, so its target features are never needed
        float2 gs[NSMAX];
        bool any_tap = false;
#pragma unroll
        for (int s = 0; s < NSMAX; s++) {
            if (s < ns) {
                gs[s] = __ldg(reinterpret_cast<const float2*>(T) + (b * ns + s) * hw + p);
                const Taps t0 = bilinear_taps(gs[s].x, gs[s].y, h, w);
                any_tap |= (t0.off[0] >= 0) | (t0.off[1] >= 0) | (t0.off[2] >= 0) | (t0.off[3] >= 0);
            }
        }
        float xv[8];
        if (any_tap) {
            load_planes8(xt, xt_planes, xt_plane_stride, pc * xt_pitch + xt_coff + cg * 8, xv);
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) xv[j] = 0.f;
        }
        // online softmax over the sources: running max m, denominator den and weighted value sum o8
        float m = -INFINITY, den = 0.f, o8[8];
#pragma unroll
        for (int j = 0; j < 8; j++) o8[j] = 0.f;
#pragma unroll 1
        for (int s = 0; s < ns; s++) {
            float2 g = gs[0];
#pragma unroll
            for (int q = 1; q < NSMAX; q++) g = (s == q) ? gs[q] : g;      // register select, no local-memory indexing
            const Taps t = bilinear_taps(g.x, g.y, h, w);
            const float* src = kv + (size_t)s * hw * KVP;
            float kk[8], vv[8], k0 = 0.f;
            {
                const float4 b0 = __ldg(bvp), b1 = __ldg(bvp + 1);
                vv[0] = b0.x; vv[1] = b0.y; vv[2] = b0.z; vv[3] = b0.w; vv[4] = b1.x; vv[5] = b1.y; vv[6] = b1.z; vv[7] = b1.w;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) kk[j] = 0.f;
            if constexpr (WIDE == 1) {
                float4 q[4][4];
                float k0c[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float* pxf = src + (size_t)(t.off[i] < 0 ? 0 : t.off[i]) * KVP;     // weight 0 when out of range
                    const float4* px = reinterpret_cast<const float4*>(pxf + cg * 8);
                    q[i][0] = __ldg(px); q[i][1] = __ldg(px + 1); q[i][2] = __ldg(px + C / 4); q[i][3] = __ldg(px + C / 4 + 1);
                    k0c[i] = (cg == 0) ? __ldg(pxf + 2 * C) : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (t.off[i] < 0) continue;
                    const float wt = t.wt[i];
                    k0 += k0c[i] * wt;
                    kk[0] += q[i][0].x * wt; kk[1] += q[i][0].y * wt; kk[2] += q[i][0].z * wt; kk[3] += q[i][0].w * wt;
                    kk[4] += q[i][1].x * wt; kk[5] += q[i][1].y * wt; kk[6] += q[i][1].z * wt; kk[7] += q[i][1].w * wt;
                    vv[0] += q[i][2].x * wt; vv[1] += q[i][2].y * wt; vv[2] += q[i][2].z * wt; vv[3] += q[i][2].w * wt;
                    vv[4] += q[i][3].x * wt; vv[5] += q[i][3].y * wt; vv[6] += q[i][3].z * wt; vv[7] += q[i][3].w * wt;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (t.off[i] < 0) continue;
                    const float* pxf = src + (size_t)t.off[i] * KVP;
                    const float4* px = reinterpret_cast<const float4*>(pxf + cg * 8);
                    const float4 k0v = __ldg(px), k1 = __ldg(px + 1), v0 = __ldg(px + C / 4), v1 = __ldg(px + C / 4 + 1);
                    const float wt = t.wt[i];
                    if (cg == 0) k0 += __ldg(pxf + 2 * C) * wt;
                    kk[0] += k0v.x * wt; kk[1] += k0v.y * wt; kk[2] += k0v.z * wt; kk[3] += k0v.w * wt;
                    kk[4] += k1.x * wt; kk[5] += k1.y * wt; kk[6] += k1.z * wt; kk[7] += k1.w * wt;
                    vv[0] += v0.x * wt; vv[1] += v0.y * wt; vv[2] += v0.z * wt; vv[3] += v0.w * wt;
                    vv[4] += v1.x * wt; vv[5] += v1.y * wt; vv[6] += v1.z * wt; vv[7] += v1.w * wt;
                }
            }
            float dot = k0;                                  // only the cg == 0 lane carries warp(k0)
#pragma unroll
            for (int j = 0; j < 8; j++) dot += kk[j] * xv[j];
#pragma unroll
            for (int o = LPP / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            const float logit = dot * inv_sqrt;
            const float m_new = fmaxf(m, logit);
            const float scale = expf(m - m_new), e = expf(logit - m_new);      // first source: exp(-inf) = 0
            den = den * scale + e;
#pragma unroll
            for (int j = 0; j < 8; j++) o8[j] = o8[j] * scale + e * vv[j];
            m = m_new;
        }
        const float rden = 1.f / den;
#pragma unroll
        for (int j = 0; j < 8; j++) o8[j] *= rden;
        if (live) store_planes8(out, out_planes, out_plane_stride, pix * out_pitch + out_coff + cg * 8, o8);
    }
}


// Chunked schedule (default): a warp takes 32 CONSECUTIVE pixels.  Lane i reads the flows of pixel i (coalesced) and classifies
// it; the ballot splits the chunk into background pixels (no tap of any source lands inside a map: their output is the same
// constant for every such pixel — softmax over equal logits of the value bias — so the warp just streams that row out, no
// dependent chain) and foreground pixels, which run the gather / online-softmax path PPW at a time with their flows
// broadcast by shuffle.  On real frames ~85-90 % of the pixels are background: the per-pixel latency chain of the
// pixel-per-warp schedule (flow load -> taps -> [gathers] -> shuffles -> exp -> store) is paid only where it is needed.
template <int C, int NSMAX>
__global__ void __launch_bounds__(256, 3) warp_attention_chunk_kernel(const __half* __restrict__ xt, int xt_planes,
                                                                      long long xt_plane_stride, int xt_pitch, int xt_coff,
                                                                      const float* __restrict__ kv,
                                                                      const float* __restrict__ bias_v,
                                                                      const float* __restrict__ T, int B, int ns, int h, int w,
                                                                      __half* __restrict__ out, int out_planes,
                                                                      long long out_plane_stride, int out_pitch, int out_coff) {
    constexpr int LPP = C / 8, PPW = 32 / LPP, KVP = 2 * C + 64;
    // pixels per chunk: a foreground pixel costs a full gather chain of PPW-wide steps, so the wider a pixel is in lanes the
    // fewer pixels a warp takes (more warps share the foreground of a region): 32 / 16 / 8 for C = 64 / 128 / 256
    constexpr int CH = 8 * PPW;
    const int lane = threadIdx.x & 31, cg = lane % LPP, sub = lane / LPP;
    const size_t hw = (size_t)h * w, total = (size_t)B * hw;
    const float inv_sqrt = 1.f / sqrtf((float)C);
    const float4* bvp = reinterpret_cast<const float4*>(bias_v + cg * 8);
    float bv8[8];
    {
        const float4 b0 = __ldg(bvp), b1 = __ldg(bvp + 1);
        bv8[0] = b0.x; bv8[1] = b0.y; bv8[2] = b0.z; bv8[3] = b0.w; bv8[4] = b1.x; bv8[5] = b1.y; bv8[6] = b1.z; bv8[7] = b1.w;
    }
    // background row: the same online-softmax arithmetic as the foreground path with logit 0 and value bias_v for every source
    float bg8[8];
    {
        float m = -INFINITY, den = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) bg8[j] = 0.f;
        for (int s = 0; s < ns; s++) {
            const float m_new = fmaxf(m, 0.f);
            const float scale = expf(m - m_new), e = expf(0.f - m_new);
            den = den * scale + e;
#pragma unroll
            for (int j = 0; j < 8; j++) bg8[j] = bg8[j] * scale + e * bv8[j];
            m = m_new;
        }
        const float rden = 1.f / den;
#pragma unroll
        for (int j = 0; j < 8; j++) bg8[j] *= rden;
    }
    const size_t nchunks = (total + CH - 1) / CH, warps = (size_t)gridDim.x * 8;
    for (size_t ch = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5); ch < nchunks; ch += warps) {
        const size_t pix_l = ch * CH + lane;
        const bool live_l = lane < CH && pix_l < total;
        const size_t pl = live_l ? pix_l : total - 1;
        const size_t b_l = pl / hw, p_l = pl % hw;
        float2 gs[NSMAX];
        bool any_l = false;
#pragma unroll
        for (int s = 0; s < NSMAX; s++) {
            gs[s] = make_float2(-2.f, -2.f);
            if (s < ns) {
                gs[s] = __ldg(reinterpret_cast<const float2*>(T) + (b_l * ns + s) * hw + p_l);
                const Taps t0 = bilinear_taps(gs[s].x, gs[s].y, h, w);
                any_l |= (t0.off[0] >= 0) | (t0.off[1] >= 0) | (t0.off[2] >= 0) | (t0.off[3] >= 0);
            }
        }
        const unsigned live_mask = __ballot_sync(0xffffffffu, live_l);
        const unsigned fg_mask = __ballot_sync(0xffffffffu, live_l && any_l);
        const unsigned bg_mask = live_mask & ~fg_mask;
        // ---- background: stream the constant row, PPW pixels per store instruction ----
        const int nbg = __popc(bg_mask);
        for (int k = sub; k < nbg; k += PPW) {
            const int j = __fns(bg_mask, 0, k + 1);
            store_planes8(out, out_planes, out_plane_stride, (ch * CH + j) * out_pitch + out_coff + cg * 8, bg8);
        }
        // ---- foreground: PPW pixels at a time (every lane takes part in the shuffles) ----
        const int nfg = __popc(fg_mask);
        for (int k0 = 0; k0 < nfg; k0 += PPW) {
            const int k = k0 + sub;
            const bool act = k < nfg;
            const int j = act ? __fns(fg_mask, 0, k + 1) : __ffs(fg_mask) - 1;      // idle sub-groups shadow a valid pixel
            const size_t pix = ch * CH + j;
            float xv[8];
            load_planes8(xt, xt_planes, xt_plane_stride, pix * xt_pitch + xt_coff + cg * 8, xv);
            float m = -INFINITY, den = 0.f, o8[8];
#pragma unroll
            for (int q = 0; q < 8; q++) o8[q] = 0.f;
#pragma unroll
            for (int s = 0; s < NSMAX; s++) {
                if (s < ns) {                                   // ns is warp-uniform
                    const float gx = __shfl_sync(0xffffffffu, gs[s].x, j), gy = __shfl_sync(0xffffffffu, gs[s].y, j);
                    const Taps t = bilinear_taps(gx, gy, h, w);
                    const float* src = kv + (size_t)s * hw * KVP;
                    float kk[8], vv[8], k0v = 0.f;
#pragma unroll
                    for (int q = 0; q < 8; q++) { kk[q] = 0.f; vv[q] = bv8[q]; }
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (t.off[i] < 0) continue;
                        const float* pxf = src + (size_t)t.off[i] * KVP;
                        const float4* px = reinterpret_cast<const float4*>(pxf + cg * 8);
                        const float4 ka = __ldg(px), kb = __ldg(px + 1), va = __ldg(px + C / 4), vb = __ldg(px + C / 4 + 1);
                        const float wt = t.wt[i];
                        if (cg == 0) k0v += __ldg(pxf + 2 * C) * wt;
                        kk[0] += ka.x * wt; kk[1] += ka.y * wt; kk[2] += ka.z * wt; kk[3] += ka.w * wt;
                        kk[4] += kb.x * wt; kk[5] += kb.y * wt; kk[6] += kb.z * wt; kk[7] += kb.w * wt;
                        vv[0] += va.x * wt; vv[1] += va.y * wt; vv[2] += va.z * wt; vv[3] += va.w * wt;
                        vv[4] += vb.x * wt; vv[5] += vb.y * wt; vv[6] += vb.z * wt; vv[7] += vb.w * wt;
                    }
                    float dot = k0v;
#pragma unroll
                    for (int q = 0; q < 8; q++) dot += kk[q] * xv[q];
#pragma unroll
                    for (int o = LPP / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                    const float logit = dot * inv_sqrt;
                    const float m_new = fmaxf(m, logit);
                    const float scale = expf(m - m_new), e = expf(logit - m_new);
                    den = den * scale + e;
#pragma unroll
                    for (int q = 0; q < 8; q++) o8[q] = o8[q] * scale + e * vv[q];
                    m = m_new;
                }
            }
            const float rden = 1.f / den;
#pragma unroll
            for (int q = 0; q < 8; q++) o8[q] *= rden;
            if (act) store_planes8(out, out_planes, out_plane_stride, pix * out_pitch + out_coff + cg * 8, o8);
        }
    }
}

// plain warp (LWB.transform) on NHWC fp32, for the seam and for debugging
__global__ void __launch_bounds__(256) warp_nhwc_kernel(const float* __restrict__ src, const float* __restrict__ T, int B,
                                                        int ns, int h, int w, int C, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const size_t hw = (size_t)h * w, total = (size_t)B * ns * hw;
    const int nvec = C / 4;
    for (size_t i = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < total; i += (size_t)gridDim.x * 8) {
        const size_t s = (i / hw) % ns;
        const float2 g = reinterpret_cast<const float2*>(T)[i];
        const Taps t = bilinear_taps(g.x, g.y, h, w);
        const float4* sp = reinterpret_cast<const float4*>(src + s * hw * C);
        for (int cv = lane; cv < nvec; cv += 32) {
            float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (t.off[k] < 0) continue;
                const float4 a = __ldg(sp + (size_t)t.off[k] * nvec + cv);
                acc.x += a.x * t.wt[k]; acc.y += a.y * t.wt[k]; acc.z += a.z * t.wt[k]; acc.w += a.w * t.wt[k];
            }
            reinterpret_cast<float4*>(out + i * C)[cv] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// layout converters
// ------------------------------------------------------------------------------------------------------------
__global__ void nchw_to_planes_kernel(const float* __restrict__ in, int N, int C, int HW, __half* __restrict__ out,
                                      int out_planes, long long out_plane_stride, int out_pitch, int out_coff) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, p = p0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && p < HW) ? in[((size_t)n * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int p = p0 + r, c = c0 + threadIdx.x;
        if (p < HW && c < C)
            store_plane_val(out, out_planes, out_plane_stride, ((size_t)n * HW + p) * out_pitch + out_coff + c,
                            tile[threadIdx.x][r]);
    }
}
__global__ void planes_to_nchw_kernel(const __half* __restrict__ x, int x_planes, long long x_plane_stride, int N, int C,
                                      int HW, int x_pitch, int x_coff, const float* __restrict__ xf,
                                      float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int p = p0 + r, c = c0 + threadIdx.x;
        float v = 0.f;
        if (p < HW && c < C) {
            const size_t off = ((size_t)n * HW + p) * x_pitch + x_coff + c;
            v = xf ? xf[off] : load_plane_val(x, x_planes, x_plane_stride, off);
        }
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, p = p0 + threadIdx.x;
        if (c < C && p < HW) out[((size_t)n * C + c) * HW + p] = tile[threadIdx.x][r];
    }
}

__global__ void pred_to_u8_kernel(const float* __restrict__ pred, int B, size_t SS, uint8_t* __restrict__ out) {
    const size_t total = (size_t)B * SS;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / SS, p = i % SS;
        uint8_t bgr[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float v = (pred[(b * 3 + c) * SS + p] + 1.f) / 2.0f * 255.f;   // cv_utils.py:111-113, then astype(uint8)
            v = fminf(fmaxf(v, 0.f), 255.f);
            bgr[2 - c] = (uint8_t)v;                                      // RGB -> BGR (cv_utils.py:104-105)
        }
        out[i * 3 + 0] = bgr[0]; out[i * 3 + 1] = bgr[1]; out[i * 3 + 2] = bgr[2];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Mask morphology of source_setup (iPERCore/tools/utils/morphology/morph_ops.py:7-61): ks x ks box sum of a
// (N,1,H,W) mask with constant border padding, then a threshold.  One CTA per 32x32 output tile: the padded
// (32+ks-1)^2 input patch lives in shared memory, a horizontal pass builds row sums, a vertical pass finishes —
// O(ks) per pass instead of the reference's ks^2 dense convolution.  Exact for 0/1 masks (integer sums in fp32).
// ------------------------------------------------------------------------------------------------------------
constexpr int MORPH_T = 32, MORPH_MAX_KS = 63;
__global__ void __launch_bounds__(256) morph_kernel(const float* __restrict__ mask, int H, int W, int ks, int mode,
                                                    float* __restrict__ out) {
    extern __shared__ float msm[];
    const int p = ks / 2, PW = MORPH_T + ks - 1;       // padded patch is PW x PW
    float* s_in = msm;                                  // [PW][PW]
    float* s_h = msm + PW * PW;                         // [PW][MORPH_T] horizontal sums
    const int n = blockIdx.z, y0 = blockIdx.y * MORPH_T, x0 = blockIdx.x * MORPH_T;
    const float padv = (mode == 0) ? 1.f : 0.f;         // erode pads with 1, dilate / soft dilate with 0
    const float* src = mask + (size_t)n * H * W;
    for (int i = threadIdx.x; i < PW * PW; i += 256) {
        const int r = i / PW, c = i % PW, y = y0 + r - p, x = x0 + c - p;
        s_in[i] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(src + (size_t)y * W + x) : padv;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PW * MORPH_T; i += 256) {
        const int r = i / MORPH_T, c = i % MORPH_T;
        float acc = 0.f;
        for (int k = 0; k < ks; k++) acc += s_in[r * PW + c + k];
        s_h[i] = acc;
    }
    __syncthreads();
    const float n_ks = (float)(ks * ks);
    for (int i = threadIdx.x; i < MORPH_T * MORPH_T; i += 256) {
        const int r = i / MORPH_T, c = i % MORPH_T, y = y0 + r, x = x0 + c;
        if (y >= H || x >= W) continue;
        float acc = 0.f;
        for (int k = 0; k < ks; k++) acc += s_h[(r + k) * MORPH_T + c];
        const bool on = mode == 0 ? (acc == n_ks) : (mode == 1 ? (acc >= 1.f) : (acc >= n_ks / 2));
        out[(size_t)n * H * W + (size_t)y * W + x] = on ? 1.f : 0.f;
    }
}

}  // namespace iper

using namespace iper;

extern "C" int iper_morph(const float* mask, int N, int H, int W, int ks, int mode, float* out, iper_stream_t stream) {
    IPER_REQUIRE(N >= 0 && H > 0 && W > 0, "iper_morph: bad sizes");
    IPER_REQUIRE(ks >= 1 && ks <= MORPH_MAX_KS && (ks & 1), "iper_morph: ks=%d must be odd and <= %d", ks, MORPH_MAX_KS);
    IPER_REQUIRE(mode >= 0 && mode <= 2, "iper_morph: mode %d not in {0 erode, 1 dilate, 2 soft dilate}", mode);
    if (N == 0) return 0;
    IPER_REQUIRE(mask && out, "iper_morph: null pointer");
    const int PW = MORPH_T + ks - 1;
    const size_t smem = sizeof(float) * ((size_t)PW * PW + (size_t)PW * MORPH_T);
    IPER_CHECK_CUDA(cudaFuncSetAttribute(morph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((W + MORPH_T - 1) / MORPH_T, (H + MORPH_T - 1) / MORPH_T, N);
    morph_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(mask, H, W, ks, mode, out);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_conv_direct(const iper_conv_gemm_desc* g, const float* w_f32, int Cout, iper_stream_t stream) {
    IPER_REQUIRE(g && w_f32 && g->a && g->out, "iper_conv_direct: null pointer");
    IPER_REQUIRE(g->epi != IPER_EPI_HEADS, "iper_conv_direct: heads epilogue not supported; use IPER_EPI_F32 and post-process");
    DirectArgs d = {};
    d.a = reinterpret_cast<const __half*>(g->a); d.a_planes = g->a_planes; d.a_plane_stride = g->a_plane_stride;
    d.N = g->N; d.H = g->H; d.W = g->W; d.a_pitch = g->a_pitch; d.a_coff = g->a_coff; d.Cin = g->Cin;
    d.mode = g->mode; d.ksize = g->ksize; d.Cout = Cout;
    d.oH = g->mode == IPER_CONV_S2 ? g->H / 2 : (g->mode == IPER_CONVT_4S2 ? 2 * g->H : g->H);
    d.oW = g->mode == IPER_CONV_S2 ? g->W / 2 : (g->mode == IPER_CONVT_4S2 ? 2 * g->W : g->W);
    d.w = w_f32; d.bias = g->bias; d.relu = g->relu; d.epi = g->epi;
    d.out = g->out; d.out_planes = g->out_planes; d.out_plane_stride = g->out_plane_stride;
    d.out_pitch = g->out_pitch; d.out_coff = g->out_coff;
    d.x = reinterpret_cast<const __half*>(g->x); d.x_planes = g->x_planes; d.x_plane_stride = g->x_plane_stride;
    d.x_pitch = g->x_pitch; d.x_coff = g->x_coff; d.mean_rstd = g->mean_rstd; d.spade_C = g->spade_C;
    if (g->epi == IPER_EPI_SPADE)
        IPER_REQUIRE(g->x && g->mean_rstd && g->bias && Cout == 2 * g->spade_C, "iper_conv_direct: bad SPADE arguments");
    const size_t total = (size_t)d.N * d.oH * d.oW * (g->epi == IPER_EPI_SPADE ? d.spade_C : Cout);
    if (total == 0) return 0;
    const int blocks = (int)min((size_t)148 * 32, (total + 255) / 256);
    conv_direct_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_conv_stem(const float* in_nchw, int N, int Cin, int H, int W, const float* w_f32, const float* bias,
                              int Cout, void* out, int out_planes, long long out_plane_stride, int out_pitch,
                              int out_coff, double* stats_ws, iper_stream_t stream) {
    IPER_REQUIRE(in_nchw && w_f32 && out, "iper_conv_stem: null pointer");
    IPER_REQUIRE(Cin >= 1 && Cin <= STEM_MAXC, "iper_conv_stem: Cin=%d not in [1,8]", Cin);
    IPER_REQUIRE(Cout % 8 == 0 && Cout <= 128, "iper_conv_stem: Cout=%d must be a multiple of 8, <= 128", Cout);
    IPER_REQUIRE(H % 2 == 0 && W % 2 == 0, "iper_conv_stem: H, W must be even");
    IPER_REQUIRE(out_pitch % 8 == 0 && out_coff % 8 == 0, "iper_conv_stem: output window must be 8-aligned");
    if (stats_ws) IPER_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * (size_t)N * Cout, (cudaStream_t)stream));
    const size_t smem = sizeof(float) * ((size_t)stem_halo_floats(Cin) + (size_t)9 * Cin * Cout);
    dim3 grid((W / 2 + STEM_TW - 1) / STEM_TW, (H / 2 + STEM_TH - 1) / STEM_TH, N);
#define IPER_STEM_CASE(CI)                                                                                              \
    case CI:                                                                                                            \
        IPER_CHECK_CUDA(cudaFuncSetAttribute(conv_stem_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        conv_stem_kernel<CI><<<grid, 256, smem, (cudaStream_t)stream>>>(in_nchw, N, H, W, w_f32, bias, Cout,              \
                                                                       reinterpret_cast<__half*>(out), out_planes,     \
                                                                       out_plane_stride, out_pitch, out_coff, stats_ws); \
        break;
    switch (Cin) {
        IPER_STEM_CASE(1) IPER_STEM_CASE(2) IPER_STEM_CASE(3) IPER_STEM_CASE(4)
        IPER_STEM_CASE(5) IPER_STEM_CASE(6) IPER_STEM_CASE(7) IPER_STEM_CASE(8)
    }
#undef IPER_STEM_CASE
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_instnorm_finalize(const double* workspace, int N, int C, int HW, float eps, float* mean_rstd,
                                      iper_stream_t stream) {
    IPER_REQUIRE(workspace && mean_rstd, "iper_instnorm_finalize: null pointer");
    if (N * C == 0) return 0;
    instnorm_final_kernel<<<(N * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(workspace, N * C, HW, eps, mean_rstd);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_instnorm_stats(const void* x, int x_planes, long long x_plane_stride, int N, int HW, int C,
                                   int x_pitch, int x_coff, float eps, double* workspace, float* mean_rstd,
                                   iper_stream_t stream) {
    IPER_REQUIRE(x && mean_rstd && workspace, "iper_instnorm_stats: null pointer");
    IPER_REQUIRE(C % 8 == 0 && C >= 8 && C <= 256 && 256 % (C / 8) == 0,
                 "iper_instnorm_stats: C=%d must be 8..256 with C/8 a power of two", C);
    IPER_REQUIRE(x_pitch % 8 == 0 && x_coff % 8 == 0, "iper_instnorm_stats: channel window must be 8-aligned");
    if (N == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    IPER_CHECK_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * (size_t)N * C, s));
    // enough CTAs to fill the machine (~4 per SM), at least 256 pixels per CTA
    int splits = (148 * 4 + N - 1) / N;
    const int max_splits = (HW + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int chunk = (HW + splits - 1) / splits;
    dim3 grid((HW + chunk - 1) / chunk, N);
    instnorm_partial_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(x), x_planes, x_plane_stride, HW, C,
                                                 x_pitch, x_coff, chunk, workspace);
    IPER_CHECK_CUDA(cudaGetLastError());
    instnorm_final_kernel<<<(N * C + 255) / 256, 256, 0, s>>>(workspace, N * C, HW, eps, mean_rstd);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_warp_attention(const void* xt, int xt_planes, long long xt_plane_stride, int xt_pitch, int xt_coff,
                                   const float* kv, const float* bias_v, const float* T, int B, int ns, int h, int w,
                                   int C, void* out, int out_planes, long long out_plane_stride, int out_pitch,
                                   int out_coff, iper_stream_t stream) {
    IPER_REQUIRE(xt && kv && bias_v && T && out, "iper_warp_attention: null pointer");
    IPER_REQUIRE(ns >= 1 && ns <= ATT_MAX_NS, "iper_warp_attention: ns=%d not in [1,%d]", ns, ATT_MAX_NS);
    IPER_REQUIRE(C == 64 || C == 128 || C == 256, "iper_warp_attention: C=%d not in {64,128,256}", C);
    IPER_REQUIRE(out_pitch % 8 == 0 && out_coff % 8 == 0 && xt_pitch % 8 == 0 && xt_coff % 8 == 0,
                 "iper_warp_attention: channel windows must be 8-aligned");
    const size_t total = (size_t)B * h * w;
    if (total == 0) return 0;
    const int ppw = 32 / (C / 8);
    const size_t nwarps = (total + ppw - 1) / ppw;
    const int blocks = (int)min((size_t)148 * 16, (nwarps + 7) / 8);
    __half* o = reinterpret_cast<__half*>(out);
    const __half* x = reinterpret_cast<const __half*>(xt);
    cudaStream_t st = (cudaStream_t)stream;
    // gather schedule (see warp_attention_kernel).  Hoisting the 16 gathers only pays on dense synthetic flows at C = 64
    // (1.58 vs 1.73 ms per 50 frames); on real frames most pixels are background whose taps the narrow schedule skips
    // outright (ncu: 406 us narrow vs 542 us wide for the C = 64 stage of a 20-frame batch), so narrow is the default.
    // default: the chunked schedule (warp_attention_chunk_kernel); IPER_ATT_WIDE = 0 / 1 / 2 select the pixel-per-warp
    // schedules above for comparison (read once)
    static const int wide = [] { const char* v = getenv("IPER_ATT_WIDE"); return v ? atoi(v) : 3; }();
#define IPER_ATT(CV, NV)                                                                                             \
    do {                                                                                                             \
        if (wide == 3) {                                                                                             \
            const size_t cpx = (size_t)(CV == 256 ? 8 : (CV == 128 ? 16 : 32));       /* pixels per warp chunk */       \
            const int cblocks = (int)min((size_t)148 * 12, ((total + cpx - 1) / cpx + 7) / 8);                        \
            warp_attention_chunk_kernel<CV, NV><<<cblocks, 256, 0, st>>>(x, xt_planes, xt_plane_stride, xt_pitch, xt_coff, kv, \
                                                                         bias_v, T, B, ns, h, w, o, out_planes,         \
                                                                         out_plane_stride, out_pitch, out_coff);       \
        } else if (wide == 1)                                                                                               \
            warp_attention_kernel<CV, NV, 1><<<blocks, 256, 0, st>>>(x, xt_planes, xt_plane_stride, xt_pitch, xt_coff, kv, \
                                                                     bias_v, T, B, ns, h, w, o, out_planes,             \
                                                                     out_plane_stride, out_pitch, out_coff);           \
        else if (wide == 2)                                                                                          \
            warp_attention_kernel<CV, NV, 2><<<blocks, 256, 0, st>>>(x, xt_planes, xt_plane_stride, xt_pitch, xt_coff, kv, \
                                                                     bias_v, T, B, ns, h, w, o, out_planes,             \
                                                                     out_plane_stride, out_pitch, out_coff);           \
        else                                                                                                         \
            warp_attention_kernel<CV, NV, 0><<<blocks, 256, 0, st>>>(x, xt_planes, xt_plane_stride, xt_pitch, xt_coff, kv, \
                                                                     bias_v, T, B, ns, h, w, o, out_planes,             \
                                                                     out_plane_stride, out_pitch, out_coff);           \
    } while (0)
#define IPER_ATT_C(CV)                                                                                               \
    do {                                                                                                             \
        if (ns <= 2) IPER_ATT(CV, 2); else if (ns <= 4) IPER_ATT(CV, 4); else IPER_ATT(CV, 8);                       \
    } while (0)
    if (C == 64) IPER_ATT_C(64); else if (C == 128) IPER_ATT_C(128); else IPER_ATT_C(256);
#undef IPER_ATT_C
#undef IPER_ATT
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_stem_im2col(const float* in_nchw, int N, int Cin, int H, int W, void* out, int out_planes,
                                long long out_plane_stride, int out_pitch, int out_coff, iper_stream_t stream) {
    IPER_REQUIRE(in_nchw && out, "iper_stem_im2col: null pointer");
    IPER_REQUIRE(N >= 0 && Cin >= 1 && 9 * Cin <= 64 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0,
                 "iper_stem_im2col: needs Cin <= 7 and even H, W (got Cin=%d, %dx%d)", Cin, H, W);
    IPER_REQUIRE(out_pitch % 8 == 0 && out_coff % 8 == 0 && out_coff + 64 <= out_pitch, "iper_stem_im2col: bad output channel window");
    IPER_REQUIRE(out_planes == 1 || out_planes == 2, "iper_stem_im2col: output format %d not in {1,2}", out_planes);
    const size_t total = (size_t)N * (H / 2) * (W / 2) * 8;
    if (total == 0) return 0;
    const int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
#define IPER_IM2COL(CV)                                                                                                     \
    stem_im2col_kernel<CV><<<blocks, 256, 0, (cudaStream_t)stream>>>(in_nchw, N, H, W, reinterpret_cast<__half*>(out), out_planes, \
                                                                     out_plane_stride, out_pitch, out_coff)
    switch (Cin) {
        case 1: IPER_IM2COL(1); break; case 2: IPER_IM2COL(2); break; case 3: IPER_IM2COL(3); break; case 4: IPER_IM2COL(4); break;
        case 5: IPER_IM2COL(5); break; case 6: IPER_IM2COL(6); break; default: IPER_IM2COL(7); break;
    }
#undef IPER_IM2COL
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_warp_nhwc(const float* src, const float* T, int B, int ns, int h, int w, int C, float* out,
                              iper_stream_t stream) {
    IPER_REQUIRE(src && T && out, "iper_warp_nhwc: null pointer");
    IPER_REQUIRE(C % 4 == 0, "iper_warp_nhwc: C must be a multiple of 4");
    const size_t total = (size_t)B * ns * h * w;
    if (total == 0) return 0;
    const int blocks = (int)min((size_t)148 * 8, (total + 7) / 8);
    warp_nhwc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, T, B, ns, h, w, C, out);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_nchw_to_planes(const float* in, int N, int C, int HW, void* out, int out_planes,
                                   long long out_plane_stride, int out_pitch, int out_coff, iper_stream_t stream) {
    IPER_REQUIRE(in && out, "iper_nchw_to_planes: null pointer");
    if (N == 0) return 0;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
    nchw_to_planes_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, N, C, HW, reinterpret_cast<__half*>(out),
                                                                   out_planes, out_plane_stride, out_pitch, out_coff);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_planes_to_nchw(const void* x, int x_planes, long long x_plane_stride, int N, int C, int HW,
                                   int x_pitch, int x_coff, float* out, iper_stream_t stream) {
    IPER_REQUIRE(x && out, "iper_planes_to_nchw: null pointer");
    if (N == 0) return 0;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
    planes_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(x), x_planes,
                                                                   x_plane_stride, N, C, HW, x_pitch, x_coff, nullptr, out);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_nhwc_f32_to_nchw(const float* in, int N, int C, int HW, int pitch, int coff, float* out,
                                     iper_stream_t stream) {
    IPER_REQUIRE(in && out, "iper_nhwc_f32_to_nchw: null pointer");
    if (N == 0) return 0;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
    planes_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(nullptr, 0, 0, N, C, HW, pitch, coff, in, out);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_pred_to_u8(const float* pred, int B, int S, uint8_t* out, iper_stream_t stream) {
    IPER_REQUIRE(pred && out, "iper_pred_to_u8: null pointer");
    const size_t total = (size_t)B * S * S;
    if (total == 0) return 0;
    const int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
    pred_to_u8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, B, (size_t)S * S, out);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_instnorm_apply(const void* x, int x_planes, long long x_plane_stride, int x_pitch, int x_coff,
                                   const float* mean_rstd, int N, int HW, int C, int relu, const void* res,
                                   int res_planes, long long res_plane_stride, int res_pitch, int res_coff, void* out,
                                   int out_planes, long long out_plane_stride, int out_pitch, int out_coff,
                                   iper_stream_t stream) {
    IPER_REQUIRE(x && mean_rstd && out, "iper_instnorm_apply: null pointer");
    const size_t total = (size_t)N * HW * C;
    if (total == 0) return 0;
    const int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
    instnorm_apply_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __half*>(x), x_planes, x_plane_stride, x_pitch, x_coff, mean_rstd, N, HW, C, relu,
        reinterpret_cast<const __half*>(res), res_planes, res_plane_stride, res_pitch, res_coff,
        reinterpret_cast<__half*>(out), out_planes, out_plane_stride, out_pitch, out_coff);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_tanh_nhwc_to_nchw(const float* in, int N, int HW, int C, int pitch, float* out, iper_stream_t stream) {
    IPER_REQUIRE(in && out, "iper_tanh_nhwc_to_nchw: null pointer");
    const size_t total = (size_t)N * HW * C;
    if (total == 0) return 0;
    const int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
    tanh_nhwc_to_nchw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, N, HW, C, pitch, out);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}
