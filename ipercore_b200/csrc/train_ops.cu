// Training-step kernels, part 2: the non-GEMM pieces of SelfAttentionLWB and the instance norms, forward
// and backward, on NHWC bf16 (torch channels_last) — all HBM-bound, so every kernel reads each tensor once with 16-byte
// accesses (8 channels per thread) and never converts a layout or a dtype outside its own loads/stores:
//   warp            LWB.transform's grid_sample (attlwb_spade_resunet.py:184-191: bilinear, zeros, align_corners=False) and its
//                   gradient w.r.t. the source features (float4 atomics into an fp32 buffer; the flow carries no gradient).
//   att_combine     softmax over the ns sources of (K_s . q) / sqrt(C) and a = sum_s alpha_s V_s (attlwb_spade_resunet.py:121-139,
//                   232-240), with alpha kept for the backward pass (dK, dV, dq in one kernel).
//   norm            InstanceNorm2d(affine=False) statistics + [SPADE modulation IN(x) (1 + gamma) + beta | plain] + [ReLU |
//                   LeakyReLU] in one apply pass (attlwb_spade_resunet.py:80-93, bg_inpaintor.py, patch_dis.py); backward in two
//                   passes (per-(n,c) sums of dxh and dxh*xh, then dx), which also emits dgamma and dbeta.
// The reference runs these through ATen (grid_sampler_2d, batch_norm, softmax, a dozen elementwise kernels per block,
// iPERCore/tools/trainers/lwg_trainer.py:699-833).
#include <cuda_bf16.h>

#include "common.cuh"
#include "iper_b200.h"

namespace iper {

struct Vec8 { float v[8]; };

IPER_DEVINL Vec8 ld8(const __nv_bfloat16* p) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    Vec8 r;
#pragma unroll
    for (int i = 0; i < 4; i++) { const float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
    return r;
}
IPER_DEVINL void st8(__nv_bfloat16* p, const Vec8& a) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(a.v[2 * i], a.v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}

// ---------------------------------------------------------------------------------------------------------------------
// warp
// ---------------------------------------------------------------------------------------------------------------------
struct Taps { int x0, y0; float wx, wy; bool any; };
IPER_DEVINL Taps taps_of(const float* T, long long pix, int h, int w) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(T) + pix);
    const float ix = ((t.x + 1.f) * w - 1.f) * 0.5f, iy = ((t.y + 1.f) * h - 1.f) * 0.5f;   // align_corners = False
    Taps r;
    const float fx = floorf(ix), fy = floorf(iy);
    r.x0 = (int)fx; r.y0 = (int)fy; r.wx = ix - fx; r.wy = iy - fy;
    r.any = r.x0 >= -1 && r.x0 < w && r.y0 >= -1 && r.y0 < h;      // false for NaN too
    return r;
}

template <bool BWD>
__global__ void __launch_bounds__(256) warp_bf16_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ T, int M, int h,
                                                        int w, int C, __nv_bfloat16* __restrict__ out, float* __restrict__ dsrc) {
    const int groups = C / 8;
    const long long total = (long long)M * h * w * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % groups);
        const long long pix = i / groups;
        const int m = (int)(pix / ((long long)h * w));
        const Taps t = taps_of(T, pix, h, w);
        Vec8 acc;
#pragma unroll
        for (int c = 0; c < 8; c++) acc.v[c] = 0.f;
        Vec8 g;
        if (BWD) { if (!t.any) continue; g = ld8(in + pix * C + cg * 8); }
        if (t.any) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
                if (xx < 0 || xx >= w || yy < 0 || yy >= h) continue;
                const float wt = ((k & 1) ? t.wx : 1.f - t.wx) * ((k >> 1) ? t.wy : 1.f - t.wy);
                const long long off = (((long long)m * h + yy) * w + xx) * C + cg * 8;
                if (BWD) {
                    float4* d = reinterpret_cast<float4*>(dsrc + off);
                    atomicAdd(d, make_float4(wt * g.v[0], wt * g.v[1], wt * g.v[2], wt * g.v[3]));
                    atomicAdd(d + 1, make_float4(wt * g.v[4], wt * g.v[5], wt * g.v[6], wt * g.v[7]));
                } else {
                    const Vec8 s = ld8(in + off);
#pragma unroll
                    for (int c = 0; c < 8; c++) acc.v[c] += wt * s.v[c];
                }
            }
        }
        if (!BWD) st8(out + pix * C + cg * 8, acc);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// attention combine: threads of one pixel are `groups` = C/8 consecutive lanes (8, 16 or 32), reductions over C by shuffles
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ATT_MAX_NS = 8;

IPER_DEVINL float group_sum(float v, int groups) {
    for (int o = groups >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool BWD>
__global__ void __launch_bounds__(256) att_combine_kernel(const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                                                          const __nv_bfloat16* __restrict__ q, int bs, int ns, long long HW, int C,
                                                          float scale, __nv_bfloat16* __restrict__ a, float* __restrict__ alpha,
                                                          const __nv_bfloat16* __restrict__ da, __nv_bfloat16* __restrict__ dk,
                                                          __nv_bfloat16* __restrict__ dv, __nv_bfloat16* __restrict__ dq) {
    const int groups = C / 8;
    const long long total = (long long)bs * HW * groups;           // a multiple of 32: whole warps stay together
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % groups);
        const long long pix = i / groups;                           // b * HW + p
        const int b = (int)(pix / HW);
        const long long p = pix - (long long)b * HW;
        const Vec8 qv = ld8(q + pix * C + cg * 8);
        if (!BWD) {
            float logit[ATT_MAX_NS], mx = -INFINITY;
            for (int s = 0; s < ns; s++) {
                const Vec8 kv = ld8(k + (((long long)b * ns + s) * HW + p) * C + cg * 8);
                float d = 0.f;
#pragma unroll
                for (int c = 0; c < 8; c++) d += kv.v[c] * qv.v[c];
                logit[s] = group_sum(d, groups) * scale;
                mx = fmaxf(mx, logit[s]);
            }
            float den = 0.f;
            for (int s = 0; s < ns; s++) { logit[s] = __expf(logit[s] - mx); den += logit[s]; }
            Vec8 acc;
#pragma unroll
            for (int c = 0; c < 8; c++) acc.v[c] = 0.f;
            for (int s = 0; s < ns; s++) {
                const float al = logit[s] / den;
                const Vec8 vv = ld8(v + (((long long)b * ns + s) * HW + p) * C + cg * 8);
#pragma unroll
                for (int c = 0; c < 8; c++) acc.v[c] += al * vv.v[c];
                if (cg == 0) alpha[((long long)b * ns + s) * HW + p] = al;
            }
            st8(a + pix * C + cg * 8, acc);
        } else {
            const Vec8 g = ld8(da + pix * C + cg * 8);
            float al[ATT_MAX_NS], dal[ATT_MAX_NS], dot = 0.f;
            for (int s = 0; s < ns; s++) {
                const long long off = (((long long)b * ns + s) * HW + p) * C + cg * 8;
                al[s] = __ldg(alpha + ((long long)b * ns + s) * HW + p);
                const Vec8 vv = ld8(v + off);
                float d = 0.f;
                Vec8 o;
#pragma unroll
                for (int c = 0; c < 8; c++) { d += g.v[c] * vv.v[c]; o.v[c] = al[s] * g.v[c]; }
                st8(dv + off, o);
                dal[s] = group_sum(d, groups);
                dot += al[s] * dal[s];
            }
            Vec8 dqv;
#pragma unroll
            for (int c = 0; c < 8; c++) dqv.v[c] = 0.f;
            for (int s = 0; s < ns; s++) {
                const long long off = (((long long)b * ns + s) * HW + p) * C + cg * 8;
                const float dl = al[s] * (dal[s] - dot) * scale;
                const Vec8 kv = ld8(k + off);
                Vec8 o;
#pragma unroll
                for (int c = 0; c < 8; c++) { o.v[c] = dl * qv.v[c]; dqv.v[c] += dl * kv.v[c]; }
                st8(dk + off, o);
            }
            st8(dq + pix * C + cg * 8, dqv);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// instance norm (+ SPADE modulation, + activation)
// ---------------------------------------------------------------------------------------------------------------------
// block = (pixel chunk, image); thread = (8-channel group, pixel lane).  Block partial sums -> one double atomic per channel.
__global__ void __launch_bounds__(256) norm_stats_kernel(const __nv_bfloat16* __restrict__ x, long long HW, int C, double* __restrict__ stats) {
    __shared__ float red[256][17];
    const int groups = C / 8, lanes = 256 / groups;
    const int cg = threadIdx.x % groups, pl = threadIdx.x / groups, n = blockIdx.y;
    float s[8], ss[8];
#pragma unroll
    for (int c = 0; c < 8; c++) s[c] = ss[c] = 0.f;
    if (pl < lanes)
        for (long long p = (long long)blockIdx.x * lanes + pl; p < HW; p += (long long)gridDim.x * lanes) {
            const Vec8 a = ld8(x + ((long long)n * HW + p) * C + cg * 8);
#pragma unroll
            for (int c = 0; c < 8; c++) { s[c] += a.v[c]; ss[c] += a.v[c] * a.v[c]; }
        }
#pragma unroll
    for (int c = 0; c < 8; c++) { red[threadIdx.x][c] = s[c]; red[threadIdx.x][8 + c] = ss[c]; }
    __syncthreads();
    for (int t = threadIdx.x; t < groups * 16; t += 256) {
        const int g = t / 16, j = t % 16;
        float acc = 0.f;
        for (int l = 0; l < lanes; l++) acc += red[l * groups + g][j];
        atomicAdd(stats + ((long long)n * C + g * 8 + (j & 7)) * 2 + (j >> 3), (double)acc);
    }
}

struct MeanRstd { float mean[8], rstd[8]; };
IPER_DEVINL MeanRstd mean_rstd(const double* stats, int n, int C, int cg, long long HW, float eps) {
    MeanRstd r;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const double s = stats[((long long)n * C + cg * 8 + c) * 2], ss = stats[((long long)n * C + cg * 8 + c) * 2 + 1];
        const double m = s / (double)HW;
        double var = ss / (double)HW - m * m;
        if (var < 0.0) var = 0.0;
        r.mean[c] = (float)m; r.rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    }
    return r;
}

// y = act( xh * (1 + gamma) + beta ),  xh = (x - mean) * rstd     (gamma / beta NULL: plain instance norm)
__global__ void __launch_bounds__(256) norm_apply_kernel(const __nv_bfloat16* __restrict__ x, const double* __restrict__ stats,
                                                         const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                                                         long long HW, int C, float eps, int act, float slope, __nv_bfloat16* __restrict__ y) {
    const int groups = C / 8, lanes = 256 / groups;
    const int cg = threadIdx.x % groups, pl = threadIdx.x / groups, n = blockIdx.y;
    if (pl >= lanes) return;
    const MeanRstd mr = mean_rstd(stats, n, C, cg, HW, eps);
    for (long long p = (long long)blockIdx.x * lanes + pl; p < HW; p += (long long)gridDim.x * lanes) {
        const long long off = ((long long)n * HW + p) * C + cg * 8;
        Vec8 a = ld8(x + off);
#pragma unroll
        for (int c = 0; c < 8; c++) a.v[c] = (a.v[c] - mr.mean[c]) * mr.rstd[c];
        if (gamma) {
            const Vec8 g = ld8(gamma + off), b = ld8(beta + off);
#pragma unroll
            for (int c = 0; c < 8; c++) a.v[c] = a.v[c] * (1.f + g.v[c]) + b.v[c];
        }
        if (act) {
#pragma unroll
            for (int c = 0; c < 8; c++) a.v[c] = a.v[c] > 0.f ? a.v[c] : a.v[c] * slope;
        }
        st8(y + off, a);
    }
}

// PASS 0: dgamma = g * xh, dbeta = g, sums of dxh = g (1 + gamma) and dxh * xh per (n, c)   (g = dout through the activation)
// PASS 1: dx = rstd * (dxh - mean(dxh) - xh * mean(dxh * xh))
template <int PASS>
__global__ void __launch_bounds__(256) norm_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                                                       const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ gamma,
                                                       const double* __restrict__ stats, double* __restrict__ sums, long long HW, int C,
                                                       float eps, int act, float slope, __nv_bfloat16* __restrict__ dgamma,
                                                       __nv_bfloat16* __restrict__ dbeta, __nv_bfloat16* __restrict__ dx) {
    __shared__ float red[PASS == 0 ? 256 : 1][17];
    const int groups = C / 8, lanes = 256 / groups;
    const int cg = threadIdx.x % groups, pl = threadIdx.x / groups, n = blockIdx.y;
    float s1[8], s2[8];
#pragma unroll
    for (int c = 0; c < 8; c++) s1[c] = s2[c] = 0.f;
    if (pl < lanes) {
        const MeanRstd mr = mean_rstd(stats, n, C, cg, HW, eps);
        float m1[8], m2[8];
        if (PASS == 1) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                m1[c] = (float)(sums[((long long)n * C + cg * 8 + c) * 2] / (double)HW);
                m2[c] = (float)(sums[((long long)n * C + cg * 8 + c) * 2 + 1] / (double)HW);
            }
        }
        for (long long p = (long long)blockIdx.x * lanes + pl; p < HW; p += (long long)gridDim.x * lanes) {
            const long long off = ((long long)n * HW + p) * C + cg * 8;
            Vec8 g = ld8(dout + off);
            Vec8 xh = ld8(x + off);
#pragma unroll
            for (int c = 0; c < 8; c++) xh.v[c] = (xh.v[c] - mr.mean[c]) * mr.rstd[c];
            if (act) {
                const Vec8 yy = ld8(y + off);
#pragma unroll
                for (int c = 0; c < 8; c++) g.v[c] = yy.v[c] > 0.f ? g.v[c] : g.v[c] * slope;
            }
            Vec8 dxh = g;
            if (gamma) {
                const Vec8 gm = ld8(gamma + off);
#pragma unroll
                for (int c = 0; c < 8; c++) dxh.v[c] = g.v[c] * (1.f + gm.v[c]);
            }
            if (PASS == 0) {
                if (gamma) {
                    Vec8 dg;
#pragma unroll
                    for (int c = 0; c < 8; c++) dg.v[c] = g.v[c] * xh.v[c];
                    st8(dgamma + off, dg);
                    st8(dbeta + off, g);
                }
#pragma unroll
                for (int c = 0; c < 8; c++) { s1[c] += dxh.v[c]; s2[c] += dxh.v[c] * xh.v[c]; }
            } else {
                Vec8 o;
#pragma unroll
                for (int c = 0; c < 8; c++) o.v[c] = mr.rstd[c] * (dxh.v[c] - m1[c] - xh.v[c] * m2[c]);
                st8(dx + off, o);
            }
        }
    }
    if (PASS == 0) {
#pragma unroll
        for (int c = 0; c < 8; c++) { red[threadIdx.x][c] = s1[c]; red[threadIdx.x][8 + c] = s2[c]; }
        __syncthreads();
        for (int t = threadIdx.x; t < groups * 16; t += 256) {
            const int g = t / 16, j = t % 16;
            float acc = 0.f;
            for (int l = 0; l < lanes; l++) acc += red[l * groups + g][j];
            atomicAdd(sums + ((long long)n * C + g * 8 + (j & 7)) * 2 + (j >> 3), (double)acc);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// (N, C, H, W) fp32 | bf16 with ANY element strides -> dense NHWC bf16 with the channels zero-padded to Cpad (a multiple of 8):
// the cast + F.pad + channels_last copy of the 1/3/4/6-channel ends of the step in one pass (16-byte stores)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pad_nhwc_kernel(const T* __restrict__ x, int N, int C, int H, int W, long long sn, long long sc,
                                                       long long sh, long long sw, int Cpad, __nv_bfloat16* __restrict__ out) {
    const int groups = Cpad / 8;
    const long long total = (long long)N * H * W * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % groups);
        const long long pix = i / groups;
        Vec8 v;
#pragma unroll
        for (int c = 0; c < 8; c++) v.v[c] = 0.f;
        if (cg * 8 < C) {
            const int xw = (int)(pix % W), yh = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
            const T* src = x + n * sn + yh * sh + xw * sw;
#pragma unroll
            for (int c = 0; c < 8; c++)
                if (cg * 8 + c < C) v.v[c] = (float)src[(long long)(cg * 8 + c) * sc];
        }
        st8(out + pix * Cpad + cg * 8, v);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// weight gradient of the THIN ends (5x5 heads 64 -> 3 / 1, 7x7 BGNet ends 4 -> 64 / 64 -> 3) on CUDA cores.  Through the tensor-core
// wgrad these layers stream 25-49 taps x 262 144 pixels of a 64-channel operand of which 1-4 channels are real (measured 410-470 us
// per launch, the slowest launches of the step); here the wide tensor's halo tile sits in shared memory once per tap group and
// every thread owns one wide channel: acc[tap][ct] += thin[p][ct] * wide[p + sign * (tap - pad)][cw] over the tile's pixels.
//   thin = dY (co <= 4):  dW[co, tap, ci] = sum_p dY[p, co] X[p + (tap - pad), ci]        sign +1, wide = X
//   thin = X  (ci <= 4):  dW[co, tap, ci] = sum_q dY[q - (tap - pad), co] X[q, ci]        sign -1, wide = dY
// grid (tiles [persistent], tap groups of <= 16, 64-channel chunks of the wide tensor); 256 threads = 64 channels x 4 pixel groups.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int THIN_TG = 16;

struct ThinArgs {
    const __nv_bfloat16* wide; const __nv_bfloat16* thin;
    int N, H, W, wide_pitch, thin_pitch, Ct, ks, pad, sign;
    float* dW; long long s_t, s_w, s_tap;
};

__global__ void __launch_bounds__(256) thin_wgrad_kernel(const ThinArgs a) {
    extern __shared__ __align__(16) uint8_t thin_smem[];
    const int P = max(a.pad, a.ks - 1 - a.pad), TW = 16 + 2 * P, TH = 8 + 2 * P;
    __nv_bfloat16* sw = reinterpret_cast<__nv_bfloat16*>(thin_smem);                       // [TH][TW][64]
    float* st = reinterpret_cast<float*>(thin_smem + (size_t)TH * TW * 128);                 // [128][4]
    float* sacc = st + 128 * 4;                                                              // [THIN_TG * 4][64]
    const int cw = threadIdx.x & 63, g = threadIdx.x >> 6;
    const int taps = a.ks * a.ks, tap0 = blockIdx.y * THIN_TG, nt = min(THIN_TG, taps - tap0), chunk = blockIdx.z;
    float acc[THIN_TG][4];
#pragma unroll
    for (int t = 0; t < THIN_TG; t++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[t][c] = 0.f;
    int toff[THIN_TG];                       // offset of tap t inside the halo tile, in pixels, relative to the thin pixel's halo position
#pragma unroll
    for (int t = 0; t < THIN_TG; t++) {
        const int tt = min(tap0 + t, taps - 1);
        toff[t] = (a.sign * (tt / a.ks - a.pad)) * TW + a.sign * (tt % a.ks - a.pad);
    }
    const int tiles_x = (a.W + 15) / 16, tiles_y = (a.H + 7) / 8, n_tiles = tiles_x * tiles_y * a.N;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int x0 = (tile % tiles_x) * 16, y0 = ((tile / tiles_x) % tiles_y) * 8, n = tile / (tiles_x * tiles_y);
        __syncthreads();                     // the previous tile's readers are done
        for (int i = threadIdx.x; i < TH * TW * 8; i += 256) {          // wide halo tile, 16 bytes per thread, zero outside the image
            const int c8 = i & 7, hp = i >> 3, hx = hp % TW, hy = hp / TW;
            const int gx = x0 - P + hx, gy = y0 - P + hy;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (gx >= 0 && gx < a.W && gy >= 0 && gy < a.H)
                v = __ldg(reinterpret_cast<const uint4*>(a.wide + (((size_t)n * a.H + gy) * a.W + gx) * a.wide_pitch + chunk * 64 + c8 * 8));
            *reinterpret_cast<uint4*>(sw + (size_t)hp * 64 + c8 * 8) = v;
        }
        if (threadIdx.x < 128) {                                          // thin tile: the first 4 channels of each pixel as floats
            const int px = threadIdx.x & 15, py = threadIdx.x >> 4, gx = x0 + px, gy = y0 + py;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gx < a.W && gy < a.H) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(a.thin + (((size_t)n * a.H + gy) * a.W + gx) * a.thin_pitch));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
                const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
                v = make_float4(f0.x, a.Ct > 1 ? f0.y : 0.f, a.Ct > 2 ? f1.x : 0.f, a.Ct > 3 ? f1.y : 0.f);
            }
            reinterpret_cast<float4*>(st)[threadIdx.x] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int q = 0; q < 32; q++) {                                    // this thread's 32 pixels: rows 2g, 2g + 1 of the tile
            const int py = 2 * g + (q >> 4), px = q & 15;
            const float4 th = reinterpret_cast<const float4*>(st)[py * 16 + px];
            const __nv_bfloat16* base = sw + ((size_t)(py + P) * TW + (px + P)) * 64 + cw;
#pragma unroll
            for (int t = 0; t < THIN_TG; t++) {
                if (t < nt) {
                    const float x = __bfloat162float(base[toff[t] * 64]);
                    acc[t][0] += th.x * x; acc[t][1] += th.y * x; acc[t][2] += th.z * x; acc[t][3] += th.w * x;
                }
            }
        }
    }
    // combine the four pixel groups in shared memory, then one atomic per (tap, thin channel, wide channel)
    __syncthreads();
    for (int i = threadIdx.x; i < THIN_TG * 4 * 64; i += 256) sacc[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int t = 0; t < THIN_TG; t++)
#pragma unroll
        for (int c = 0; c < 4; c++) atomicAdd(&sacc[(t * 4 + c) * 64 + cw], acc[t][c]);
    __syncthreads();
    for (int i = threadIdx.x; i < THIN_TG * 4 * 64; i += 256) {
        const int w = i & 63, c = (i >> 6) & 3, t = i >> 8;
        if (t < nt && c < a.Ct)
            atomicAdd(a.dW + (long long)c * a.s_t + (long long)(tap0 + t) * a.s_tap + (long long)(chunk * 64 + w) * a.s_w, sacc[i]);
    }
}

static int grid_for(long long work_items, int per_block) {
    long long b = (work_items + per_block - 1) / per_block;
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    return (int)b;
}
static bool ok_channels(int C) { return C >= 64 && C % 8 == 0 && C <= 2048 && (256 % (C / 8) == 0 || C / 8 > 32); }

}  // namespace iper

using namespace iper;

#define IPER_A16(p) (((uintptr_t)(p) & 15) == 0)

extern "C" int iper_warp_bf16(const void* src_nhwc, const float* T, int M, int h, int w, int C, void* out_nhwc, iper_stream_t stream) {
    IPER_REQUIRE(src_nhwc && T && out_nhwc && M > 0 && h > 0 && w > 0 && C > 0 && C % 8 == 0 && IPER_A16(src_nhwc) && IPER_A16(out_nhwc) && IPER_A16(T),
                 "iper_warp_bf16: null / unaligned pointer or C %% 8 != 0 (C = %d)", C);
    const long long total = (long long)M * h * w * (C / 8);
    warp_bf16_kernel<false><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(src_nhwc), T, M, h, w, C,
                                                                                  reinterpret_cast<__nv_bfloat16*>(out_nhwc), nullptr);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_warp_bwd_bf16(const void* dout_nhwc, const float* T, int M, int h, int w, int C, float* dsrc_f32, iper_stream_t stream) {
    IPER_REQUIRE(dout_nhwc && T && dsrc_f32 && M > 0 && h > 0 && w > 0 && C > 0 && C % 8 == 0 && IPER_A16(dout_nhwc) && IPER_A16(dsrc_f32) && IPER_A16(T),
                 "iper_warp_bwd_bf16: null / unaligned pointer or C %% 8 != 0 (C = %d)", C);
    cudaStream_t st = (cudaStream_t)stream;
    IPER_CHECK_CUDA(cudaMemsetAsync(dsrc_f32, 0, sizeof(float) * (size_t)M * h * w * C, st));
    const long long total = (long long)M * h * w * (C / 8);
    warp_bf16_kernel<true><<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dout_nhwc), T, M, h, w, C, nullptr, dsrc_f32);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_att_combine_bf16(const void* k, const void* v, const void* q, int bs, int ns, long long HW, int C, void* a, float* alpha,
                                     iper_stream_t stream) {
    IPER_REQUIRE(k && v && q && a && alpha && IPER_A16(k) && IPER_A16(v) && IPER_A16(q) && IPER_A16(a), "iper_att_combine_bf16: null / unaligned pointer");
    IPER_REQUIRE(bs > 0 && ns > 0 && ns <= ATT_MAX_NS && HW > 0 && (C == 64 || C == 128 || C == 256) && (bs * HW * (C / 8)) % 32 == 0,
                 "iper_att_combine_bf16: needs ns <= %d, C in {64,128,256}, bs*HW*C/8 %% 32 == 0 (got ns %d, C %d, HW %lld)", ATT_MAX_NS, ns, C, HW);
    const long long total = (long long)bs * HW * (C / 8);
    att_combine_kernel<false><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(k), reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<const __nv_bfloat16*>(q), bs, ns, HW, C,
        1.0f / sqrtf((float)C), reinterpret_cast<__nv_bfloat16*>(a), alpha, nullptr, nullptr, nullptr, nullptr);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_att_combine_bwd_bf16(const void* da, const void* k, const void* v, const void* q, const float* alpha, int bs, int ns,
                                         long long HW, int C, void* dk, void* dv, void* dq, iper_stream_t stream) {
    IPER_REQUIRE(da && k && v && q && alpha && dk && dv && dq && IPER_A16(da) && IPER_A16(k) && IPER_A16(v) && IPER_A16(q) && IPER_A16(dk) &&
                 IPER_A16(dv) && IPER_A16(dq), "iper_att_combine_bwd_bf16: null / unaligned pointer");
    IPER_REQUIRE(bs > 0 && ns > 0 && ns <= ATT_MAX_NS && HW > 0 && (C == 64 || C == 128 || C == 256) && (bs * HW * (C / 8)) % 32 == 0,
                 "iper_att_combine_bwd_bf16: needs ns <= %d, C in {64,128,256}, bs*HW*C/8 %% 32 == 0 (got ns %d, C %d, HW %lld)", ATT_MAX_NS, ns, C, HW);
    const long long total = (long long)bs * HW * (C / 8);
    att_combine_kernel<true><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(k), reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<const __nv_bfloat16*>(q), bs, ns, HW, C,
        1.0f / sqrtf((float)C), nullptr, const_cast<float*>(alpha), reinterpret_cast<const __nv_bfloat16*>(da), reinterpret_cast<__nv_bfloat16*>(dk),
        reinterpret_cast<__nv_bfloat16*>(dv), reinterpret_cast<__nv_bfloat16*>(dq));
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_norm_stats_bf16(const void* x_nhwc, int N, long long HW, int C, double* stats, iper_stream_t stream) {
    IPER_REQUIRE(x_nhwc && stats && IPER_A16(x_nhwc) && N > 0 && HW > 0 && ok_channels(C) && C <= 2048 && 256 % (C / 8) == 0,
                 "iper_norm_stats_bf16: null / unaligned pointer or unsupported C %d (64..2048, C/8 a divisor of 256)", C);
    cudaStream_t st = (cudaStream_t)stream;
    IPER_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)N * C, st));
    const int lanes = 256 / (C / 8);
    dim3 grid(grid_for(HW, lanes * 16) , N);
    norm_stats_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x_nhwc), HW, C, stats);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_norm_apply_bf16(const void* x_nhwc, const double* stats, const void* gamma, const void* beta, int N, long long HW, int C,
                                    float eps, int act, float slope, void* y_nhwc, iper_stream_t stream) {
    IPER_REQUIRE(x_nhwc && stats && y_nhwc && IPER_A16(x_nhwc) && IPER_A16(y_nhwc) && IPER_A16(gamma) && IPER_A16(beta) && N > 0 && HW > 0 &&
                 ok_channels(C) && 256 % (C / 8) == 0 && ((gamma == nullptr) == (beta == nullptr)),
                 "iper_norm_apply_bf16: null / unaligned pointer, gamma without beta, or unsupported C %d", C);
    const int lanes = 256 / (C / 8);
    dim3 grid(grid_for(HW, lanes * 8), N);
    norm_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x_nhwc), stats,
                                                             reinterpret_cast<const __nv_bfloat16*>(gamma), reinterpret_cast<const __nv_bfloat16*>(beta),
                                                             HW, C, eps, act, slope, reinterpret_cast<__nv_bfloat16*>(y_nhwc));
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_norm_bwd_bf16(const void* dout, const void* x, const void* y, const void* gamma, const double* stats, int N, long long HW,
                                  int C, float eps, int act, float slope, double* sums_ws, void* dgamma, void* dbeta, void* dx,
                                  iper_stream_t stream) {
    IPER_REQUIRE(dout && x && stats && sums_ws && dx && IPER_A16(dout) && IPER_A16(x) && IPER_A16(y) && IPER_A16(gamma) && IPER_A16(dgamma) &&
                 IPER_A16(dbeta) && IPER_A16(dx), "iper_norm_bwd_bf16: null / unaligned pointer");
    IPER_REQUIRE(N > 0 && HW > 0 && ok_channels(C) && 256 % (C / 8) == 0 && (!act || y) && (!gamma || (dgamma && dbeta)),
                 "iper_norm_bwd_bf16: unsupported C %d, activation without the saved output, or gamma without dgamma / dbeta", C);
    cudaStream_t st = (cudaStream_t)stream;
    IPER_CHECK_CUDA(cudaMemsetAsync(sums_ws, 0, sizeof(double) * 2 * (size_t)N * C, st));
    const int lanes = 256 / (C / 8);
    dim3 grid(grid_for(HW, lanes * 8), N);
    auto B = [](const void* p) { return reinterpret_cast<const __nv_bfloat16*>(p); };
    auto Bm = [](void* p) { return reinterpret_cast<__nv_bfloat16*>(p); };
    norm_bwd_kernel<0><<<grid, 256, 0, st>>>(B(dout), B(x), B(y), B(gamma), stats, sums_ws, HW, C, eps, act, slope, Bm(dgamma), Bm(dbeta), nullptr);
    norm_bwd_kernel<1><<<grid, 256, 0, st>>>(B(dout), B(x), B(y), B(gamma), stats, sums_ws, HW, C, eps, act, slope, nullptr, nullptr, Bm(dx));
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_pad_nhwc_bf16(const void* x, int is_bf16, int N, int C, int H, int W, long long stride_n, long long stride_c,
                                  long long stride_h, long long stride_w, int Cpad, void* out_nhwc, iper_stream_t stream) {
    IPER_REQUIRE(x && out_nhwc && IPER_A16(out_nhwc) && N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C && Cpad % 8 == 0,
                 "iper_pad_nhwc_bf16: null / unaligned pointer or Cpad %d not a multiple of 8 >= C %d", Cpad, C);
    const long long total = (long long)N * H * W * (Cpad / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (is_bf16)
        pad_nhwc_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), N, C, H, W, stride_n, stride_c,
                                                                            stride_h, stride_w, Cpad, reinterpret_cast<__nv_bfloat16*>(out_nhwc));
    else
        pad_nhwc_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const float*>(x), N, C, H, W, stride_n, stride_c, stride_h,
                                                                    stride_w, Cpad, reinterpret_cast<__nv_bfloat16*>(out_nhwc));
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_thin_wgrad_bf16(const void* wide_nhwc, const void* thin_nhwc, int N, int H, int W, int Cw, int thin_pitch, int Ct,
                                    int ksize, int pad, int sign, float* dW, long long stride_thin, long long stride_wide,
                                    long long stride_tap, iper_stream_t stream) {
    IPER_REQUIRE(wide_nhwc && thin_nhwc && dW && IPER_A16(wide_nhwc) && (((uintptr_t)thin_nhwc) & 7) == 0, "iper_thin_wgrad_bf16: null / unaligned pointer");
    IPER_REQUIRE(N > 0 && H > 0 && W > 0 && Cw > 0 && Cw % 64 == 0 && Ct >= 1 && Ct <= 4 && thin_pitch >= 4 && thin_pitch % 4 == 0 &&
                 ksize >= 1 && ksize <= 7 && pad >= 0 && pad < ksize && (sign == 1 || sign == -1),
                 "iper_thin_wgrad_bf16: needs Cw %% 64 == 0, 1 <= Ct <= 4, thin_pitch %% 4 == 0, k <= 7, sign +-1 (got Cw %d Ct %d k %d)", Cw, Ct, ksize);
    const int P = pad > ksize - 1 - pad ? pad : ksize - 1 - pad;
    const int smem = (8 + 2 * P) * (16 + 2 * P) * 128 + 128 * 16 + THIN_TG * 4 * 64 * 4;
    static int have[64] = {};
    int dev = 0;
    IPER_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && have[dev] < smem) {
        IPER_CHECK_CUDA(cudaFuncSetAttribute(thin_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        have[dev] = smem;
    }
    ThinArgs a;
    a.wide = reinterpret_cast<const __nv_bfloat16*>(wide_nhwc); a.thin = reinterpret_cast<const __nv_bfloat16*>(thin_nhwc);
    a.N = N; a.H = H; a.W = W; a.wide_pitch = Cw; a.thin_pitch = thin_pitch; a.Ct = Ct; a.ks = ksize; a.pad = pad; a.sign = sign;
    a.dW = dW; a.s_t = stride_thin; a.s_w = stride_wide; a.s_tap = stride_tap;
    const int n_tiles = ((W + 15) / 16) * ((H + 7) / 8) * N, groups = (ksize * ksize + THIN_TG - 1) / THIN_TG, chunks = Cw / 64;
    int bx = (148 * 2) / (groups * chunks);
    if (bx < 1) bx = 1;
    if (bx > n_tiles) bx = n_tiles;
    thin_wgrad_kernel<<<dim3(bx, groups, chunks), 256, smem, (cudaStream_t)stream>>>(a);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}
