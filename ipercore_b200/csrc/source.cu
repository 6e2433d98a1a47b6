// One-time-per-source kernels of Imitator.source_setup, sm_100a.
//
// Replaces, in FlowComposition.process_source (iPERCore/models/flowcomposition.py:452-512):
//   * CannyFilter.forward(confidant_sil, 0.1, 0.9, True)          iPERCore/tools/utils/morphology/canny_ops.py:129-192
//   * make_morph_image / cal_top_k_ids / morph_image               flowcomposition.py:264-386  (O(n1*n2) distance matrix)
//   * make_uv_img                                                   flowcomposition.py:87-137   (2x cal_bc_transform, 2x grid_sample, merge)
// All of it is HBM/latency-trivial work (a few MB once per source set); the point is to remove the reference's
// materialised (n1, n2, 2) int64 distance tensors, its `nonzero()` host synchronisations and ~40 tiny launches.
#include "common.cuh"
#include "iper_b200.h"

namespace iper {

// ------------------------------------------------------------------------------------------------------------
// Canny on a 1-channel map (canny_ops.py:129-192 with C = 1).  Every stage is a 3x3 cross-correlation with ZERO padding
// of its own input (nn.Conv2d(padding=1)), so a stage sees zeros outside the image even where the previous stage would
// have been non-zero there.  Tap order is row-major (dy, dx), plain multiply-add chains.
// ------------------------------------------------------------------------------------------------------------
struct CannyW {
    float g[9];        // gaussian_filter.weight
    float sx[9];       // sobel_filter_x.weight   (sobel_filter_y = transpose)
    float d[8][9];     // directional_filter.weight
    float hyst;        // hysteresis.weight (all taps equal: 1.25)
};

// stage 1+2: blurred -> (grad_x, grad_y) -> magnitude, orientation index.  One thread per pixel, 5x5 input footprint.
__global__ void __launch_bounds__(256) canny_grad_kernel(const float* __restrict__ img, int H, int W, CannyW w,
                                                         float* __restrict__ mag, signed char* __restrict__ pidx) {
    const int n = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const float* src = img + (size_t)n * H * W;
    auto px = [&](int yy, int xx) -> float { return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(src + (size_t)yy * W + xx) : 0.f; };
    float gx = 0.f, gy = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++) {
            const int by = y + dy - 1, bx = x + dx - 1;
            float b = 0.f;                                 // blurred outside the image is the zero padding of the sobel conv
            if (by >= 0 && by < H && bx >= 0 && bx < W) {
#pragma unroll
                for (int ey = 0; ey < 3; ey++)
#pragma unroll
                    for (int ex = 0; ex < 3; ex++) b = __fadd_rn(b, __fmul_rn(w.g[ey * 3 + ex], px(by + ey - 1, bx + ex - 1)));
            }
            gx = __fadd_rn(gx, __fmul_rn(w.sx[dy * 3 + dx], b));
            gy = __fadd_rn(gy, __fmul_rn(w.sx[dx * 3 + dy], b));
        }
    const float m = sqrtf(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
    // orientation: atan(gy/gx) * (360/pi) + 180, rounded to a multiple of 45; index = (ori/45) % 8 (NaN -> no orientation)
    float ori = __fadd_rn(__fmul_rn(atanf(__fdiv_rn(gy, gx)), (float)(360.0 / 3.14159265358979323846)), 180.f);
    ori = __fmul_rn(rintf(__fdiv_rn(ori, 45.f)), 45.f);
    const float q = __fdiv_rn(ori, 45.f);
    signed char pi = -1;
    if (q == q) pi = (signed char)(((int)q) & 7);
    mag[(size_t)n * H * W + (size_t)y * W + x] = m;
    pidx[(size_t)n * H * W + (size_t)y * W + x] = pi;
}

// stage 3+4: non-maximum suppression along the orientation, double threshold -> {0, 0.5, 1}
__global__ void __launch_bounds__(256) canny_thin_kernel(const float* __restrict__ mag, const signed char* __restrict__ pidx,
                                                         int H, int W, CannyW w, float low, float high,
                                                         float* __restrict__ tri) {
    const int n = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const float* src = mag + (size_t)n * H * W;
    float nb[9];
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++) {
            const int yy = y + dy - 1, xx = x + dx - 1;
            nb[dy * 3 + dx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(src + (size_t)yy * W + xx) : 0.f;
        }
    float thin = nb[4];
    const int pi = pidx[(size_t)n * H * W + (size_t)y * W + x];
    if (pi >= 0) {
        const int pos = pi & 3, neg = pos + 4;           // the pair (pos_i, pos_i + 4) this pixel is oriented along
        float dp = 0.f, dn = 0.f;
#pragma unroll
        for (int t = 0; t < 9; t++) {
            dp = __fadd_rn(dp, __fmul_rn(w.d[pos][t], nb[t]));
            dn = __fadd_rn(dn, __fmul_rn(w.d[neg][t], nb[t]));
        }
        if (!(fminf(dp, dn) > 0.f)) thin = 0.f;          // not a local maximum along its direction
    }
    const float lo = thin > low ? 0.5f : 0.f, hi = thin > high ? 0.5f : 0.f;
    tri[(size_t)n * H * W + (size_t)y * W + x] = lo + hi;
}

// stage 5: hysteresis — weak pixels (0.5) next to enough strong mass become edges; output {0, 1}
__global__ void __launch_bounds__(256) canny_hyst_kernel(const float* __restrict__ tri, int H, int W, float hw,
                                                         float* __restrict__ edges) {
    const int n = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const float* src = tri + (size_t)n * H * W;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++) {
            const int yy = y + dy - 1, xx = x + dx - 1;
            const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(src + (size_t)yy * W + xx) : 0.f;
            acc = __fadd_rn(acc, __fmul_rn(hw, v));
        }
    const float c = src[(size_t)y * W + x];
    const bool high = c == 1.0f, weak = c == 0.5f;
    edges[(size_t)n * H * W + (size_t)y * W + x] = (high || (weak && acc > 1.0f)) ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------------------------------
// make_morph_image (flowcomposition.py:335-386): every "uncertain" pixel (outpad_sil * (1 - confidant_sil) != 0) takes the
// colour  sum_k w_k * src[nn_k]  of its top_k = 3 nearest boundary (thin-edge) pixels with w = d_k / sum(d)  (squared pixel
// distances, flowcomposition.py:264-293 — farther points weigh more, as upstream), all other pixels src * confidant_sil.
// Pass 1 compacts the boundary pixels of each image; pass 2 streams them through shared memory, each thread keeping the
// three smallest (distance, row-major index) pairs — the tie-break torch.topk leaves unspecified is fixed to the lowest
// index.  Images with fewer than top_k boundary pixels keep src * confidant_sil everywhere (torch.topk would raise).
// ------------------------------------------------------------------------------------------------------------
__global__ void boundary_compact_kernel(const float* __restrict__ edges, int HW, int* __restrict__ count, int* __restrict__ list) {
    const int n = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
        if (edges[(size_t)n * HW + i] != 0.f) list[(size_t)n * HW + atomicAdd(count + n, 1)] = i;
}

constexpr int NN_TILE = 1024;
__global__ void __launch_bounds__(256) morph_image_kernel(const float* __restrict__ src, const float* __restrict__ conf,
                                                          const float* __restrict__ outpad, const int* __restrict__ count,
                                                          const int* __restrict__ list, int H, int W, float* __restrict__ out) {
    __shared__ int s_pts[NN_TILE];
    const int n = blockIdx.y, HW = H * W;
    const int p = blockIdx.x * 256 + threadIdx.x;
    const bool in = p < HW;
    const float cf = in ? conf[(size_t)n * HW + p] : 1.f;
    const float unc = in ? outpad[(size_t)n * HW + p] * (1.f - cf) : 0.f;
    const int nb = count[n];
    const bool work = in && unc != 0.f && nb >= 3;
    const int py = p / W, px = p - py * W;
    long long d0 = 0x7fffffffffffffffLL, d1 = d0, d2 = d0;       // (distance << 32 | index): smaller = nearer, then lower index
    if (__syncthreads_or(work)) {
        for (int base = 0; base < nb; base += NN_TILE) {
            const int m = min(NN_TILE, nb - base);
            __syncthreads();
            for (int i = threadIdx.x; i < m; i += 256) s_pts[i] = list[(size_t)n * HW + base + i];
            __syncthreads();
            if (work) {
                for (int i = 0; i < m; i++) {
                    const int q = s_pts[i];
                    const int qy = q / W, qx = q - qy * W;
                    const int dy = py - qy, dx = px - qx;
                    const long long key = ((long long)(dy * dy + dx * dx) << 32) | (unsigned)q;
                    if (key < d2) {
                        if (key < d1) {
                            d2 = d1;
                            if (key < d0) { d1 = d0; d0 = key; } else d1 = key;
                        } else d2 = key;
                    }
                }
            }
        }
    }
    if (!in) return;
    const float* s = src + (size_t)n * 3 * HW;
    float* o = out + (size_t)n * 3 * HW;
    if (work) {
        const float v0 = (float)(d0 >> 32), v1 = (float)(d1 >> 32), v2 = (float)(d2 >> 32);
        const float sum = v0 + v1 + v2;
        const float w0 = v0 / sum, w1 = v1 / sum, w2 = v2 / sum;
        const int q0 = (int)(d0 & 0xffffffffLL), q1 = (int)(d1 & 0xffffffffLL), q2 = (int)(d2 & 0xffffffffLL);
#pragma unroll
        for (int c = 0; c < 3; c++)
            o[(size_t)c * HW + p] = s[(size_t)c * HW + q0] * w0 + s[(size_t)c * HW + q1] * w1 + s[(size_t)c * HW + q2] * w2;
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) o[(size_t)c * HW + p] = s[(size_t)c * HW + p] * cf;
    }
}

// ------------------------------------------------------------------------------------------------------------
// make_uv_img (flowcomposition.py:87-137).  Pass A, per source n and UV pixel: T = sum_k uv_wim[p,k] * f2pts[n, uv_fim[p], k]
// (cal_bc_transform, -2 on uncovered UV pixels) for the full and the visible-only corner sets, then
//   src_warp[n] = grid_sample(src_img[n], T_full)     vis_warp[n] = grid_sample(ones, T_vis)  (= sum of in-range bilinear weights)
// The reference dilates vis_warp with a 13x13 box (iper_morph) before pass B merges:
//   vis_sum = sum_{s>=1} vis[s];  temp = sum_{s>=1} src_warp[s]*vis[s] / (vis_sum + 1e-5)
//   front_invisible = (1 - vis[0]) * (vis_sum >= 1);  uv = src_warp[0] * (1 - front_invisible) + temp * front_invisible
// ------------------------------------------------------------------------------------------------------------
IPER_DEVINL void bilinear4(float gx, float gy, int H, int W, int (&off)[4], float (&wt)[4]) {
    const float ix = ((gx + 1.f) * W - 1.f) * 0.5f, iy = ((gy + 1.f) * H - 1.f) * 0.5f;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const bool sane = fabsf(ix) < 1e8f && fabsf(iy) < 1e8f;
    const int x0 = sane ? (int)fx0 : -10, y0 = sane ? (int)fy0 : -10, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix, wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
    const int xs[4] = {x0, x1, x0, x1}, ys[4] = {y0, y0, y1, y1};
    const float ws[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool ok = xs[i] >= 0 && xs[i] < W && ys[i] >= 0 && ys[i] < H;
        off[i] = ok ? ys[i] * W + xs[i] : -1;
        wt[i] = ok ? ws[i] : 0.f;
    }
}

__global__ void __launch_bounds__(256) uv_warp_kernel(const float* __restrict__ src, const float* __restrict__ f2pts,
                                                      const float* __restrict__ vis_f2pts, const int32_t* __restrict__ uv_fim,
                                                      const float* __restrict__ uv_wim, int nf, int H, int W,
                                                      float* __restrict__ src_warp, float* __restrict__ vis_warp) {
    const int n = blockIdx.y, HW = H * W;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= HW) return;
    const int fn = __ldg(uv_fim + p);
    float tx = -2.f, ty = -2.f, vx = -2.f, vy = -2.f;
    if (fn >= 0) {
        const float w0 = __ldg(uv_wim + 3 * (size_t)p), w1 = __ldg(uv_wim + 3 * (size_t)p + 1), w2 = __ldg(uv_wim + 3 * (size_t)p + 2);
        const float* a = f2pts + ((size_t)n * nf + fn) * 6;
        const float* b = vis_f2pts + ((size_t)n * nf + fn) * 6;
        tx = __fadd_rn(__fadd_rn(__fmul_rn(a[0], w0), __fmul_rn(a[2], w1)), __fmul_rn(a[4], w2));
        ty = __fadd_rn(__fadd_rn(__fmul_rn(a[1], w0), __fmul_rn(a[3], w1)), __fmul_rn(a[5], w2));
        vx = __fadd_rn(__fadd_rn(__fmul_rn(b[0], w0), __fmul_rn(b[2], w1)), __fmul_rn(b[4], w2));
        vy = __fadd_rn(__fadd_rn(__fmul_rn(b[1], w0), __fmul_rn(b[3], w1)), __fmul_rn(b[5], w2));
    }
    int off[4]; float wt[4];
    bilinear4(tx, ty, H, W, off, wt);
    const float* s = src + (size_t)n * 3 * HW;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (off[i] >= 0) acc += s[(size_t)c * HW + off[i]] * wt[i];
        src_warp[((size_t)n * 3 + c) * HW + p] = acc;
    }
    bilinear4(vx, vy, H, W, off, wt);
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (off[i] >= 0) v += wt[i];
    vis_warp[(size_t)n * HW + p] = v;
}

__global__ void __launch_bounds__(256) uv_merge_kernel(const float* __restrict__ src_warp, const float* __restrict__ vis, int bs,
                                                       int ns, int HW, float* __restrict__ uv) {
    const size_t total = (size_t)bs * HW;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const size_t b = i / HW, p = i % HW;
        float vis_sum = 0.f, t[3] = {0.f, 0.f, 0.f};
        for (int s = 1; s < ns; s++) {
            const float v = vis[(b * ns + s) * HW + p];
            vis_sum += v;
#pragma unroll
            for (int c = 0; c < 3; c++) t[c] += src_warp[((b * ns + s) * 3 + c) * HW + p] * v;
        }
        const float front_invisible = (1.f - vis[(b * ns) * HW + p]) * (vis_sum >= 1.f ? 1.f : 0.f);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float temp = t[c] / (vis_sum + 1e-5f);
            uv[(b * 3 + c) * HW + p] = src_warp[((b * ns) * 3 + c) * HW + p] * (1.f - front_invisible) + temp * front_invisible;
        }
    }
}

}  // namespace iper

using namespace iper;

extern "C" int iper_canny_edges(const float* img, int N, int H, int W, const float* gaussian_w, const float* sobel_x_w,
                                const float* directional_w, float hysteresis_w, float low, float high, float* mag_ws,
                                int8_t* ori_ws, float* tri_ws, float* edges, iper_stream_t stream) {
    IPER_REQUIRE(N >= 0 && H > 0 && W > 0, "iper_canny_edges: bad sizes");
    if (N == 0) return 0;
    IPER_REQUIRE(img && gaussian_w && sobel_x_w && directional_w && mag_ws && ori_ws && tri_ws && edges, "iper_canny_edges: null pointer");
    CannyW w;
    cudaStream_t st = (cudaStream_t)stream;
    // the 9 + 9 + 72 filter taps are HOST pointers (read here, passed by value): they are module constants of CannyFilter
    for (int i = 0; i < 9; i++) { w.g[i] = gaussian_w[i]; w.sx[i] = sobel_x_w[i]; }
    for (int k = 0; k < 8; k++)
        for (int i = 0; i < 9; i++) w.d[k][i] = directional_w[k * 9 + i];
    w.hyst = hysteresis_w;
    dim3 grid((W + 31) / 32, (H + 7) / 8, N);
    canny_grad_kernel<<<grid, 256, 0, st>>>(img, H, W, w, mag_ws, reinterpret_cast<signed char*>(ori_ws));
    IPER_CHECK_CUDA(cudaGetLastError());
    canny_thin_kernel<<<grid, 256, 0, st>>>(mag_ws, reinterpret_cast<const signed char*>(ori_ws), H, W, w, low, high, tri_ws);
    IPER_CHECK_CUDA(cudaGetLastError());
    canny_hyst_kernel<<<grid, 256, 0, st>>>(tri_ws, H, W, hysteresis_w, edges);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_morph_image(const float* src_img, const float* confidant_sil, const float* outpad_sil, const float* edges,
                                int N, int H, int W, int32_t* count_ws, int32_t* list_ws, float* out, iper_stream_t stream) {
    IPER_REQUIRE(N >= 0 && H > 0 && W > 0 && (long long)H * W < (1LL << 31), "iper_morph_image: bad sizes");
    if (N == 0) return 0;
    IPER_REQUIRE(src_img && confidant_sil && outpad_sil && edges && count_ws && list_ws && out, "iper_morph_image: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    IPER_CHECK_CUDA(cudaMemsetAsync(count_ws, 0, sizeof(int32_t) * N, st));
    boundary_compact_kernel<<<dim3(min((HW + 255) / 256, 256), N), 256, 0, st>>>(edges, HW, count_ws, list_ws);
    IPER_CHECK_CUDA(cudaGetLastError());
    morph_image_kernel<<<dim3((HW + 255) / 256, N), 256, 0, st>>>(src_img, confidant_sil, outpad_sil, count_ws, list_ws, H, W, out);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_uv_warp(const float* src_img, const float* f2pts, const float* vis_f2pts, const int32_t* uv_fim,
                            const float* uv_wim, int N, int nf, int H, int W, float* src_warp, float* vis_warp,
                            iper_stream_t stream) {
    IPER_REQUIRE(N >= 0 && nf > 0 && H > 0 && W > 0, "iper_uv_warp: bad sizes");
    if (N == 0) return 0;
    IPER_REQUIRE(src_img && f2pts && vis_f2pts && uv_fim && uv_wim && src_warp && vis_warp, "iper_uv_warp: null pointer");
    uv_warp_kernel<<<dim3((H * W + 255) / 256, N), 256, 0, (cudaStream_t)stream>>>(src_img, f2pts, vis_f2pts, uv_fim, uv_wim, nf, H, W,
                                                                                    src_warp, vis_warp);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_uv_merge(const float* src_warp, const float* vis_dilated, int bs, int ns, int H, int W, float* uv_img,
                             iper_stream_t stream) {
    IPER_REQUIRE(bs >= 0 && ns >= 1 && H > 0 && W > 0, "iper_uv_merge: bad sizes");
    if (bs == 0) return 0;
    IPER_REQUIRE(src_warp && vis_dilated && uv_img, "iper_uv_merge: null pointer");
    const size_t total = (size_t)bs * H * W;
    uv_merge_kernel<<<(unsigned)min((total + 255) / 256, (size_t)148 * 8), 256, 0, (cudaStream_t)stream>>>(src_warp, vis_dilated, bs, ns,
                                                                                                           H * W, uv_img);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}
