// SMPL mesh rasteriser + per-frame geometry (rows a1-a8 of SURVEY.md §8a), sm_100a.
//
// Replaces, for the hot path:
//   * nr.rasterize_face_index_map_and_weight_map   (reference call sites nmr.py:337,356; third-party CUDA ext)
//   * SMPLRenderer.render_fim_wim glue              (nmr.py:319-342: projection, y flip, look_at, vertices_to_faces)
//   * SMPLRenderer.encode_fim                       (nmr.py:390-401)
//   * SMPLRenderer.cal_bc_transform                 (nmr.py:713-757)
//   * FlowComposition.make_tsf_inputs / make_trans_flow (flowcomposition.py:206-248, 514-582)
//
// Design (B200), two launches per batch of frames:
//   raster_setup_kernel  one thread per (frame, face): project the three vertices (a1), write the face's corners
//                        (B,nf,3,3) and f2pts (a3), cull back faces, and BIN the face: its conservative pixel bounding box
//                        (grown by one pixel) becomes a packed tile range (4 x 8 bits) — done once per face per frame.
//   raster_kernel        one CTA per 64x32 pixel tile per frame: scans the 4-byte bins (coalesced, L2-resident) into a
//                        compact shared list of the faces touching the tile, then each WARP takes binned faces and
//                        scan-converts them with its lanes spread over the pixels of (bbox ∩ tile) — no per-face divergence —
//                        straight into a shared memory z-buffer with a 64-bit atomicMin on (depth bits << 32 | face index) — which is exactly the upstream
// rule "strictly smaller depth wins, lowest face index wins a tie".  A resolve pass then recomputes the winner's
// barycentric weights (same float sequence, hence identical bits) and writes fim / wim / cond / flow / sampled
// UV image with fully coalesced stores.  Work is proportional to covered area, not faces x pixels
// (upstream: 262 144 px x 13 776 faces per 512^2 frame); 25 KB of shared memory per CTA.
//
// Parity contract: every float operation below that feeds the inside test, the weights or the depth is an
// explicitly rounded binary32 op (__fmul_rn/__fadd_rn/__fdiv_rn — never contracted into FMA) in the same order
// as oracle/raster_ref.c, including the two places where upstream's double literals promote to double.
#include <cstdlib>

#include "common.cuh"
#include "iper_b200.h"

namespace iper {

constexpr int TILE_W = 64;
constexpr int TILE_H = 32;
constexpr int RASTER_THREADS = 256;
constexpr unsigned long long ZBUF_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr int BIN_CHUNK = 4096;            // faces binned per pass (compact list of 16-bit offsets)
constexpr int RASTER_ZBUF_BYTES = TILE_W * TILE_H * 8;
constexpr int RASTER_FIXED_SMEM = RASTER_ZBUF_BYTES + BIN_CHUNK * 2 + (TILE_W + TILE_H) * 4 + 16;

struct FaceGeom {
    float v[9];  // x0 y0 z0 x1 y1 z1 x2 y2 z2 (NDC, +y up after the reference's flip)
};

IPER_DEVINL bool is_backface(const float* f) {
    return __fmul_rn(__fsub_rn(f[7], f[1]), __fsub_rn(f[3], f[0])) <
           __fmul_rn(__fsub_rn(f[4], f[1]), __fsub_rn(f[6], f[0]));
}

// Rounding model of the contraction-sensitive expressions (see oracle/raster_ref.c): `fma` = 0 rounds every product
// (source-level IEEE semantics, the default), `fma` = 1 fuses exactly where nvcc -fmad=true would contract upstream's
// source (the oracle's -DORACLE_FMA variant).  Selected per process by iper_raster_set_contraction / IPER_RASTER_FMA.
IPER_DEVINL float mul_add(float a, float b, float c, int fma) {            // a*b + c
    return fma ? __fmaf_rn(a, b, c) : __fadd_rn(__fmul_rn(a, b), c);
}
IPER_DEVINL float mul_sub2(float a, float b, float c, float d, int fma) {  // a*b - c*d
    return fma ? __fmaf_rn(a, b, -__fmul_rn(c, d)) : __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}

// upstream kernel 1: pixel-space corners and the adjugate/determinant inverse
IPER_DEVINL void face_inverse(const float* f, int is, float* inv, int fma) {
    float p[3][2];
    const float fis = (float)is;
#pragma unroll
    for (int n = 0; n < 3; n++)
#pragma unroll
        for (int d = 0; d < 2; d++) {
            float t = mul_add(f[3 * n + d], fis, fis, fma);
            t = __fsub_rn(t, 1.0f);
            p[n][d] = __fmul_rn(0.5f, t);  // exact
        }
    float a[9];
    a[0] = __fsub_rn(p[1][1], p[2][1]);
    a[1] = __fsub_rn(p[2][0], p[1][0]);
    a[2] = mul_sub2(p[1][0], p[2][1], p[2][0], p[1][1], fma);
    a[3] = __fsub_rn(p[2][1], p[0][1]);
    a[4] = __fsub_rn(p[0][0], p[2][0]);
    a[5] = mul_sub2(p[2][0], p[0][1], p[0][0], p[2][1], fma);
    a[6] = __fsub_rn(p[0][1], p[1][1]);
    a[7] = __fsub_rn(p[1][0], p[0][0]);
    a[8] = mul_sub2(p[0][0], p[1][1], p[1][0], p[0][1], fma);
    float den;
    if (fma) {      // m1 + m2 + m3 -> fma(x3, y3, fma(x1, y1, x2*y2))
        den = __fmul_rn(p[0][0], __fsub_rn(p[1][1], p[2][1]));
        den = __fmaf_rn(p[2][0], __fsub_rn(p[0][1], p[1][1]), den);
        den = __fmaf_rn(p[1][0], __fsub_rn(p[2][1], p[0][1]), den);
    } else {
        den = __fmul_rn(p[2][0], __fsub_rn(p[0][1], p[1][1]));
        den = __fadd_rn(den, __fmul_rn(p[0][0], __fsub_rn(p[1][1], p[2][1])));
        den = __fadd_rn(den, __fmul_rn(p[1][0], __fsub_rn(p[2][1], p[0][1])));
    }
#pragma unroll
    for (int k = 0; k < 9; k++) inv[k] = __fdiv_rn(a[k], den);
}

IPER_DEVINL float pixel_centre(int i, int is) {
    // (float)((2. * i + 1 - is) / is): evaluated in double, rounded once
    return __double2float_rn(__ddiv_rn(2.0 * (double)i + 1.0 - (double)is, (double)is));
}

IPER_DEVINL bool inside_face(const float* f, float xp, float yp) {
    if (__fmul_rn(__fsub_rn(yp, f[1]), __fsub_rn(f[3], f[0])) < __fmul_rn(__fsub_rn(xp, f[0]), __fsub_rn(f[4], f[1])))
        return false;
    if (__fmul_rn(__fsub_rn(yp, f[4]), __fsub_rn(f[6], f[3])) < __fmul_rn(__fsub_rn(xp, f[3]), __fsub_rn(f[7], f[4])))
        return false;
    if (__fmul_rn(__fsub_rn(yp, f[7]), __fsub_rn(f[0], f[6])) < __fmul_rn(__fsub_rn(xp, f[6]), __fsub_rn(f[1], f[7])))
        return false;
    return true;
}

// clamped + renormalised barycentric weights and perspective-correct depth; returns false when rejected
IPER_DEVINL bool weights_depth(const float* f, const float* inv, int xi, int yi, float near_, float far_, float* w,
                               float& zp, int fma) {
    const float fx = (float)xi, fy = (float)yi;
    float ws = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float t = mul_add(inv[3 * k + 0], fx, __fmul_rn(inv[3 * k + 1], fy), fma);
        t = __fadd_rn(t, inv[3 * k + 2]);
        double d = fmax((double)t, 0.0);  // min(max(w, 0.), 1.) — NaN -> 0 like upstream's fmax/fmin
        d = fmin(d, 1.0);
        w[k] = (float)d;
        ws = __fadd_rn(ws, w[k]);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) w[k] = __fdiv_rn(w[k], ws);
    float s = __fdiv_rn(w[0], f[2]);
    s = __fadd_rn(s, __fdiv_rn(w[1], f[5]));
    s = __fadd_rn(s, __fdiv_rn(w[2], f[8]));
    zp = __double2float_rn(__ddiv_rn(1.0, (double)s));
    // upstream: reject zp <= near or far <= zp, then keep only if zp < depth_min (initially far): a NaN depth
    // passes the first test but never the second, so eligibility is exactly near < zp < far
    return (zp > near_) && (zp < far_);
}

// F.grid_sample(img, grid) for one location, bilinear / zeros / align_corners=False, C planes of an (C,S,S) image
template <int C>
IPER_DEVINL void grid_sample_chw(const float* __restrict__ img, int S, float gx, float gy, float* out) {
    const float ix = ((gx + 1.f) * S - 1.f) * 0.5f;
    const float iy = ((gy + 1.f) * S - 1.f) * 0.5f;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix;
    const float wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
    const float nw = wx0 * wy0, ne = wx1 * wy0, sw = wx0 * wy1, se = wx1 * wy1;
    const bool vx0 = (x0 >= 0 && x0 < S), vx1 = (x1 >= 0 && x1 < S);
    const bool vy0 = (y0 >= 0 && y0 < S), vy1 = (y1 >= 0 && y1 < S);
#pragma unroll
    for (int c = 0; c < C; c++) {
        const float* p = img + (size_t)c * S * S;
        float acc = 0.f;
        if (vy0 && vx0) acc += p[y0 * S + x0] * nw;
        if (vy0 && vx1) acc += p[y0 * S + x1] * ne;
        if (vy1 && vx0) acc += p[y1 * S + x0] * sw;
        if (vy1 && vx1) acc += p[y1 * S + x1] * se;
        out[c] = acc;
    }
}

struct RasterArgs {
    // geometry source A: per-frame vertices + camera + shared topology (engine path, staged in smem)
    const float* verts;    // (B, nv, 3) or null
    const float* cams;     // (B, 3)
    const int32_t* faces;  // (nf, 3)
    // geometry source B: pre-gathered, already projected faces (neural_renderer seam)
    const float* face_verts;  // (B, nf, 3, 3): given (seam B1) or written by the setup kernel into the workspace
    const uint32_t* bins;     // (B, nf) packed tile ranges from the setup kernel (0xFFFFFFFF = culled)
    int B, nv, nf, S;
    int fma;               // rounding model of the rasteriser arithmetic (0: every op rounded, 1: nvcc -fmad contraction)
    float near_, far_, eye_z;
    // outputs (any may be null)
    int32_t* fim;  // (B,S,S)
    float* wim;    // (B,S,S,3)
    float* f2pts;  // (B,nf,3,2)   image-space (x, y) of every face corner, y flipped back (nmr.py:339-340)
    // fused per-frame generator inputs (all-or-nothing: enabled when tsf_inputs != null)
    const float* map_fn;     // (nf+1, 3)
    const float* f_uvs2img;  // (nf, 3, 2)
    const float* uv_img;     // (3, S, S)
    const float* src_f2pts;  // (ns, nf, 3, 2)
    int ns;
    float* tsf_inputs;  // (B, 6, S, S)  = cat[syn_img, cond]
    float* Tst;         // (B, ns, S, S, 2)
};

constexpr uint32_t BIN_CULLED = 0xFFFFFFFFu;

// one thread per (frame, face): projection (FROM_VERTS) / corners, f2pts, back-face cull, tile-range bin
template <bool FROM_VERTS>
__global__ void __launch_bounds__(256) raster_setup_kernel(const RasterArgs a, float* __restrict__ face_out,
                                                           uint32_t* __restrict__ bins) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (f >= a.nf) return;
    FaceGeom g;
    if (FROM_VERTS) {
        // a1: orthographic_proj_withz_idrot + y flip + look_at (identity rotation, z -> z - eye_z)
        const float s = a.cams[3 * b + 0], tx = a.cams[3 * b + 1], ty = a.cams[3 * b + 2];
        const float* v = a.verts + (size_t)b * a.nv * 3;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int vi = __ldg(a.faces + 3 * f + k);
            const float x = v[3 * vi + 0], y = v[3 * vi + 1], z = v[3 * vi + 2];
            g.v[3 * k + 0] = __fmul_rn(s, __fadd_rn(x, tx));
            g.v[3 * k + 1] = -__fmul_rn(s, __fadd_rn(y, ty));
            g.v[3 * k + 2] = __fsub_rn(z, a.eye_z);
        }
        float* o = face_out + ((size_t)b * a.nf + f) * 9;
#pragma unroll
        for (int k = 0; k < 9; k++) o[k] = g.v[k];
        if (a.f2pts) {      // a3: image-space corners with y flipped back (nmr.py:339-340)
            float* p = a.f2pts + ((size_t)b * a.nf + f) * 6;
#pragma unroll
            for (int k = 0; k < 3; k++) { p[2 * k] = g.v[3 * k]; p[2 * k + 1] = -g.v[3 * k + 1]; }
        }
    } else {
        const float* src = a.face_verts + ((size_t)b * a.nf + f) * 9;
#pragma unroll
        for (int k = 0; k < 9; k++) g.v[k] = __ldg(src + k);
    }
    uint32_t bin = BIN_CULLED;
    if (!is_backface(g.v)) {
        const int S = a.S;
        const float fS = (float)S;
        const float xmin = fminf(g.v[0], fminf(g.v[3], g.v[6])), xmax = fmaxf(g.v[0], fmaxf(g.v[3], g.v[6]));
        const float ymin = fminf(g.v[1], fminf(g.v[4], g.v[7])), ymax = fmaxf(g.v[1], fmaxf(g.v[4], g.v[7]));
        int x0 = 0, x1 = S - 1, y0 = 0, y1 = S - 1;     // pixel-index range in the upstream (xi, yi) frame
        const bool finite = (xmin == xmin) && (xmax == xmax) && (ymin == ymin) && (ymax == ymax) &&
                            fabsf(xmin) < 1e6f && fabsf(xmax) < 1e6f && fabsf(ymin) < 1e6f && fabsf(ymax) < 1e6f;
        if (finite) {       // NaN / huge corners: keep the whole image (such faces never win, like upstream)
            x0 = max(x0, (int)floorf((xmin * fS + fS - 1.f) * 0.5f) - 1);
            x1 = min(x1, (int)ceilf((xmax * fS + fS - 1.f) * 0.5f) + 1);
            y0 = max(y0, (int)floorf((ymin * fS + fS - 1.f) * 0.5f) - 1);
            y1 = min(y1, (int)ceilf((ymax * fS + fS - 1.f) * 0.5f) + 1);
        }
        if (x0 <= x1 && y0 <= y1) {
            // image rows r = S-1-yi: rows [S-1-y1, S-1-y0]
            const int r0 = S - 1 - y1, r1 = S - 1 - y0;
            bin = (uint32_t)(x0 / TILE_W) | ((uint32_t)(x1 / TILE_W) << 8) | ((uint32_t)(r0 / TILE_H) << 16) |
                  ((uint32_t)(r1 / TILE_H) << 24);
        }
    }
    bins[(size_t)b * a.nf + f] = bin;
}

__global__ void __launch_bounds__(RASTER_THREADS) raster_kernel(const RasterArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* zbuf = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned short* s_list = reinterpret_cast<unsigned short*>(smem_raw + RASTER_ZBUF_BYTES);
    float* s_xp = reinterpret_cast<float*>(smem_raw + RASTER_ZBUF_BYTES + BIN_CHUNK * sizeof(unsigned short));
    float* s_yp = s_xp + TILE_W;
    int* s_count = reinterpret_cast<int*>(s_yp + TILE_H);

    const int S = a.S, nf = a.nf;
    const int b = blockIdx.y;
    const int tiles_x = (S + TILE_W - 1) / TILE_W;
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int tx0 = tile_x * TILE_W;
    const int r0 = tile_y * TILE_H;  // image rows r0 .. r0+TILE_H-1
    const int tid = threadIdx.x;

    for (int i = tid; i < TILE_W * TILE_H; i += RASTER_THREADS) zbuf[i] = ZBUF_EMPTY;

    __syncthreads();

    // pixel-index bounds of this tile in the upstream (xi, yi) frame: yi = S-1-row
    const int xi_lo = tx0, xi_hi = min(tx0 + TILE_W, S) - 1;
    const int yi_hi = S - 1 - r0, yi_lo = S - 1 - (min(r0 + TILE_H, S) - 1);
    const float fS = (float)S;
    // pixel-centre tables of the tile (the double-precision divisions happen once per CTA, not per test)
    if (tid < TILE_W) s_xp[tid] = pixel_centre(tx0 + tid, S);
    else if (tid < TILE_W + TILE_H) s_yp[tid - TILE_W] = pixel_centre(S - 1 - (r0 + tid - TILE_W), S);

    auto load_face = [&](int f, FaceGeom& g) {
        const float* src = a.face_verts + ((size_t)b * nf + f) * 9;
#pragma unroll
        for (int k = 0; k < 9; k++) g.v[k] = __ldg(src + k);
    };
    // conservative pixel range of a face's bounding box (grown by one pixel) clipped to the tile
    auto face_range = [&](const FaceGeom& g, int& x0, int& x1, int& y0, int& y1) {
        const float xmin = fminf(g.v[0], fminf(g.v[3], g.v[6])), xmax = fmaxf(g.v[0], fmaxf(g.v[3], g.v[6]));
        const float ymin = fminf(g.v[1], fminf(g.v[4], g.v[7])), ymax = fmaxf(g.v[1], fmaxf(g.v[4], g.v[7]));
        x0 = xi_lo; x1 = xi_hi; y0 = yi_lo; y1 = yi_hi;
        const bool finite = (xmin == xmin) && (xmax == xmax) && (ymin == ymin) && (ymax == ymax) &&
                            fabsf(xmin) < 1e6f && fabsf(xmax) < 1e6f && fabsf(ymin) < 1e6f && fabsf(ymax) < 1e6f;
        if (finite) {
            x0 = max(x0, (int)floorf((xmin * fS + fS - 1.f) * 0.5f) - 1);
            x1 = min(x1, (int)ceilf((xmax * fS + fS - 1.f) * 0.5f) + 1);
            y0 = max(y0, (int)floorf((ymin * fS + fS - 1.f) * 0.5f) - 1);
            y1 = min(y1, (int)ceilf((ymax * fS + fS - 1.f) * 0.5f) + 1);
        }
    };

    // ---- per chunk of faces: (1) BIN front faces whose bbox touches the tile into a compact shared list,
    //      (2) scan-convert the list one WARP per face, lanes over the pixels of (bbox ∩ tile) ----
    const int warp = tid >> 5, lane = tid & 31;
    for (int base = 0; base < nf; base += BIN_CHUNK) {
        if (tid == 0) *s_count = 0;
        __syncthreads();
        const int end = min(base + BIN_CHUNK, nf);
        for (int f = base + tid; f < end; f += RASTER_THREADS) {
            const uint32_t bin = __ldg(a.bins + (size_t)b * nf + f);
            if (bin == BIN_CULLED) continue;
            const int bx0 = bin & 0xFF, bx1 = (bin >> 8) & 0xFF, by0 = (bin >> 16) & 0xFF, by1 = bin >> 24;
            if (tile_x < bx0 || tile_x > bx1 || tile_y < by0 || tile_y > by1) continue;
            s_list[atomicAdd(s_count, 1)] = (unsigned short)(f - base);
        }
        __syncthreads();
        const int count = *s_count;
        for (int c = warp; c < count; c += RASTER_THREADS / 32) {
            const int f = base + s_list[c];
            FaceGeom g;
            load_face(f, g);
            int x0, x1, y0, y1;
            face_range(g, x0, x1, y0, y1);
            float inv[9];
            face_inverse(g.v, S, inv, a.fma);
            const int bw = x1 - x0 + 1, npx = bw * (y1 - y0 + 1);
            for (int i = lane; i < npx; i += 32) {
                const int xi = x0 + i % bw, yi = y0 + i / bw;
                const float xp = s_xp[xi - tx0], yp = s_yp[(S - 1 - yi) - r0];
                if (!inside_face(g.v, xp, yp)) continue;
                float w[3], zp;
                if (!weights_depth(g.v, inv, xi, yi, a.near_, a.far_, w, zp, a.fma)) continue;
                const unsigned long long key = ((unsigned long long)__float_as_uint(zp) << 32) | (unsigned)f;
                atomicMin(&zbuf[(S - 1 - yi - r0) * TILE_W + (xi - tx0)], key);
            }
        }
    }
    __syncthreads();

    // ---- resolve: coalesced writes of fim / wim / cond / flow / sampled UV image ----
    const size_t SS = (size_t)S * S;
    for (int i = tid; i < TILE_W * TILE_H; i += RASTER_THREADS) {
        const int lx = i % TILE_W, ly = i / TILE_W;
        const int xi = tx0 + lx, row = r0 + ly;
        if (xi >= S || row >= S) continue;
        const unsigned long long key = zbuf[i];
        const int fn = (key == ZBUF_EMPTY) ? -1 : (int)(unsigned)(key & 0xFFFFFFFFull);
        float w[3] = {0.f, 0.f, 0.f};
        if (fn >= 0) {
            FaceGeom g;
            load_face(fn, g);
            float inv[9], zp;
            face_inverse(g.v, S, inv, a.fma);
            weights_depth(g.v, inv, xi, S - 1 - row, a.near_, a.far_, w, zp, a.fma);
        }
        const size_t pix = (size_t)row * S + xi;
        if (a.fim) a.fim[(size_t)b * SS + pix] = fn;
        if (a.wim) {
            float* o = a.wim + ((size_t)b * SS + pix) * 3;
            o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
        }
        if (a.tsf_inputs) {
            // a4: cond = map_fn[fim] (index -1 -> last row)
            const float* mf = a.map_fn + 3 * (size_t)(fn >= 0 ? fn : nf);
            float* ti = a.tsf_inputs + (size_t)b * 6 * SS + pix;
            ti[3 * SS] = __ldg(mf + 0);
            ti[4 * SS] = __ldg(mf + 1);
            ti[5 * SS] = __ldg(mf + 2);
            // a6/a7: Tuv2t then syn_img = grid_sample(uv_img, Tuv2t); background T = -2 samples zeros
            float syn[3] = {0.f, 0.f, 0.f};
            if (fn >= 0) {
                const float* fu = a.f_uvs2img + (size_t)fn * 6;
                const float gx = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(fu + 0), w[0]), __fmul_rn(__ldg(fu + 2), w[1])),
                                           __fmul_rn(__ldg(fu + 4), w[2]));
                const float gy = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(fu + 1), w[0]), __fmul_rn(__ldg(fu + 3), w[1])),
                                           __fmul_rn(__ldg(fu + 5), w[2]));
                grid_sample_chw<3>(a.uv_img, S, gx, gy, syn);
            }
            ti[0] = syn[0];
            ti[SS] = syn[1];
            ti[2 * SS] = syn[2];
            // a8: Tst[b, s] from each source's f2pts
            for (int s = 0; s < a.ns; s++) {
                float2 t = make_float2(-2.f, -2.f);
                if (fn >= 0) {
                    const float* fp = a.src_f2pts + ((size_t)s * nf + fn) * 6;
                    t.x = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(fp + 0), w[0]), __fmul_rn(__ldg(fp + 2), w[1])),
                                    __fmul_rn(__ldg(fp + 4), w[2]));
                    t.y = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(fp + 1), w[0]), __fmul_rn(__ldg(fp + 3), w[1])),
                                    __fmul_rn(__ldg(fp + 5), w[2]));
                }
                reinterpret_cast<float2*>(a.Tst)[((size_t)b * a.ns + s) * SS + pix] = t;
            }
        }
    }
}

// cal_bc_transform (nmr.py:713-757): T[p] = sum_k wim[p,k] * f2pts[fim[p],k,:], background (-2,-2)
__global__ void flow_kernel(const float* __restrict__ f2pts, int f2pts_per_item, const int32_t* __restrict__ fim,
                            const float* __restrict__ wim, int nb, int nsrc, int nf, size_t SS, float* __restrict__ T) {
    const size_t total = (size_t)nb * nsrc * SS;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = i % SS;
        const int s = (int)((i / SS) % nsrc);
        const int b = (int)(i / (SS * nsrc));
        const int fn = fim[(size_t)b * SS + pix];
        float2 t = make_float2(-2.f, -2.f);
        if (fn >= 0) {
            const float* w = wim + ((size_t)b * SS + pix) * 3;
            const float* fp = f2pts + ((size_t)(f2pts_per_item ? b : s) * nf + fn) * 6;
            t.x = __fadd_rn(__fadd_rn(__fmul_rn(fp[0], w[0]), __fmul_rn(fp[2], w[1])), __fmul_rn(fp[4], w[2]));
            t.y = __fadd_rn(__fadd_rn(__fmul_rn(fp[1], w[0]), __fmul_rn(fp[3], w[1])), __fmul_rn(fp[5], w[2]));
        }
        reinterpret_cast<float2*>(T)[i] = t;
    }
}

// encode_fim (nmr.py:390-401): out[b, c, y, x] = map_fn[fim[b,y,x] (or last row for -1)][c]   (transpose=True)
__global__ void encode_fim_kernel(const int32_t* __restrict__ fim, const float* __restrict__ map_fn, int nb, int nf,
                                  int ch, size_t SS, int transpose, float* __restrict__ out) {
    const size_t total = (size_t)nb * SS;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int fn = fim[i];
        const float* mf = map_fn + (size_t)ch * (fn >= 0 ? fn : nf);
        const size_t b = i / SS, pix = i % SS;
        for (int c = 0; c < ch; c++) {
            if (transpose) out[(b * ch + c) * SS + pix] = mf[c];
            else out[i * ch + c] = mf[c];
        }
    }
}

// LWB.resize_trans (attlwb_spade_resunet.py:175-182): bilinear, align_corners=True, (N,S,S,2) -> (N,h,w,2)
__global__ void flow_resize_kernel(const float* __restrict__ T, int n, int S, int h, int w, float* __restrict__ out) {
    const float sy = (h > 1) ? (float)(S - 1) / (float)(h - 1) : 0.f;
    const float sx = (w > 1) ? (float)(S - 1) / (float)(w - 1) : 0.f;
    const size_t total = (size_t)n * h * w;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w), y = (int)((i / w) % h);
        const size_t b = i / ((size_t)w * h);
        const float fy = sy * y, fx = sx * x;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < S - 1 ? 1 : 0), x1 = x0 + (x0 < S - 1 ? 1 : 0);
        const float ly1 = fy - y0, ly0 = 1.f - ly1, lx1 = fx - x0, lx0 = 1.f - lx1;
        const float2* src = reinterpret_cast<const float2*>(T) + b * S * S;
        const float2 a = src[(size_t)y0 * S + x0], bb = src[(size_t)y0 * S + x1];
        const float2 c = src[(size_t)y1 * S + x0], d = src[(size_t)y1 * S + x1];
        float2 o;
        o.x = ly0 * (lx0 * a.x + lx1 * bb.x) + ly1 * (lx0 * c.x + lx1 * d.x);
        o.y = ly0 * (lx0 * a.y + lx1 * bb.y) + ly1 * (lx0 * c.y + lx1 * d.y);
        reinterpret_cast<float2*>(out)[i] = o;
    }
}

// ---- SMPLRenderer.get_vis_f2pts (nmr.py:639-681) ---------------------------------------------------------------------
// faces visible in fim, minus the SMALLEST unique value of the map (`fim.unique()[1:]`: the background -1 when present,
// otherwise the lowest visible face id), together with their top_k UV-nearest neighbours keep their corner coordinates;
// every other face is set to -2.  Three passes over byte flags instead of two torch.unique sorts with host syncs.
__global__ void vis_mark_kernel(const int32_t* __restrict__ fim, int nf, size_t SS, unsigned char* __restrict__ vis,
                                int* __restrict__ min_id) {
    const int b = blockIdx.y;
    int local_min = 0x7fffffff;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < SS; i += (size_t)gridDim.x * blockDim.x) {
        const int fn = __ldg(fim + (size_t)b * SS + i);
        local_min = min(local_min, fn);
        if (fn >= 0 && fn < nf) vis[(size_t)b * nf + fn] = 1;
    }
    for (int o = 16; o > 0; o >>= 1) local_min = min(local_min, __shfl_xor_sync(0xffffffffu, local_min, o));
    if ((threadIdx.x & 31) == 0) atomicMin(min_id + b, local_min);
}
__global__ void vis_spread_kernel(const unsigned char* __restrict__ vis, const int* __restrict__ min_id,
                                  const long long* __restrict__ nearest, int top_k, int nf, unsigned char* __restrict__ keep) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (f >= nf || !vis[(size_t)b * nf + f] || f == min_id[b]) return;
    for (int k = 0; k < top_k; k++) {
        const long long nb = __ldg(nearest + (size_t)f * top_k + k);
        if (nb >= 0 && nb < nf) keep[(size_t)b * nf + nb] = 1;
    }
}
__global__ void vis_apply_kernel(const float* __restrict__ f2pts, const unsigned char* __restrict__ keep, int elems, size_t total,
                                 float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = keep[i / elems] ? f2pts[i] : -2.0f;
}

static size_t vis_ws_bytes(int B, int nf) {
    // min_id (B) ints, padded to 16 bytes, then vis (B, nf) and keep (B, nf) byte flags
    return (((size_t)B * sizeof(int) + 15) & ~(size_t)15) + 2 * (size_t)B * nf;
}

static size_t raster_ws_bytes(int B, int nf, bool from_verts) {
    // bins (B, nf) u32, then — when the corners are computed here — face corners (B, nf, 9) f32
    return (size_t)B * nf * sizeof(uint32_t) + (from_verts ? (size_t)B * nf * 9 * sizeof(float) : 0);
}

// process-wide rounding model: -1 = not chosen yet (first launch reads IPER_RASTER_FMA, default 0)
static int g_raster_fma = -1;
static int raster_fma_mode() {
    if (g_raster_fma < 0) {
        const char* v = getenv("IPER_RASTER_FMA");
        g_raster_fma = (v && atoi(v) != 0) ? 1 : 0;
    }
    return g_raster_fma;
}

static int launch_raster(RasterArgs a, bool from_verts, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int S = a.S;
    a.fma = raster_fma_mode();
    const int tiles_x = (S + TILE_W - 1) / TILE_W, tiles_y = (S + TILE_H - 1) / TILE_H;
    IPER_REQUIRE(tiles_x <= 256 && tiles_y <= 256, "rasteriser: image size %d exceeds the 8-bit tile index range", S);
    IPER_REQUIRE(a.nf <= (1 << 24), "rasteriser: too many faces (%d)", a.nf);
    IPER_REQUIRE(workspace && workspace_bytes >= raster_ws_bytes(a.B, a.nf, from_verts),
                 "rasteriser: workspace of %zu bytes needed (iper_raster_workspace_bytes), got %zu",
                 raster_ws_bytes(a.B, a.nf, from_verts), workspace_bytes);
    uint32_t* bins = reinterpret_cast<uint32_t*>(workspace);
    float* corners = reinterpret_cast<float*>(bins + (size_t)a.B * a.nf);
    dim3 sgrid((a.nf + 255) / 256, a.B);
    if (from_verts) {
        raster_setup_kernel<true><<<sgrid, 256, 0, stream>>>(a, corners, bins);
        a.face_verts = corners;
    } else {
        raster_setup_kernel<false><<<sgrid, 256, 0, stream>>>(a, nullptr, bins);
    }
    IPER_CHECK_CUDA(cudaGetLastError());
    a.bins = bins;
    dim3 grid(tiles_x * tiles_y, a.B);
    raster_kernel<<<grid, RASTER_THREADS, RASTER_FIXED_SMEM, stream>>>(a);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace iper

using namespace iper;

extern "C" size_t iper_vis_f2pts_workspace_bytes(int B, int nf) {
    if (B <= 0 || nf <= 0) return 0;
    return vis_ws_bytes(B, nf);
}

extern "C" int iper_vis_f2pts(const float* f2pts, int elems_per_face, const int32_t* fim, const int64_t* face_k_nearest,
                              int top_k, int B, int nf, int S, float* out, void* workspace, size_t workspace_bytes,
                              iper_stream_t stream) {
    IPER_REQUIRE(B >= 0 && nf > 0 && S > 0 && top_k > 0 && elems_per_face > 0, "iper_vis_f2pts: bad sizes");
    if (B == 0) return 0;
    IPER_REQUIRE(f2pts && fim && face_k_nearest && out, "iper_vis_f2pts: null pointer");
    IPER_REQUIRE(workspace && workspace_bytes >= vis_ws_bytes(B, nf), "iper_vis_f2pts: workspace of %zu bytes needed, got %zu",
                 vis_ws_bytes(B, nf), workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    int* min_id = reinterpret_cast<int*>(workspace);
    const size_t head = ((size_t)B * sizeof(int) + 15) & ~(size_t)15;
    unsigned char* vis = reinterpret_cast<unsigned char*>(workspace) + head;
    unsigned char* keep = vis + (size_t)B * nf;
    IPER_CHECK_CUDA(cudaMemsetAsync(min_id, 0x7f, head, st));                    // 0x7f7f7f7f: larger than any face id
    IPER_CHECK_CUDA(cudaMemsetAsync(vis, 0, 2 * (size_t)B * nf, st));
    const size_t SS = (size_t)S * S;
    vis_mark_kernel<<<dim3((unsigned)min((SS + 255) / 256, (size_t)64), B), 256, 0, st>>>(fim, nf, SS, vis, min_id);
    IPER_CHECK_CUDA(cudaGetLastError());
    vis_spread_kernel<<<dim3((nf + 255) / 256, B), 256, 0, st>>>(vis, min_id, reinterpret_cast<const long long*>(face_k_nearest),
                                                                top_k, nf, keep);
    IPER_CHECK_CUDA(cudaGetLastError());
    const size_t total = (size_t)B * nf * elems_per_face;
    vis_apply_kernel<<<(unsigned)min((total + 255) / 256, (size_t)148 * 16), 256, 0, st>>>(f2pts, keep, elems_per_face, total, out);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_raster_set_contraction(int mode) {
    IPER_REQUIRE(mode == 0 || mode == 1, "iper_raster_set_contraction: mode %d not in {0 = every op rounded, 1 = fmad model}", mode);
    g_raster_fma = mode;
    return 0;
}
extern "C" int iper_raster_get_contraction(void) { return raster_fma_mode(); }

extern "C" size_t iper_raster_workspace_bytes(int B, int nf, int from_verts) {
    if (B <= 0 || nf <= 0) return 0;
    return raster_ws_bytes(B, nf, from_verts != 0);
}

extern "C" int iper_rasterize_faces(const float* faces, int B, int nf, int S, float near_, float far_, int32_t* fim,
                                    float* wim, void* workspace, size_t workspace_bytes, iper_stream_t stream) {
    IPER_REQUIRE(B >= 0 && nf > 0 && S > 0, "iper_rasterize_faces: bad sizes B=%d nf=%d S=%d", B, nf, S);
    if (B == 0) return 0;   // empty batch: nothing to do (pointers may be null)
    IPER_REQUIRE(faces && fim && wim, "iper_rasterize_faces: null pointer");
    RasterArgs a = {};
    a.face_verts = faces; a.B = B; a.nf = nf; a.S = S; a.near_ = near_; a.far_ = far_;
    a.fim = fim; a.wim = wim;
    return launch_raster(a, false, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int iper_raster_frames(const float* verts, const float* cams, const int32_t* faces, int B, int nv, int nf,
                                  int S, float eye_z, float near_, float far_, int32_t* fim, float* wim, float* f2pts,
                                  const float* map_fn, const float* f_uvs2img, const float* uv_img,
                                  const float* src_f2pts, int ns, float* tsf_inputs, float* Tst,
                                  void* workspace, size_t workspace_bytes, iper_stream_t stream) {
    IPER_REQUIRE(B >= 0 && nv > 0 && nf > 0 && S > 0, "iper_raster_frames: bad sizes");
    if (B == 0) return 0;
    IPER_REQUIRE(verts && cams && faces, "iper_raster_frames: null geometry pointer");
    if (tsf_inputs || Tst) {
        IPER_REQUIRE(tsf_inputs && Tst && map_fn && f_uvs2img && uv_img && src_f2pts && ns > 0,
                     "iper_raster_frames: the fused frame-input outputs need map_fn, f_uvs2img, uv_img, src_f2pts, ns");
    }
    RasterArgs a = {};
    a.verts = verts; a.cams = cams; a.faces = faces; a.B = B; a.nv = nv; a.nf = nf; a.S = S;
    a.near_ = near_; a.far_ = far_; a.eye_z = eye_z;
    a.fim = fim; a.wim = wim; a.f2pts = f2pts;
    a.map_fn = map_fn; a.f_uvs2img = f_uvs2img; a.uv_img = uv_img; a.src_f2pts = src_f2pts; a.ns = ns;
    a.tsf_inputs = tsf_inputs; a.Tst = Tst;
    return launch_raster(a, true, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int iper_flow_from_fim_wim(const float* f2pts, int f2pts_per_item, const int32_t* fim, const float* wim,
                                      int nb, int nsrc, int nf, int S, float* T, iper_stream_t stream) {
    IPER_REQUIRE(f2pts && fim && wim && T, "iper_flow_from_fim_wim: null pointer");
    IPER_REQUIRE(!f2pts_per_item || nsrc == 1, "iper_flow_from_fim_wim: per-item f2pts needs nsrc == 1");
    const size_t total = (size_t)nb * nsrc * S * S;
    if (total == 0) return 0;
    const int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
    flow_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(f2pts, f2pts_per_item, fim, wim, nb, nsrc, nf, (size_t)S * S, T);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_encode_fim(const int32_t* fim, const float* map_fn, int nb, int nf, int ch, int S, int transpose,
                               float* out, iper_stream_t stream) {
    IPER_REQUIRE(fim && map_fn && out, "iper_encode_fim: null pointer");
    const size_t total = (size_t)nb * S * S;
    if (total == 0) return 0;
    const int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
    encode_fim_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(fim, map_fn, nb, nf, ch, (size_t)S * S, transpose, out);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_flow_resize(const float* T, int n, int S, int h, int w, float* out, iper_stream_t stream) {
    IPER_REQUIRE(T && out, "iper_flow_resize: null pointer");
    const size_t total = (size_t)n * h * w;
    if (total == 0) return 0;
    const int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
    flow_resize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(T, n, S, h, w, out);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}
