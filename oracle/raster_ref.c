/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (ipercore_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: this is a CPU restatement of the face-index-map + weight-map rasteriser of the
 * third-party CUDA extension `neural_renderer` (fork iPERDance/neural_renderer pinned at commit
 * e5f54f71a8941acf372514eb92e289872f272653 by /root/reference/requirements/build.txt:3), which the reference
 * calls at iPERCore/tools/human_digitalizer/renders/nmr.py:337 and :356
 * (`nr.rasterize_face_index_map_and_weight_map(faces, image_size, False)`).  The extension is NOT vendored
 * under /root/reference, is not installed here and cannot be fetched (no network); the reference's own test
 * for this seam (tests/test_human_digitalizer/test_renders.py) only visualises and holds no golden values.
 * The algorithm below restates the published upstream kernel pair
 * (`forward_face_index_map_cuda_kernel_1/_2`, daniilidis-group/neural_renderer, rasterize_cuda_kernel.cu)
 * at source-level IEEE semantics: every float operation is a separately rounded binary32 operation
 * (no FMA contraction; build with -ffp-contract=off), and the places where upstream's double literals
 * (`2.`, `0.5`, `1.`, `0.`) promote an expression to double are evaluated in double and rounded once.
 *
 * CONTRACTION VARIANT (-DORACLE_FMA, symbols suffixed _fma): upstream is compiled by nvcc, whose default -fmad=true lets
 * the compiler contract a*b+c into one fused multiply-add, so the shipped extension most likely does NOT round every
 * product.  Which products fuse is the compiler's choice; the variant models the LLVM/NVPTX DAG-combiner rules
 * (fadd(fmul x y, z) -> fma(x,y,z), fsub(fmul x y, z) -> fma(x,y,-z), no reassociation) on upstream's expressions:
 *   p          = 0.5 * (fma(v, is, is) - 1)
 *   adjugate   a*b - c*d        -> fma(a, b, -(c*d))
 *   determinant m1 + m2 + m3    -> fma(x3, y3, fma(x1, y1, x2*y2))
 *   weights    i0*xi + i1*yi + i2 -> fma(i0, xi, i1*yi) + i2
 * Comparisons of two products (back-face and edge tests) and the divisions have no add to fuse and are unchanged.  It
 * cannot be checked here either (the extension is absent); it exists so that the CUDA rasteriser's matching switch
 * (iper_raster_set_contraction) can be validated against a real build of the fork later.  tests/ run both.
 *
 * Semantics:
 *   kernel 1, per face : skip when back-facing; pixel-space corners p = 0.5*(v*is + is - 1);
 *                        face_inv = inverse of [[x0,x1,x2],[y0,y1,y2],[1,1,1]] (adjugate / determinant).
 *   kernel 2, per pixel: row r of the output image is yi = is-1-r (vertical flip), xi = column;
 *                        pixel centre (xp,yp) = ((2xi+1-is)/is, (2yi+1-is)/is);
 *                        faces visited in ascending index; skip back-facing; three edge inequalities;
 *                        w = face_inv * [xi, yi, 1], clamp to [0,1], renormalise by the sum;
 *                        zp = 1/(w0/z0 + w1/z1 + w2/z2); reject zp <= near or far <= zp;
 *                        keep when zp < depth_min (strict: lowest face index wins a tie).
 *   outputs            : face_index_map int32 (init -1), weight_map float32 x3 (init 0).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORACLE_FMA
#define NAME(x) x##_fma
#define MULADD(a, b, c) fmaf((a), (b), (c))                 /* a*b + c, fused        */
#define MULSUB2(a, b, c, d) fmaf((a), (b), -((c) * (d)))    /* a*b - c*d             */
#else
#define NAME(x) x
#define MULADD(a, b, c) ((a) * (b) + (c))                   /* separately rounded (-ffp-contract=off) */
#define MULSUB2(a, b, c, d) ((a) * (b) - (c) * (d))
#endif

static inline int backface(const float *f)
{
    return (f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]);
}

/* kernel 1: per-face inverse, written only for front-facing faces (others stay 0, never read) */
static void face_setup(const float *f, int is, float *inv)
{
    float p[3][2];
    for (int n = 0; n < 3; n++)
        for (int d = 0; d < 2; d++) {
            /* 0.5 * (face * is + is - 1): float chain, the final 0.5* is exact in either precision */
            float t = MULADD(f[3 * n + d], (float)is, (float)is);
            t = t - 1.0f;
            p[n][d] = (float)(0.5 * (double)t);
        }
    float a[9] = {
        p[1][1] - p[2][1], p[2][0] - p[1][0], MULSUB2(p[1][0], p[2][1], p[2][0], p[1][1]),
        p[2][1] - p[0][1], p[0][0] - p[2][0], MULSUB2(p[2][0], p[0][1], p[0][0], p[2][1]),
        p[0][1] - p[1][1], p[1][0] - p[0][0], MULSUB2(p[0][0], p[1][1], p[1][0], p[0][1])};
#ifdef ORACLE_FMA
    float den = p[0][0] * (p[1][1] - p[2][1]);
    den = fmaf(p[2][0], p[0][1] - p[1][1], den);
    den = fmaf(p[1][0], p[2][1] - p[0][1], den);
#else
    float den = p[2][0] * (p[0][1] - p[1][1]);
    den = den + p[0][0] * (p[1][1] - p[2][1]);
    den = den + p[1][0] * (p[2][1] - p[0][1]);
#endif
    for (int k = 0; k < 9; k++) inv[k] = a[k] / den;
}

static inline float clamp01(float w)
{
    /* min(max(w, 0.), 1.) with fmax/fmin NaN semantics (NaN -> 0) */
    double d = fmax((double)w, 0.0);
    d = fmin(d, 1.0);
    return (float)d;
}

/*
 * faces : (bs, nf, 3, 3) float32, NDC x,y in [-1,1] (+y up), z = depth along the view axis
 * fim   : (bs, is, is) int32     wim : (bs, is, is, 3) float32
 * returns 0
 */
int NAME(oracle_rasterize_fim_wim)(const float *faces, int bs, int nf, int is, float near_, float far_,
                             int32_t *fim, float *wim)
{
    float *inv = (float *)calloc((size_t)bs * nf * 9, sizeof(float));
    unsigned char *back = (unsigned char *)malloc((size_t)bs * nf);
    if (!inv || !back) return 1;
    for (long i = 0; i < (long)bs * nf; i++) {
        back[i] = (unsigned char)backface(faces + 9 * i);
        if (!back[i]) face_setup(faces + 9 * i, is, inv + 9 * i);
    }
    for (int bn = 0; bn < bs; bn++) {
        for (int pn = 0; pn < is * is; pn++) {
            const int yi = is - 1 - (pn / is);
            const int xi = pn % is;
            const float yp = (float)((2. * yi + 1 - is) / is);
            const float xp = (float)((2. * xi + 1 - is) / is);
            float depth_min = far_;
            int fmin_ = -1;
            float wmin[3] = {0.f, 0.f, 0.f};
            const float *fb = faces + (size_t)bn * nf * 9;
            const float *ib = inv + (size_t)bn * nf * 9;
            const unsigned char *bb = back + (size_t)bn * nf;
            for (int fn = 0; fn < nf; fn++) {
                if (bb[fn]) continue;
                const float *f = fb + 9 * fn;
                if (((yp - f[1]) * (f[3] - f[0]) < (xp - f[0]) * (f[4] - f[1])) ||
                    ((yp - f[4]) * (f[6] - f[3]) < (xp - f[3]) * (f[7] - f[4])) ||
                    ((yp - f[7]) * (f[0] - f[6]) < (xp - f[6]) * (f[1] - f[7])))
                    continue;
                const float *fi = ib + 9 * fn;
                float w[3];
                for (int k = 0; k < 3; k++) {
                    float t = MULADD(fi[3 * k + 0], (float)xi, fi[3 * k + 1] * (float)yi);
                    t = t + fi[3 * k + 2];
                    w[k] = t;
                }
                float ws = 0.f;
                for (int k = 0; k < 3; k++) { w[k] = clamp01(w[k]); ws = ws + w[k]; }
                for (int k = 0; k < 3; k++) w[k] = w[k] / ws;
                float s = w[0] / f[2];
                s = s + w[1] / f[5];
                s = s + w[2] / f[8];
                const float zp = (float)(1. / (double)s);
                if (zp <= near_ || far_ <= zp) continue;
                if (zp < depth_min) {
                    depth_min = zp; fmin_ = fn;
                    wmin[0] = w[0]; wmin[1] = w[1]; wmin[2] = w[2];
                }
            }
            const size_t o = (size_t)bn * is * is + pn;
            fim[o] = fmin_;
            wim[3 * o + 0] = fmin_ >= 0 ? wmin[0] : 0.f;
            wim[3 * o + 1] = fmin_ >= 0 ? wmin[1] : 0.f;
            wim[3 * o + 2] = fmin_ >= 0 ? wmin[2] : 0.f;
        }
    }
    free(inv); free(back);
    return 0;
}

/*
 * Same result, faster: per-face bounding boxes cull the inner loop (a face whose NDC bounding box, grown
 * by one pixel, does not contain the pixel centre cannot pass the three edge tests up to rounding that is
 * orders of magnitude below one pixel).  Used by the CPU baseline and by large-size tests; the plain loop
 * above is the definition and tests/ checks the two against each other.
 */
int NAME(oracle_rasterize_fim_wim_fast)(const float *faces, int bs, int nf, int is, float near_, float far_,
                                  int32_t *fim, float *wim)
{
    const size_t npix = (size_t)is * is;
    for (size_t i = 0; i < (size_t)bs * npix; i++) fim[i] = -1;
    memset(wim, 0, sizeof(float) * 3 * bs * npix);
    float *depth = (float *)malloc(sizeof(float) * npix);
    if (!depth) return 1;
    const double margin = 2.0 / is;
    for (int bn = 0; bn < bs; bn++) {
        for (size_t i = 0; i < npix; i++) depth[i] = far_;
        int32_t *fimb = fim + (size_t)bn * npix;
        float *wimb = wim + 3 * (size_t)bn * npix;
        for (int fn = 0; fn < nf; fn++) {
            const float *f = faces + ((size_t)bn * nf + fn) * 9;
            if (backface(f)) continue;
            float inv[9];
            face_setup(f, is, inv);
            double xmin = fmin(f[0], fmin(f[3], f[6])) - margin, xmax = fmax(f[0], fmax(f[3], f[6])) + margin;
            double ymin = fmin(f[1], fmin(f[4], f[7])) - margin, ymax = fmax(f[1], fmax(f[4], f[7])) + margin;
            if (!(xmin == xmin) || !(xmax == xmax) || !(ymin == ymin) || !(ymax == ymax)) {
                xmin = ymin = -1e30; xmax = ymax = 1e30;  /* NaN corner: no culling */
            }
            /* xp = (2xi+1-is)/is  ->  xi = (xp*is + is - 1)/2 */
            double lo = floor((xmin * is + is - 1) / 2.0), hi = ceil((xmax * is + is - 1) / 2.0);
            int x0 = lo < 0 ? 0 : (lo > is - 1 ? is : (int)lo), x1 = hi > is - 1 ? is - 1 : (hi < 0 ? -1 : (int)hi);
            lo = floor((ymin * is + is - 1) / 2.0); hi = ceil((ymax * is + is - 1) / 2.0);
            int y0 = lo < 0 ? 0 : (lo > is - 1 ? is : (int)lo), y1 = hi > is - 1 ? is - 1 : (hi < 0 ? -1 : (int)hi);
            for (int yi = y0; yi <= y1; yi++) {
                const float yp = (float)((2. * yi + 1 - is) / is);
                for (int xi = x0; xi <= x1; xi++) {
                    const float xp = (float)((2. * xi + 1 - is) / is);
                    if (((yp - f[1]) * (f[3] - f[0]) < (xp - f[0]) * (f[4] - f[1])) ||
                        ((yp - f[4]) * (f[6] - f[3]) < (xp - f[3]) * (f[7] - f[4])) ||
                        ((yp - f[7]) * (f[0] - f[6]) < (xp - f[6]) * (f[1] - f[7])))
                        continue;
                    float w[3];
                    for (int k = 0; k < 3; k++) {
                        float t = MULADD(inv[3 * k + 0], (float)xi, inv[3 * k + 1] * (float)yi);
                        t = t + inv[3 * k + 2];
                        w[k] = t;
                    }
                    float ws = 0.f;
                    for (int k = 0; k < 3; k++) { w[k] = clamp01(w[k]); ws = ws + w[k]; }
                    for (int k = 0; k < 3; k++) w[k] = w[k] / ws;
                    float s = w[0] / f[2];
                    s = s + w[1] / f[5];
                    s = s + w[2] / f[8];
                    const float zp = (float)(1. / (double)s);
                    if (zp <= near_ || far_ <= zp) continue;
                    const size_t o = (size_t)(is - 1 - yi) * is + xi;
                    /* ascending fn + strict '<' == lowest index wins ties, as in the definition */
                    if (zp < depth[o]) {
                        depth[o] = zp; fimb[o] = fn;
                        wimb[3 * o] = w[0]; wimb[3 * o + 1] = w[1]; wimb[3 * o + 2] = w[2];
                    }
                }
            }
        }
    }
    free(depth);
    return 0;
}
