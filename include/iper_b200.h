/*
 * iper_b200 — C ABI of the B200-native (sm_100a) kernels for iPERCore's motion-imitation hot path.
 *
 * Boundary rules: plain C, raw DEVICE pointers + sizes + a CUDA stream; no torch types; the
 * library never allocates device memory, never synchronises and launches only on the stream it is given.
 * Every function returns 0 on success; non-zero = error, message via iper_last_error() (thread-local).
 * The Python shims under ipercore_b200/ keep the reference's operator API on top of these entry points; the
 * reference-side binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Reference citations are relative to /root/reference/ (iPERDance/iPERCore @ fcf9a18).
 */
#ifndef IPER_B200_H
#define IPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* iper_stream_t; /* == cudaStream_t */

const char* iper_last_error(void);
int iper_abi_version(void);

/* ------------------------------------------------------------------------------------------------------------
 * Seam B1 — replaces neural_renderer.rasterize_face_index_map_and_weight_map(faces, image_size, False)
 * (third-party CUDA ext; call sites iPERCore/tools/human_digitalizer/renders/nmr.py:337 and :356).
 * faces (B,nf,3,3) f32 NDC  ->  fim (B,S,S) i32 (-1 = background), wim (B,S,S,3) f32 (0 on background).
 * Correct for any B (the reference loops around an upstream B==3 bug, nmr.py:892-918).
 * workspace: device scratch of iper_raster_workspace_bytes(B, nf, from_verts) bytes (per-face tile bins, and the
 * projected face corners when the library computes them: from_verts=0 here, 1 for iper_raster_frames); the caller
 * owns it so the library never allocates — it may be reused by the next call on the same stream.
 * ---------------------------------------------------------------------------------------------------------- */
size_t iper_raster_workspace_bytes(int B, int nf, int from_verts);
/* Rounding model of the rasteriser arithmetic, process-wide.  The upstream extension is built by nvcc, whose default
 * -fmad=true may contract a*b+c; which model the installed fork follows can only be checked against a real build of it:
 *   0 (default) every float op rounded separately — bit-exact with oracle/raster_ref.c
 *   1           fused where nvcc would contract upstream's source — bit-exact with the oracle's -DORACLE_FMA variant
 * Also selectable by the environment variable IPER_RASTER_FMA=1 (read at the first launch). */
int iper_raster_set_contraction(int mode);
int iper_raster_get_contraction(void);
int iper_rasterize_faces(const float* faces, int B, int nf, int S, float near_, float far_, int32_t* fim, float* wim,
                         void* workspace, size_t workspace_bytes, iper_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Seam B2 / engine — replaces SMPLRenderer.render_fim_wim (nmr.py:319-342: orthographic projection :34-52, y flip,
 * nr.look_at with eye=(0,0,eye_z), nr.vertices_to_faces, rasterise, f2pts) and, when tsf_inputs/Tst are given,
 * also encode_fim (nmr.py:390-401), cal_bc_transform (nmr.py:713-757), FlowComposition.make_tsf_inputs
 * (iPERCore/models/flowcomposition.py:206-248) and make_trans_flow (:514-582) for a batch of independent frames.
 *   verts (B,nv,3) f32, cams (B,3) f32 [s,tx,ty], faces (nf,3) i32
 *   optional outputs: fim (B,S,S) i32, wim (B,S,S,3) f32, f2pts (B,nf,3,2) f32
 *   fused outputs   : tsf_inputs (B,6,S,S) f32 = cat[grid_sample(uv_img, Tuv2t), cond], Tst (B,ns,S,S,2) f32
 *                     from map_fn (nf+1,3), f_uvs2img (nf,3,2), uv_img (3,S,S), src_f2pts (ns,nf,3,2)
 * ---------------------------------------------------------------------------------------------------------- */
int iper_raster_frames(const float* verts, const float* cams, const int32_t* faces, int B, int nv, int nf, int S,
                       float eye_z, float near_, float far_, int32_t* fim, float* wim, float* f2pts,
                       const float* map_fn, const float* f_uvs2img, const float* uv_img, const float* src_f2pts,
                       int ns, float* tsf_inputs, float* Tst, void* workspace, size_t workspace_bytes,
                       iper_stream_t stream);

/* SMPLRenderer.cal_bc_transform (nmr.py:713-757).  T (nb,nsrc,S,S,2): item (b,s) combines fim/wim[b] with
 * f2pts[s] (f2pts_per_item=0, shape (nsrc,nf,3,2)) or f2pts[b] (f2pts_per_item=1, nsrc must be 1). */
int iper_flow_from_fim_wim(const float* f2pts, int f2pts_per_item, const int32_t* fim, const float* wim, int nb,
                           int nsrc, int nf, int S, float* T, iper_stream_t stream);

/* SMPLRenderer.get_vis_f2pts (nmr.py:639-681): faces that appear in fim — except the smallest unique value of the map, which
 * `fim.unique()[1:]` drops: the background -1 when present, else the lowest visible face id — plus their top_k UV-nearest
 * neighbours (face_k_nearest (nf, top_k) int64, the reference buffer) keep their corner coordinates, all other faces become -2.
 * f2pts / out (B, nf, elems_per_face) f32 (elems 6 = (3,2) or 9 = (3,3)); fim (B,S,S) i32.  workspace: byte flags,
 * iper_vis_f2pts_workspace_bytes(B, nf).  No host synchronisation (the reference sorts twice per item with torch.unique). */
size_t iper_vis_f2pts_workspace_bytes(int B, int nf);
int iper_vis_f2pts(const float* f2pts, int elems_per_face, const int32_t* fim, const int64_t* face_k_nearest, int top_k, int B,
                   int nf, int S, float* out, void* workspace, size_t workspace_bytes, iper_stream_t stream);

/* SMPLRenderer.encode_fim (nmr.py:390-401): out = map_fn[fim] (fim == -1 -> row nf); transpose -> (nb,ch,S,S). */
int iper_encode_fim(const int32_t* fim, const float* map_fn, int nb, int nf, int ch, int S, int transpose, float* out,
                    iper_stream_t stream);

/* LWB.resize_trans (iPERCore/models/networks/generators/attlwb_spade_resunet.py:175-182):
 * bilinear align_corners=True resize of a flow field (n,S,S,2) -> (n,h,w,2). */
int iper_flow_resize(const float* T, int n, int S, int h, int w, float* out, iper_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Seam B3 — building blocks of the AttLWB-SPADE generator (attlwb_spade_resunet.py).  Activations live in HBM
 * as NHWC "planes"; every *_planes argument is the FORMAT of the tensor:
 *   1: fp16(x)                      2: fp16(x) + fp16(x - hi)  (split fp16, ~22 bits, 3 MMAs per K step)
 *   3: fp16(x) + two e4m3 planes stored where the lo plane would be: a8 = e4m3(8x), l8 = e4m3(2^14 (x - hi));
 *      the conv does hi*w_hi in kind::f16 and the two cross terms in kind::f8f6f4 (2 MMA-equivalents per K step).
 * A plane holds *_plane_stride elements; 8-bit planes start 2 and 3 plane_strides (in bytes) after the base.
 * ---------------------------------------------------------------------------------------------------------- */
enum { IPER_CONV_S1 = 0,    /* k x k, stride 1, pad k/2 (k = 1, 3, 5)                                     */
       IPER_CONV_S2 = 1,    /* 3x3, stride 2, pad 1                                                        */
       IPER_CONVT_4S2 = 2,  /* ConvTranspose2d(k=4, s=2, p=1) as four 2x2 phase convolutions               */
       IPER_CONV_ROW5 = 3 };/* 5x5 s1 p2 heads: K = (dy, cin), rows = (dx, out) pairs, epilogue shift-add     */

enum { IPER_EPI_PLANES = 0, /* out = [relu](acc + bias [+ residual x]) -> NHWC fp16 planes                 */
       IPER_EPI_F32 = 1,    /* out = acc + bias -> NHWC fp32                                               */
       IPER_EPI_SPADE = 2,  /* rows = [gamma block | beta block]; out = IN(x)*(1+gamma)+beta -> planes     */
       IPER_EPI_HEADS = 3 };/* (IPER_CONV_ROW5) out 0..2 tanh image, out 3 sigmoid mask, + composite with bg  */

typedef struct {
    /* A operand: input activations, NHWC fp16 planes */
    const void* a; int a_planes; long long a_plane_stride;   /* stride in elements                         */
    int N, H, W;                                             /* input spatial dims                          */
    int a_pitch, a_coff, Cin;                                /* channels per pixel in memory, window start  */
    int mode, ksize;
    /* B operand: packed weights fp16 [plane][phase][rows][taps*Cin], K ordered (tap, cin), K-major          */
    const void* w; int w_planes; long long w_plane_stride;
    int rows;                                                /* GEMM N per phase (multiple of block_n)      */
    int block_n;                                             /* 64, 128 or 256 (32 for the heads)           */
    /* epilogue */
    int epi; const float* bias; int relu;
    void* out; int out_planes; long long out_plane_stride; int out_pitch, out_coff;
    const void* x; int x_planes; long long x_plane_stride; int x_pitch, x_coff; /* residual / SPADE input   */
    const float* mean_rstd;                                  /* (N, C, 2) instance-norm statistics of x     */
    int spade_C;
    /* IPER_EPI_HEADS: NCHW fp32 outputs (N,3,H,W), (N,1,H,W), (N,3,H,W); bg (.,3,H,W) with batch stride     */
    const float* bg; long long bg_batch_stride; float* img; float* mask; float* pred;
    int max_ctas;                                            /* 0 = one persistent CTA per SM               */
    /* planes format 3 (fp16 + e4m3 cross terms): e4m3 weights [phase][rows][K]: w8 = e4m3(w*sW), wl8 =
     * e4m3((w - fp16(w))*sW*2^11); cross_scale = 1 / (2^14 * sW) turns the fp8 accumulator into the fp32 sum        */
    const void* w8; const void* wl8; float cross_scale;
    int tiles_m;                                             /* 128-pixel M tiles per CTA: 0 = auto, 1 or 2   */
    /* IPER_EPI_PLANES only: fused instance-norm statistics of the stored output — (N, rows, 2) fp64 workspace of
     * (sum, sum of squares), cleared by the call; finish with iper_instnorm_finalize                              */
    double* stats_ws;
    /* 0: one CTA per 128-pixel tile.  1: 2-CTA clusters (cta_group::2), weight tile split across the pair (formats 1/2,
     * block_n 128/256).  2: CTA pair + vertical halo — 16x8-pixel tiles whose vertical taps are views of ONE TMA box per
     * horizontal tap (3x3 stride-1, block_n >= 64; 5x5 heads as IPER_CONV_ROW5 with 32x4 tiles), and the transposed
     * conv runs its four phases fused from the same boxes (block_n 64); needs a map of >= 16x8 (heads 32x4) pixels.  */
    int cta_pair;
    /* Optional DEVICE scalar 1/s (formats 1/2).  fp16 planes have fp16 RANGE: the lo plane of a weight below ~1e-3 is subnormal
     * and the split degrades towards a single fp16.  Packing the weights as w * s with s a power of two (e.g. max|w| * s in
     * [128, 256)) and passing 1/s here keeps both planes normal; the epilogue multiplies the accumulator by 1/s before bias —
     * exact (power of two), so results are bit-identical to s = 1 wherever s = 1 was not losing bits.  NULL = 1. */
    const float* w_scale_inv;
} iper_conv_gemm_desc;

/* tcgen05/TMEM implicit-GEMM convolution with TMA im2col tile loads (conv_tc.cu). */
int iper_conv_gemm(const iper_conv_gemm_desc* d, iper_stream_t stream);

/* Host-only introspection of the cta_pair = 2 tap program for a layer (mode IPER_CONV_S1 3x3 / IPER_CONVT_4S2 /
 * IPER_CONV_ROW5; fuse_n = 1: the fused-N form of the transposed conv, several phases per MMA group):
 * out = {n_loads, acc_blocks, box_rows, phase of block 0..3, per load {ox, oy, first, count}, per entry {a_row_off, b_row,
 * b_k, acc, nblk, fb_row[2][2], fb_k[2][2]}}; returns the number of ints written (capacity >= 227) or -1.  No device work. */
int iper_conv_halo_plan(int mode, int Cin, int rows, int fuse_n, int32_t* out, int capacity);

/* CUDA-core direct convolution on the same operand formats — the on-device cross-check of iper_conv_gemm and
 * the path for tensor-core-hostile shapes.  Same descriptor; `w` is ignored, weights come as fp32 in the
 * reference's own layout: Conv2d (Cout,Cin,k,k) / ConvTranspose2d (Cin,Cout,4,4), values = hi(+lo) of planes. */
int iper_conv_direct(const iper_conv_gemm_desc* d, const float* w_f32, int Cout, iper_stream_t stream);

/* Stem: Conv2d(Cin<=8 -> Cout, 3x3, s2, p1)(+bias)+ReLU from an NCHW fp32 image to NHWC planes
 * (tsf_net_enc.layers.0 / src_net.encoders.layers.0, attlwb_spade_resunet.py:268-271). */
int iper_conv_stem(const float* in_nchw, int N, int Cin, int H, int W, const float* w_f32, const float* bias, int Cout,
                   void* out, int out_planes, long long out_plane_stride, int out_pitch, int out_coff,
                   double* stats_ws /* optional (N,Cout,2) fp64 sums, cleared by the call */, iper_stream_t stream);

/* The same stem as a tensor-core GEMM: im2col of the 3x3 / s2 / p1 patches of (N,Cin<=7,H,W) fp32 into NHWC planes
 * (N,H/2,W/2,64), channel k = (ky*3+kx)*Cin + ci, zero-padded to 64; the stem is then iper_conv_gemm(IPER_CONV_S1, ksize 1,
 * Cin 64) on weights packed (Cout, 64) in the same (tap, ci) order. */
int iper_stem_im2col(const float* in_nchw, int N, int Cin, int H, int W, void* out, int out_planes,
                     long long out_plane_stride, int out_pitch, int out_coff, iper_stream_t stream);

/* The stem on the tensor cores WITHOUT the im2col round trip (default): builder warps stage the image patch of every 16x8 output
 * tile in shared memory and write the [128 px x 64] split-fp16 A operand directly in the swizzled layout UMMA reads; weights
 * w_packed = (w_planes, Cout = 64, 64) fp16 planes with K = (ky*3+kx)*Cin + ci zero-padded to 64 (and optionally pre-scaled, see
 * iper_conv_gemm_desc.w_scale_inv); epilogue = bias + ReLU + planes store + optional instance-norm sums (stats_ws (N,64,2) fp64). */
int iper_conv_stem_tc(const float* in_nchw, int N, int Cin, int H, int W, const void* w_packed, int w_planes,
                      long long w_plane_stride, const float* w_scale_inv, const float* bias, void* out, int out_planes,
                      long long out_plane_stride, int out_pitch, int out_coff, double* stats_ws, iper_stream_t stream);

/* nn.InstanceNorm2d(affine=False) statistics (attlwb_spade_resunet.py:62, eps 1e-5, biased variance):
 * mean_rstd (N,C,2) = (mean, 1/sqrt(var+eps)) of an NHWC planes tensor.  workspace: N*C*2 doubles (fp64 sums,
 * cleared and filled by the call; two launches + one memset on the stream). */
int iper_instnorm_stats(const void* x, int x_planes, long long x_plane_stride, int N, int HW, int C, int x_pitch,
                        int x_coff, float eps, double* workspace, float* mean_rstd, iper_stream_t stream);

/* (sum, sum of squares) fp64 workspace (N*C*2) -> mean_rstd (N,C,2) fp32: mean, 1/sqrt(var_biased + eps). */
int iper_instnorm_finalize(const double* workspace, int N, int C, int HW, float eps, float* mean_rstd, iper_stream_t stream);

/* InstanceNorm2d(affine=False) apply [+ReLU] [+ residual] on NHWC planes — the BGNet blocks (bg_inpaintor.py:13-21,
 * 33-52): out = [res +] [relu]((x - mean) * rstd). */
int iper_instnorm_apply(const void* x, int x_planes, long long x_plane_stride, int x_pitch, int x_coff,
                        const float* mean_rstd, int N, int HW, int C, int relu, const void* res, int res_planes,
                        long long res_plane_stride, int res_pitch, int res_coff, void* out, int out_planes,
                        long long out_plane_stride, int out_pitch, int out_coff, iper_stream_t stream);

/* tanh + NHWC fp32 (pitch channels per pixel, first C used) -> NCHW fp32: last layer of BGNet (bg_inpaintor.py:54-55). */
int iper_tanh_nhwc_to_nchw(const float* in, int N, int HW, int C, int pitch, float* out, iper_stream_t stream);

/* Flow-guided warp + per-pixel source attention (LWB.transform :184-191, SelfAttentionBlock :102-139) with the 1x1
 * projections hoisted to the source side.  With q = Wq x_t + bq and K_s = warp(Wk x_s) + bk,
 *   K_s . q = warp((Wq^T Wk) x_s) . x_t + warp((Wk^T bq) . x_s) + bk . q,   and bk . q cancels in softmax over s.
 * kv (ns,h,w,2C+64) fp32 per source pixel: [ (Wq^T Wk) x_s | Wv x_s | (Wk^T bq) . x_s | pad ];  xt = target features
 * (B,h,w,C) planes;  T (B,ns,h,w,2) flow at this resolution;  out (B,h,w,C) planes = sum_s softmax_s(.) (warp(Wv x_s) + bv). */
int iper_warp_attention(const void* xt, int xt_planes, long long xt_plane_stride, int xt_pitch, int xt_coff,
                        const float* kv, const float* bias_v, const float* T, int B, int ns, int h, int w, int C,
                        void* out, int out_planes, long long out_plane_stride, int out_pitch, int out_coff,
                        iper_stream_t stream);

/* LWB.transform alone: grid_sample(src (ns,h,w,C) fp32 NHWC, T (B,ns,h,w,2)) -> (B,ns,h,w,C) fp32 (seam/debug). */
int iper_warp_nhwc(const float* src, const float* T, int B, int ns, int h, int w, int C, float* out,
                   iper_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * SMPL / SMPLH linear blend skinning — iPERCore/tools/human_digitalizer/smplx/lbs.py:137-227,
 * bodynets/batch_smplh.py:137-180, base_smpl.py:28-50 (link).
 * iper_lbs_shape (once per shape): v_shaped (nv,3) = v_template (+offsets) + shapedirs(nv,3,nb) . betas ;
 *                                  J_rest (nj,3) = J_regressor (nj,nv) @ v_shaped
 * iper_lbs_frames (per batch)    : pose (B,nj*3) axis-angle -> verts (B,nv,3) [+ posed joints (B,nj,3)];
 *     posedirs ((nj-1)*9, nv*3), weights_t (nj,nv) = lbs_weights^T, src_of (nv) or NULL = cloth-link source vertex of
 *     every output vertex; pose_feature (B,(nj-1)*9) and A (B,nj,12) are caller-provided scratch.
 * ---------------------------------------------------------------------------------------------------------- */
int iper_lbs_shape(const float* v_template, const float* offsets, const float* shapedirs, const float* betas,
                   const float* J_regressor, int nv, int nb, int nj, float* v_shaped, float* J_rest, iper_stream_t stream);
int iper_lbs_frames(const float* pose, int B, int nj, const float* v_shaped, const float* J_rest, const int32_t* parents,
                    const float* posedirs, const float* weights_t, const int32_t* src_of, int nv, float* pose_feature,
                    float* A, float* joints, float* verts, iper_stream_t stream);

/* layout converters between the reference's NCHW fp32 tensors and NHWC planes */
int iper_nchw_to_planes(const float* in, int N, int C, int HW, void* out, int out_planes, long long out_plane_stride,
                        int out_pitch, int out_coff, iper_stream_t stream);
int iper_planes_to_nchw(const void* x, int x_planes, long long x_plane_stride, int N, int C, int HW, int x_pitch,
                        int x_coff, float* out, iper_stream_t stream);
int iper_nhwc_f32_to_nchw(const float* in, int N, int C, int HW, int pitch, int coff, float* out, iper_stream_t stream);

/* (B,3,S,S) fp32 in [-1,1] -> (B,S,S,3) uint8 BGR, the conversion of cv_utils.save_cv2_img(normalize=True)
 * (iPERCore/tools/utils/filesio/cv_utils.py:100-116) done on device so only 0.75 MB/frame crosses PCIe. */
int iper_pred_to_u8(const float* pred, int B, int S, uint8_t* out, iper_stream_t stream);

/* Mask morphology of source_setup — iPERCore/tools/utils/morphology/morph_ops.py: morph() :7-36 (mode 0 erode: border 1,
 * box sum == ks*ks; mode 1 dilate: border 0, box sum >= 1) and soft_dilate() :39-61 (mode 2: border 0, sum >= ks*ks/2).
 * mask, out (N,1,H,W) f32; ks odd, <= 63 (deploy.toml uses 3..51).  Exact for 0/1 masks. */
int iper_morph(const float* mask, int N, int H, int W, int ks, int mode, float* out, iper_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Stand-alone generator handle — runs the AttLWB-SPADE per-frame network from C alone: weight repacking, layer graph and
 * workspace planning live in the library (csrc/generator.cu), so a non-Python host needs nothing but this header.
 * Replaces NetworksFactory.get_by_name("AttLWB-SPADE", ...) + load_state_dict (iPERCore/models/networks/__init__.py:14-16,
 * iPERCore/models/imitator.py:155-175) and BaseAttentionLWBGenerator.forward_src(only_enc=True) / forward_tsf
 * (attlwb_spade_resunet.py:450-535) + the composite of Imitator.forward (imitator.py:384-395).
 *
 *   iper_gen_create(num_filters[3] = {64,128,256}, n_res_block, precision 1 (fp16) | 2 (split fp16), mixed_spade, &gen)
 *   iper_gen_load_weight(gen, "tsf_net_enc.layers.0.0.weight", dev_ptr, shape, ndim)   for every state_dict tensor: reference
 *        fp32 layouts on the DEVICE, kept by pointer until iper_gen_pack; a "module." prefix is stripped (base_model.py:56-65);
 *        tensors the hot path does not use (bg_net.*, src_net.decoders.*, src_net.*_reg) are accepted and ignored
 *   iper_gen_packed_bytes(gen) -> size of the packed-weights buffer (0 + iper_last_error() while tensors are missing)
 *   iper_gen_pack(gen, packed, bytes, stream)        repack on the device into the caller's 256-byte aligned buffer
 *   iper_gen_src_cache_bytes / iper_gen_src_workspace_bytes / iper_gen_tsf_workspace_bytes -> caller-owned buffers
 *   iper_gen_forward_src(gen, src_inputs (ns,6,S,S) f32, ns, S, src_cache, ..)        once per source set
 *   iper_gen_forward_tsf(gen, tsf_inputs (B,6,S,S), Tst (B,ns,S,S,2), src_cache, ns, B, S, bg ((1|B),3,S,S) or NULL,
 *        bg_batched, img (B,3,S,S), mask (B,1,S,S), pred (B,3,S,S) [any may be NULL], workspace, ..)   per batch of frames
 * The library never allocates device memory and never synchronises; the handle is a host object.  Not thread-safe per handle.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct iper_gen iper_gen;
int iper_gen_create(const int* num_filters, int n_res_block, int precision, int mixed_spade, iper_gen** out);
void iper_gen_destroy(iper_gen* gen);
int iper_gen_load_weight(iper_gen* gen, const char* name, const float* dev_ptr, const int64_t* shape, int ndim);
size_t iper_gen_packed_bytes(iper_gen* gen);
int iper_gen_pack(iper_gen* gen, void* packed, size_t packed_bytes, iper_stream_t stream);
size_t iper_gen_src_cache_bytes(iper_gen* gen, int ns, int S);
size_t iper_gen_src_workspace_bytes(iper_gen* gen, int ns, int S);
size_t iper_gen_tsf_workspace_bytes(iper_gen* gen, int ns, int B, int S);
int iper_gen_forward_src(iper_gen* gen, const float* src_inputs, int ns, int S, void* src_cache, size_t src_cache_bytes,
                         void* workspace, size_t workspace_bytes, iper_stream_t stream);
int iper_gen_forward_tsf(iper_gen* gen, const float* tsf_inputs, const float* Tst, const void* src_cache, int ns, int B, int S,
                         const float* bg, int bg_batched, float* img, float* mask, float* pred, void* workspace,
                         size_t workspace_bytes, iper_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * One-time-per-source kernels of Imitator.source_setup (iPERCore/models/imitator.py:177-246 ->
 * FlowComposition.process_source, iPERCore/models/flowcomposition.py:452-512).
 *
 * iper_canny_edges — CannyFilter.forward(img, low, high, hysteresis=True) for a 1-channel map
 *   (iPERCore/tools/utils/morphology/canny_ops.py:129-192): gaussian 3x3 -> sobel x/y -> magnitude + orientation ->
 *   directional non-maximum suppression -> double threshold -> hysteresis; edges (N,H,W) in {0,1} = its `thin_edges`.
 *   gaussian_w[9], sobel_x_w[9] (sobel_y = transpose), directional_w[8*9] are HOST arrays (the module's constant
 *   filter taps, passed by value to the kernels); mag_ws/tri_ws (N*H*W f32) and ori_ws (N*H*W bytes) are device scratch.
 * iper_morph_image — make_morph_image / cal_top_k_ids / morph_image (flowcomposition.py:264-386) with top_k = 3:
 *   uncertain pixels (outpad_sil*(1-confidant_sil) != 0) take sum_k (d_k / sum d) * src[nn_k] over their 3 nearest edge
 *   pixels (squared pixel distance; ties -> lowest row-major index), every other pixel src * confidant_sil.
 *   src_img/out (N,3,H,W), sils/edges (N,H,W); count_ws (N) and list_ws (N*H*W) int32 device scratch.
 * iper_uv_warp + iper_uv_merge — make_uv_img (flowcomposition.py:87-137): per source, T = cal_bc_transform(f2pts, uv_fim,
 *   uv_wim) fused with grid_sample(src, T) -> src_warp (N,3,H,W) and grid_sample(ones, T_vis) -> vis_warp (N,H,W); the caller
 *   dilates vis_warp (iper_morph, ks 13) and iper_uv_merge blends the primary source with the others -> uv_img (bs,3,H,W).
 *   uv_fim (H,W) i32 / uv_wim (H,W,3): the constant UV-layout maps of render_uv_fim_wim; f2pts (N,nf,3,2), N = bs*ns.
 * ---------------------------------------------------------------------------------------------------------- */
int iper_canny_edges(const float* img, int N, int H, int W, const float* gaussian_w, const float* sobel_x_w,
                     const float* directional_w, float hysteresis_w, float low, float high, float* mag_ws, int8_t* ori_ws,
                     float* tri_ws, float* edges, iper_stream_t stream);
int iper_morph_image(const float* src_img, const float* confidant_sil, const float* outpad_sil, const float* edges, int N,
                     int H, int W, int32_t* count_ws, int32_t* list_ws, float* out, iper_stream_t stream);
int iper_uv_warp(const float* src_img, const float* f2pts, const float* vis_f2pts, const int32_t* uv_fim, const float* uv_wim,
                 int N, int nf, int H, int W, float* src_warp, float* vis_warp, iper_stream_t stream);
int iper_uv_merge(const float* src_warp, const float* vis_dilated, int bs, int ns, int H, int W, float* uv_img,
                  iper_stream_t stream);

/* ----------------------------------------------------------------------------------------------------------
 * Training-step kernels (BASELINE.json configs[4]): the stride-1 "same" convolutions of the step in bf16 on
 * the tensor cores — forward, data gradient, weight gradient, bias gradient — and the fused Adam + weight-repack pass.  The
 * reference trains these layers through cuDNN and torch.optim.Adam (iPERCore/tools/trainers/lwg_trainer.py:326-352, 699-833 on
 * attlwb_spade_resunet.py:14-25, 80-93, 208-252, 316-357, 605-613; bg_inpaintor.py:24-60).  Everything is NHWC bf16 (= torch
 * channels_last); channel counts are multiples of 64 (callers zero-pad the 1/3/4/6-channel ends).
 *   iper_conv_bf16        y (N,Ho,Wo,Cout) = conv_kxk(x (N,H,W,Cin), stride 1 | 2, padding pad; k <= 7) with w_packed (Cout, k*k*Cin)
 *                         bf16, K = (ky*k+kx)*Cin + ci  [+ bias fp32] [+ add_nhwc, a residual of the output's shape] [ReLU].  Stride 2
 *                         reads x through a 5-D parity view (even H, W).  dgrad of a stride-1 conv = the same call on dY with padding
 *                         k-1-pad and the weights rotated by 180 degrees, in/out transposed: w'(ci, (k-1-ky, k-1-kx), co) = w(co, ci, ky, kx);
 *                         dgrad of ConvTranspose2d(k, 2, pad) = this call with stride 2 on dY and the plain (ci_T, tap, co_T) packing.
 *   iper_conv_transposed_bf16  y (N,2H,2W,Cout) = the stride-2 transposition: ConvTranspose2d(4, 2, 1) forward, and the data gradient
 *                         of a stride-2 conv (k = 3 or 4, pad 1), as four stride-1 phase convolutions over the input grid with
 *                         interleaved stores; w_phases = the four K-major blocks (rows Cout, K = (tap in phase, ci)), phases ordered
 *                         (py,px) = (0,0) (0,1) (1,0) (1,1), taps of a phase = the (ky,kx) with (ky+pad+py), (kx+pad+px) even, ascending.
 *   iper_conv_wgrad_bf16  dW[co*stride_co + ci*stride_ci + tap*stride_tap] += sum over the pixels p of dY: dY[p, co] * X[stride*p + tap - pad, ci]
 *                         (fp32 atomics: the caller zeroes dW, or lets several calls accumulate).  Reads both operands in NHWC
 *                         (MN-major tcgen05 operands; no transposed or shifted copies).  Only co < co_valid, ci < ci_valid are
 *                         written (zero-padded ends).  (co, tap, ci) layout: strides (k*k*Cin, 1, Cin) — the fast one: a warp's
 *                         atomics then fall on consecutive floats; the parameter's own (co, ci, ky, kx) layout: strides
 *                         (Cin_real*k*k, k*k, 1).
 *   iper_bias_grad_bf16   db[c] += sum over pixels dY[p, c], c < C (fp32 atomics), `pitch` channels per pixel.
 *   iper_adam_pack        torch.optim.Adam's update (betas, eps, no weight decay; bias correction from the device-side step
 *                         counter *step_dev, so the call can be captured in a CUDA graph) over flat fp32 buffers, described by a
 *                         device-resident segment table + chunk table ((segment, first element) pairs of <= chunk_elems elements),
 *                         gradients pre-multiplied by grad_scale (1 / world size); for segments with taps > 0 (convolution weights
 *                         (co, ci, ky, kx), whose GRADIENT is read in the (co, tap, ci) layout iper_conv_wgrad_bf16 writes fastest)
 *                         it also writes the bf16 forward packing (co_pad, taps, ci_pad) at fwd_offset and, when
 *                         dgrad_offset >= 0, the dgrad packing (ci_pad, taps reversed, co_pad) — padding entries are never written
 *                         (the caller zeroes the pack buffers once).  update = 0 only repacks.
 * Every read tensor must hold one 16 x 8 TMA box: H / stride >= 8, W / stride >= 16.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct iper_adam_seg {
    long long offset, numel;        /* position in the flat buffers */
    int co, ci, taps;               /* 4-D weight (co, ci, ky, kx) with taps = k*k (a transposed convolution's (ci_T, co_T, ky, kx) tensor:
                                       co = ci_T, ci = co_T); taps = 0: not a packed weight */
    int co_pad, ci_pad;             /* channel counts of the packings (multiples of 64) */
    int reserved;                   /* kind | k << 8 | pad << 16; kind 1: stride-1 conv, 2: stride-2 conv, 3: ConvTranspose2d(k, 2, pad) */
    long long fwd_offset, dgrad_offset;   /* element offsets into pack_fwd / pack_dgrad (dgrad_offset < 0: none) */
} iper_adam_seg;

int iper_conv_bf16(const void* x_nhwc, int N, int H, int W, int Cin, const void* w_packed, int Cout, int ksize, int stride, int pad,
                   const float* bias, int relu, const void* add_nhwc, void* out_nhwc, iper_stream_t stream);
int iper_conv_transposed_bf16(const void* x_nhwc, int N, int H, int W, int Cin, const void* w_phases, int Cout, int ksize, int pad,
                              const float* bias, int relu, void* out_nhwc, iper_stream_t stream);
int iper_conv_wgrad_bf16(const void* x_nhwc, const void* dy_nhwc, int N, int H, int W, int Cin, int Cout, int ksize, int stride, int pad,
                         float* dW, long long stride_co, long long stride_ci, long long stride_tap, int co_valid, int ci_valid,
                         iper_stream_t stream);
int iper_bias_grad_bf16(const void* dy_nhwc, long long pixels, int C, int pitch, float* db, iper_stream_t stream);
int iper_adam_pack(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const iper_adam_seg* segs_dev,
                   const int* chunks_dev, int n_chunks, int chunk_elems, float lr, float beta1, float beta2, float eps,
                   float grad_scale, const float* step_dev, int update, void* pack_fwd, void* pack_dgrad, iper_stream_t stream);

/* ----------------------------------------------------------------------------------------------------------
 * Training-step kernels, part 2: the non-GEMM pieces of SelfAttentionLWB and the instance norms, forward and backward, NHWC bf16
 * (csrc/train_ops.cu).  Reference: attlwb_spade_resunet.py:80-93 (SPADE), 121-139 + 232-240 (softmax over sources), 184-191
 * (LWB.transform), bg_inpaintor.py / patch_dis.py (InstanceNorm2d + ReLU / LeakyReLU), run by ATen in the reference.
 *   iper_warp_bf16             out (M,h,w,C) = grid_sample(src (M,h,w,C), T (M,h,w,2) fp32; bilinear, zeros, align_corners=False)
 *   iper_warp_bwd_bf16         dsrc (M,h,w,C) FP32 (zeroed by the call) += scatter of dout through the same taps
 *   iper_att_combine_bf16      alpha (bs,ns,HW) fp32 = softmax_s( k_s . q / sqrt(C) ),  a (bs,HW,C) = sum_s alpha_s v_s;
 *                              k, v (bs*ns,HW,C), q (bs,HW,C); C in {64,128,256}, ns <= 8
 *   iper_att_combine_bwd_bf16  dk, dv (bs*ns,HW,C), dq (bs,HW,C) from da
 *   iper_norm_stats_bf16       stats (N,C,2) double = per-(n,c) sum and sum of squares over the HW pixels (zeroed by the call)
 *   iper_norm_apply_bf16       y = act( IN(x) * (1 + gamma) + beta ) (gamma = beta = NULL: plain instance norm); act 0 none,
 *                              1 ReLU / LeakyReLU with `slope` for the negative side (0 = ReLU)
 *   iper_norm_bwd_bf16         dx [, dgamma, dbeta] from dout (y = the saved output, needed when act != 0); sums_ws (N,C,2) double
 * C %% 8 == 0 everywhere; the norm kernels need C/8 to divide 256 (C = 64, 128, 256, 512).
 * ---------------------------------------------------------------------------------------------------------- */
/* (N,C,H,W) fp32 (is_bf16 = 0) or bf16 with arbitrary element strides -> dense NHWC bf16, channels zero-padded to Cpad (multiple of 8):
 * the cast + pad + channels_last copy of the 1/3/4/6-channel ends of the training step in one pass. */
int iper_pad_nhwc_bf16(const void* x, int is_bf16, int N, int C, int H, int W, long long stride_n, long long stride_c, long long stride_h,
                       long long stride_w, int Cpad, void* out_nhwc, iper_stream_t stream);
/* Weight gradient of the THIN ends (one side has <= 4 real channels: the 5x5 heads, the 7x7 BGNet ends), stride 1, on CUDA cores:
 *   sign +1, thin = dY (N,H,W,thin_pitch; Ct real), wide = X (N,H,W,Cw):  dW[ct*stride_thin + tap*stride_tap + cw*stride_wide] += sum_p dY[p,ct] X[p + tap - pad, cw]
 *   sign -1, thin = X, wide = dY:                                          dW[...] += sum_q dY[q - (tap - pad), cw] X[q, ct]
 * (the tensor-core wgrad streams a 64-channel zero-padded operand for these layers: 410-470 us per launch at 512x512). */
int iper_thin_wgrad_bf16(const void* wide_nhwc, const void* thin_nhwc, int N, int H, int W, int Cw, int thin_pitch, int Ct, int ksize,
                         int pad, int sign, float* dW, long long stride_thin, long long stride_wide, long long stride_tap,
                         iper_stream_t stream);
int iper_warp_bf16(const void* src_nhwc, const float* T, int M, int h, int w, int C, void* out_nhwc, iper_stream_t stream);
int iper_warp_bwd_bf16(const void* dout_nhwc, const float* T, int M, int h, int w, int C, float* dsrc_f32, iper_stream_t stream);
int iper_att_combine_bf16(const void* k, const void* v, const void* q, int bs, int ns, long long HW, int C, void* a, float* alpha,
                          iper_stream_t stream);
int iper_att_combine_bwd_bf16(const void* da, const void* k, const void* v, const void* q, const float* alpha, int bs, int ns,
                              long long HW, int C, void* dk, void* dv, void* dq, iper_stream_t stream);
int iper_norm_stats_bf16(const void* x_nhwc, int N, long long HW, int C, double* stats, iper_stream_t stream);
int iper_norm_apply_bf16(const void* x_nhwc, const double* stats, const void* gamma, const void* beta, int N, long long HW, int C,
                         float eps, int act, float slope, void* y_nhwc, iper_stream_t stream);
int iper_norm_bwd_bf16(const void* dout, const void* x, const void* y, const void* gamma, const double* stats, int N, long long HW,
                       int C, float eps, int act, float slope, double* sums_ws, void* dgamma, void* dbeta, void* dx,
                       iper_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IPER_B200_H */
