// Stand-alone C-ABI generator: the AttLWB-SPADE per-frame network (forward_src cache + forward_tsf) driven entirely from C —
// weight repacking, layer graph, workspace planning — so a non-Python host can run the hot path from include/iper_b200.h
// alone (iper_gen_create / load_weights / forward_src / forward_tsf / destroy).
//
// Mirrors BaseAttentionLWBGenerator.forward_src(only_enc=True) / forward_tsf
// (iPERCore/models/networks/generators/attlwb_spade_resunet.py:450-535) with the algebraic restructuring described in
// DESIGN.md §5 (source-side q/k/v projections hoisted into the per-source cache).  The library still never allocates DEVICE
// memory: packed weights, the per-source cache and the activation workspace are caller-owned buffers whose sizes the
// handle reports; the handle itself is a small host object.  Every launch goes through the same extern "C" entry points
// a host could call one by one (iper_conv_gemm, iper_warp_attention, ...).
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "iper_b200.h"

namespace iper {

// ------------------------------------------------------------------------------------------------------------
// weight repacking on the device: reference fp32 layouts -> K-major fp16 planes (hi [+ lo]) the tcgen05 kernels read.
// One generic gather kernel; `mode` selects how (row, k) of the packed matrix maps to the source tensor.
// ------------------------------------------------------------------------------------------------------------
enum PackMode { PK_CONV = 0, PK_CONVT = 1, PK_SPADE = 2, PK_HEADS = 3, PK_STEM = 4, PK_MATRIX = 5 };
struct PackArgs {
    int mode;
    const float* src;    // PK_CONV/STEM: (Cout,Cin,k,k); PK_CONVT: (Cin,Cout,4,4); PK_MATRIX: (rows,K) row-major fp32
    const float* src2;   // PK_SPADE: beta weight (gamma in src); PK_HEADS: att_reg weight (img_reg in src)
    int rows, K;         // packed matrix dims
    int Cout, Cin, ks;   // source dims
    int cb;              // PK_SPADE: channels per tile half
    __half* hi; __half* lo;   // lo may be null (single plane)
    // power-of-two pre-scale (iper_conv_gemm_desc.w_scale_inv): slot = {max|w| bits, s, 1/s}; measure = 1: only fill the max
    uint32_t* slot; int measure;
};
// s = 2^floor(log2(256 / max|w|)): max|w| * s in [128, 256); exact via frexp
__global__ void pack_scale_kernel(uint32_t* slot) {
    const float mx = __uint_as_float(slot[0]);
    float s = 1.f;
    if (mx > 0.f && mx < 3.0e38f) {
        int e;
        const float m = frexpf(mx, &e);             // mx = m * 2^e, m in [0.5, 1)
        s = ldexpf(1.f, (m == 0.5f ? 9 : 8) - e);
    }
    reinterpret_cast<float*>(slot)[1] = s;
    reinterpret_cast<float*>(slot)[2] = 1.f / s;
}
__global__ void pack_kernel(PackArgs a) {
    const size_t total = (size_t)a.rows * a.K;
    const float scale = a.measure ? 1.f : reinterpret_cast<const float*>(a.slot)[1];
    float local_max = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / a.K), k = (int)(i % a.K);
        float v = 0.f;
        if (a.mode == PK_CONV || a.mode == PK_STEM) {
            // m[co][(ky*ks+kx)*Cin + ci] = w[co][ci][ky][kx]; PK_STEM pads K beyond ks*ks*Cin, PK_CONV may pad rows beyond Cout
            const int kk = a.ks * a.ks * a.Cin;
            if (r < a.Cout && k < kk) {
                const int tap = k / a.Cin, ci = k - tap * a.Cin;
                v = a.src[(((size_t)r * a.Cin + ci) * a.ks + tap / a.ks) * a.ks + tap % a.ks];
            }
        } else if (a.mode == PK_CONVT) {
            // rows = (phase p = py*2+px, co); K = (ta*2+tb, ci); ky = {py=0:(1,3), py=1:(0,2)}[ta], same for kx
            const int p = r / a.Cout, co = r - p * a.Cout, py = p >> 1, px = p & 1;
            const int t = k / a.Cin, ci = k - t * a.Cin, ta = t >> 1, tb = t & 1;
            const int ky = py == 0 ? (ta == 0 ? 1 : 3) : (ta == 0 ? 0 : 2), kx = px == 0 ? (tb == 0 ? 1 : 3) : (tb == 0 ? 0 : 2);
            v = a.src[(((size_t)ci * a.Cout + co) * 4 + ky) * 4 + kx];
        } else if (a.mode == PK_SPADE) {
            // tiles of 2*cb rows: [gamma channels t*cb..] then [beta channels t*cb..]; K = (tap, ci) of a (C,Cin,3,3) conv
            const int t = r / (2 * a.cb), q = r - t * 2 * a.cb;
            const int c = t * a.cb + (q < a.cb ? q : q - a.cb);
            const float* w = q < a.cb ? a.src : a.src2;
            const int tap = k / a.Cin, ci = k - tap * a.Cin;
            v = w[(((size_t)c * a.Cin + ci) * 3 + tap / 3) * 3 + tap % 3];
        } else if (a.mode == PK_HEADS) {
            // row = dx*4 + o (o = r,g,b from img_reg, 3 = mask from att_reg; rows 20..31 zero); K = (dy, c)
            if (r < 20) {
                const int dx = r >> 2, o = r & 3, dy = k / a.Cin, c = k - dy * a.Cin;
                const float* w = o < 3 ? a.src + (size_t)o * a.Cin * 25 : a.src2;
                v = w[((size_t)c * 5 + dy) * 5 + dx];
            }
        } else {   // PK_MATRIX
            v = a.src[i];
        }
        if (a.measure) { local_max = fmaxf(local_max, fabsf(v)); continue; }
        v *= scale;
        const __half h = __float2half_rn(v);
        a.hi[i] = h;
        if (a.lo) a.lo[i] = __float2half_rn(v - __half2float(h));
    }
    if (a.measure && local_max > 0.f) atomicMax(a.slot, __float_as_uint(local_max));     // non-negative floats order like their bits
}
// spade bias in the same tile interleave: [gamma bias cb | beta bias cb] per tile
__global__ void spade_bias_kernel(const float* __restrict__ bg, const float* __restrict__ bb, int C, int cb, float* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= 2 * C) return;
    const int t = r / (2 * cb), q = r - t * 2 * cb, c = t * cb + (q < cb ? q : q - cb);
    out[r] = q < cb ? bg[c] : bb[c];
}
// source-side attention projection: rows [ Wq^T Wk (C) | Wv (C) | Wk^T bq (1) | zeros (63) ], fp64 accumulate
__global__ void att_source_matrix_kernel(const float* __restrict__ wq, const float* __restrict__ bq, const float* __restrict__ wk,
                                         const float* __restrict__ wv, int C, float* __restrict__ out) {
    const size_t total = (size_t)(2 * C + 64) * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / C), j = (int)(i % C);
        double acc = 0.0;
        if (r < C) {
            for (int o = 0; o < C; o++) acc += (double)wq[(size_t)o * C + r] * (double)wk[(size_t)o * C + j];
        } else if (r < 2 * C) {
            acc = wv[(size_t)(r - C) * C + j];
        } else if (r == 2 * C) {
            for (int o = 0; o < C; o++) acc += (double)wk[(size_t)o * C + j] * (double)bq[o];
        }
        out[i] = (float)acc;
    }
}

struct PlanesT {           // NHWC planes tensor (ops.py: Planes)
    void* p; int fmt; int N, H, W, pitch, C, coff;
    long long plane_stride() const { return (long long)N * H * W * pitch; }
    size_t bytes() const { return (size_t)(fmt == 1 ? 1 : 2) * N * H * W * pitch * sizeof(__half); }
    PlanesT window(int off, int c) const { PlanesT w = *this; w.coff = coff + off; w.C = c; return w; }
};

struct Packed {            // one packed weight matrix inside the caller's packed-weights buffer
    size_t off = 0; int fmt = 0, rows = 0, K = 0;
    size_t slot_off = (size_t)-1;     // {max bits, s, 1/s} of the power-of-two pre-scale, or -1
    size_t bias_off = (size_t)-1;     // fp32 bias inside the packed buffer (copied / interleaved), or -1
};

struct Bump {              // dry-run capable bump allocator over a caller-owned buffer
    uint8_t* base; size_t cap, off = 0; bool dry;
    size_t take_off(size_t n) {            // 256-byte aligned offset of a new block of n bytes
        off = (off + 255) & ~(size_t)255;
        const size_t o = off;
        off += n;
        return o;
    }
    void* take(size_t n) {
        const size_t o = take_off(n);
        return dry ? nullptr : base + o;
    }
};

}  // namespace iper

using namespace iper;

struct iper_gen {
    int nf[3], n_res, precision, mixed;
    std::map<std::string, const float*> w;          // reference tensors by state_dict name (device pointers, caller-owned)
    std::map<std::string, std::vector<int64_t>> shape;
    std::map<std::string, Packed> pk;
    uint8_t* packed = nullptr; size_t packed_bytes = 0; bool packed_ok = false;
    int fmt_of(const char* group) const {            // per-group operand format of the precision plan
        if (mixed && (!strcmp(group, "spade_shared") || !strcmp(group, "spade_gb"))) return 1;
        return precision;
    }
};

static int bn_for(int rows) { return rows >= 256 ? 256 : rows; }

extern "C" int iper_gen_create(const int* num_filters, int n_res_block, int precision, int mixed_spade, iper_gen** out) {
    IPER_REQUIRE(out && num_filters, "iper_gen_create: null pointer");
    IPER_REQUIRE(num_filters[0] == 64 && num_filters[1] == 128 && num_filters[2] == 256,
                 "iper_gen_create: kernel tiling is specialised to num_filters = [64,128,256] (AttLWB-SPADE.toml)");
    IPER_REQUIRE(n_res_block >= 1 && n_res_block <= 16, "iper_gen_create: n_res_block=%d", n_res_block);
    IPER_REQUIRE(precision == 1 || precision == 2, "iper_gen_create: precision %d not in {1 = fp16, 2 = split fp16}", precision);
    iper_gen* g = new iper_gen();
    for (int i = 0; i < 3; i++) g->nf[i] = num_filters[i];
    g->n_res = n_res_block; g->precision = precision; g->mixed = (mixed_spade != 0 && precision == 2);
    *out = g;
    return 0;
}
extern "C" void iper_gen_destroy(iper_gen* g) { delete g; }

// names of the state_dict tensors forward_src / forward_tsf need (the checkpoint's other tensors may be loaded or skipped)
static std::vector<std::string> needed_names(const iper_gen* g) {
    std::vector<std::string> n;
    auto add = [&](const std::string& s, bool bias = true) { n.push_back(s + ".weight"); if (bias) n.push_back(s + ".bias"); };
    for (int i = 0; i < 3; i++) {
        add("src_net.encoders.layers." + std::to_string(i) + ".0");
        add("tsf_net_enc.layers." + std::to_string(i) + ".0", false);
        add("tsf_net_dec.upconvs." + std::to_string(i) + ".0");
    }
    for (int i = 0; i < 2; i++) add("tsf_net_dec.skippers." + std::to_string(i) + ".0");
    auto att = [&](const std::string& p) {
        add(p + ".fq"); add(p + ".fk"); add(p + ".fv");
        add(p + ".spade.mlp_shared.0"); add(p + ".spade.mlp_gamma"); add(p + ".spade.mlp_beta");
    };
    for (int i = 0; i < 3; i++) att("enc_attlwbs." + std::to_string(i));
    for (int i = 0; i < g->n_res; i++) {
        att("res_attlwbs." + std::to_string(i));
        for (const char* pre : {"src_net.res_blocks.", "res_blocks."}) {
            add(std::string(pre) + std::to_string(i) + ".main.0"); add(std::string(pre) + std::to_string(i) + ".main.2");
        }
    }
    add("tsf_img_reg.0", false); add("tsf_att_reg.0", false);
    return n;
}

extern "C" int iper_gen_load_weight(iper_gen* g, const char* name, const float* dev_ptr, const int64_t* shape, int ndim) {
    IPER_REQUIRE(g && name && dev_ptr && shape && ndim >= 1 && ndim <= 4, "iper_gen_load_weight: bad arguments");
    std::string s(name);
    if (s.rfind("module.", 0) == 0) s = s.substr(7);        // DDP prefix, stripped like base_model.py:56-65
    g->w[s] = dev_ptr;
    g->shape[s] = std::vector<int64_t>(shape, shape + ndim);
    g->packed_ok = false;
    return 0;
}

// ---- packed-weights plan: the same walk computes the size (dry) and performs the packing ---------------------------
static int pack_all(iper_gen* g, Bump& b, cudaStream_t st) {
    const int P = g->precision;
    auto need = [&](const std::string& n, std::initializer_list<int64_t> shp) -> const float* {
        auto it = g->w.find(n);
        if (it == g->w.end()) { set_last_error("iper_gen: tensor %s was not loaded", n.c_str()); return nullptr; }
        const auto& s = g->shape[n];
        if (s.size() != shp.size() || !std::equal(s.begin(), s.end(), shp.begin())) {
            set_last_error("iper_gen: tensor %s has the wrong shape", n.c_str()); return nullptr;
        }
        return it->second;
    };
    auto planes = [&](Packed& p, int fmt, int rows, int K) -> PackArgs {
        p.fmt = fmt; p.rows = rows; p.K = K;
        const size_t plane = (size_t)rows * K * sizeof(__half);
        p.off = b.take_off((fmt == 2 ? 2 : 1) * plane);
        p.slot_off = b.take_off(16);
        PackArgs a = {};
        a.rows = rows; a.K = K;
        if (!b.dry) {
            a.hi = (__half*)(b.base + p.off); a.lo = fmt == 2 ? (__half*)(b.base + p.off + plane) : nullptr;
            a.slot = (uint32_t*)(b.base + p.slot_off);
        }
        return a;
    };
    auto launch = [&](PackArgs a) -> int {          // measure max|w| -> power-of-two scale -> write the scaled planes
        if (b.dry) return 0;
        const size_t total = (size_t)a.rows * a.K;
        const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 8);
        IPER_CHECK_CUDA(cudaMemsetAsync(a.slot, 0, 16, st));
        a.measure = 1;
        pack_kernel<<<grid, 256, 0, st>>>(a);
        pack_scale_kernel<<<1, 1, 0, st>>>(a.slot);
        a.measure = 0;
        pack_kernel<<<grid, 256, 0, st>>>(a);
        IPER_CHECK_CUDA(cudaGetLastError());
        return 0;
    };
    auto bias_copy = [&](Packed& p, const float* src, int n) -> int {
        p.bias_off = b.take_off(sizeof(float) * n);
        if (!b.dry) IPER_CHECK_CUDA(cudaMemcpyAsync(b.base + p.bias_off, src, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
        return 0;
    };
    auto conv = [&](const std::string& name, int Cout, int Cin, int ks, bool bias, int fmt) -> int {
        const float* w = need(name + ".weight", {Cout, Cin, ks, ks});
        if (!w) return 1;
        Packed& p = g->pk[name];
        PackArgs a = planes(p, fmt, Cout, ks * ks * Cin);
        a.mode = PK_CONV; a.src = w; a.Cout = Cout; a.Cin = Cin; a.ks = ks;
        if (int rc = launch(a)) return rc;
        if (bias) {
            const float* bv = need(name + ".bias", {Cout});
            if (!bv) return 1;
            if (int rc = bias_copy(p, bv, Cout)) return rc;
        }
        return 0;
    };
    auto stem = [&](const std::string& name, bool bias) -> int {
        const float* w = need(name + ".weight", {64, 6, 3, 3});
        if (!w) return 1;
        Packed& p = g->pk[name];
        PackArgs a = planes(p, P, 64, 64);
        a.mode = PK_STEM; a.src = w; a.Cout = 64; a.Cin = 6; a.ks = 3;
        if (int rc = launch(a)) return rc;
        if (bias) {
            const float* bv = need(name + ".bias", {64});
            if (!bv) return 1;
            if (int rc = bias_copy(p, bv, 64)) return rc;
        }
        return 0;
    };
    auto convT = [&](const std::string& name, int Cin, int Cout) -> int {
        const float* w = need(name + ".weight", {Cin, Cout, 4, 4});
        const float* bv = need(name + ".bias", {Cout});
        if (!w || !bv) return 1;
        Packed& p = g->pk[name];
        PackArgs a = planes(p, P, 4 * Cout, 4 * Cin);
        a.mode = PK_CONVT; a.src = w; a.Cout = Cout; a.Cin = Cin;
        if (int rc = launch(a)) return rc;
        return bias_copy(p, bv, Cout);
    };
    auto att = [&](const std::string& pre, int C) -> int {
        const float* wq = need(pre + ".fq.weight", {C, C, 1, 1}); const float* bq = need(pre + ".fq.bias", {C});
        const float* wk = need(pre + ".fk.weight", {C, C, 1, 1}); const float* wv = need(pre + ".fv.weight", {C, C, 1, 1});
        const float* bv = need(pre + ".fv.bias", {C});
        if (!wq || !bq || !wk || !wv || !bv) return 1;
        // fp32 source-projection matrix (2C+64, C) in scratch, then split into planes like any other 1x1 conv weight
        float* m = (float*)b.take(sizeof(float) * (size_t)(2 * C + 64) * C);
        Packed& p = g->pk[pre + ".kv"];
        PackArgs a = planes(p, P, 2 * C + 64, C);
        a.mode = PK_MATRIX; a.src = m;
        if (!b.dry) {
            att_source_matrix_kernel<<<148 * 4, 256, 0, st>>>(wq, bq, wk, wv, C, m);
            IPER_CHECK_CUDA(cudaGetLastError());
        }
        if (int rc = launch(a)) return rc;
        if (int rc = bias_copy(g->pk[pre + ".bv"], bv, C)) return rc;
        if (int rc = conv(pre + ".spade.mlp_shared.0", 128, C, 3, true, g->fmt_of("spade_shared"))) return rc;
        const float* wg = need(pre + ".spade.mlp_gamma.weight", {C, 128, 3, 3}); const float* bg = need(pre + ".spade.mlp_gamma.bias", {C});
        const float* wb = need(pre + ".spade.mlp_beta.weight", {C, 128, 3, 3}); const float* bb = need(pre + ".spade.mlp_beta.bias", {C});
        if (!wg || !bg || !wb || !bb) return 1;
        Packed& gb = g->pk[pre + ".spade.gb"];
        PackArgs s = planes(gb, g->fmt_of("spade_gb"), 2 * C, 9 * 128);
        s.mode = PK_SPADE; s.src = wg; s.src2 = wb; s.Cin = 128; s.cb = bn_for(2 * C) / 2;
        if (int rc = launch(s)) return rc;
        gb.bias_off = b.take_off(sizeof(float) * 2 * C);
        if (!b.dry) {
            spade_bias_kernel<<<(2 * C + 255) / 256, 256, 0, st>>>(bg, bb, C, s.cb, (float*)(b.base + gb.bias_off));
            IPER_CHECK_CUDA(cudaGetLastError());
        }
        return 0;
    };
    int rc = 0;
    if ((rc = stem("src_net.encoders.layers.0.0", true))) return rc;
    if ((rc = stem("tsf_net_enc.layers.0.0", false))) return rc;
    for (int i = 1; i < 3; i++) {
        if ((rc = conv("src_net.encoders.layers." + std::to_string(i) + ".0", g->nf[i], g->nf[i - 1], 3, true, P))) return rc;
        if ((rc = conv("tsf_net_enc.layers." + std::to_string(i) + ".0", g->nf[i], g->nf[i - 1], 3, false, P))) return rc;
    }
    for (int i = 0; i < g->n_res; i++) {
        for (const char* pre : {"src_net.res_blocks.", "res_blocks."})
            for (const char* idx : {".main.0", ".main.2"})
                if ((rc = conv(std::string(pre) + std::to_string(i) + idx, 256, 256, 3, true, P))) return rc;
        if ((rc = att("res_attlwbs." + std::to_string(i), 256))) return rc;
    }
    for (int i = 0; i < 3; i++)
        if ((rc = att("enc_attlwbs." + std::to_string(i), g->nf[i]))) return rc;
    const int up_in[3] = {256, 256, 128}, up_out[3] = {256, 128, 64};
    for (int i = 0; i < 3; i++)
        if ((rc = convT("tsf_net_dec.upconvs." + std::to_string(i) + ".0", up_in[i], up_out[i]))) return rc;
    if ((rc = conv("tsf_net_dec.skippers.0.0", 256, 384, 3, true, P))) return rc;
    if ((rc = conv("tsf_net_dec.skippers.1.0", 128, 192, 3, true, P))) return rc;
    {
        const float* wi = need("tsf_img_reg.0.weight", {3, 64, 5, 5}); const float* wm = need("tsf_att_reg.0.weight", {1, 64, 5, 5});
        if (!wi || !wm) return 1;
        PackArgs a = planes(g->pk["tsf_heads"], P, 32, 5 * 64);
        a.mode = PK_HEADS; a.src = wi; a.src2 = wm; a.Cin = 64;
        if ((rc = launch(a))) return rc;
    }
    return 0;
}

extern "C" size_t iper_gen_packed_bytes(iper_gen* g) {
    if (!g) return 0;
    // sizes do not depend on the data, but the walk validates names/shapes: report 0 (with the error set) when incomplete
    Bump b{nullptr, 0, 0, true};
    std::map<std::string, Packed> saved = g->pk;
    const int rc = pack_all(g, b, nullptr);
    g->pk = saved;
    return rc ? 0 : b.off + 256;
}

extern "C" int iper_gen_pack(iper_gen* g, void* packed, size_t packed_bytes, iper_stream_t stream) {
    IPER_REQUIRE(g && packed, "iper_gen_pack: null pointer");
    Bump b{(uint8_t*)packed, packed_bytes, 0, true};
    std::map<std::string, Packed> none;
    g->pk = none;
    if (int rc = pack_all(g, b, nullptr)) return rc;
    IPER_REQUIRE(packed_bytes >= b.off, "iper_gen_pack: buffer of %zu bytes needed (iper_gen_packed_bytes), got %zu", b.off, packed_bytes);
    IPER_REQUIRE(((uintptr_t)packed & 255) == 0, "iper_gen_pack: the packed-weights buffer must be 256-byte aligned");
    g->pk = none;
    Bump r{(uint8_t*)packed, packed_bytes, 0, false};
    if (int rc = pack_all(g, r, (cudaStream_t)stream)) return rc;
    g->packed = (uint8_t*)packed; g->packed_bytes = packed_bytes; g->packed_ok = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// layer helpers (thin fillers of iper_conv_gemm_desc, like ops._fill_desc)
// ------------------------------------------------------------------------------------------------------------
namespace {
struct Ctx {
    iper_gen* g; Bump ws; cudaStream_t st; int halo = 1, pair = 1;
    PlanesT planes(int fmt, int N, int H, int W, int C) {
        PlanesT t{nullptr, fmt, N, H, W, C, C, 0};
        t.p = ws.take(t.bytes());
        return t;
    }
    const Packed& P(const std::string& n) { return g->pk.at(n); }
    const void* wptr(const Packed& p) { return g->packed + p.off; }
    const float* bias(const Packed& p) { return p.bias_off == (size_t)-1 ? nullptr : (const float*)(g->packed + p.bias_off); }
};
int cta_pair_for(const Ctx& c, int rows, int mode, int ksize, const PlanesT& a) {       // generator.py: _pair()
    if (!c.pair) return 0;
    if (c.halo && a.H >= 8 && a.W >= 16) {
        const int bn = bn_for(rows);
        if ((mode == IPER_CONV_S1 && ksize == 3 && bn >= 64) || (mode == IPER_CONVT_4S2 && bn == 64)) return 2;
        if (mode == IPER_CONV_ROW5 && a.W >= 32) return 2;
    }
    return (rows >= 128 && mode != IPER_CONV_ROW5) ? 1 : 0;
}
void fill(iper_conv_gemm_desc& d, Ctx& c, const PlanesT& a, const Packed& w, int mode, int ksize, int rows, int block_n, int epi) {
    memset(&d, 0, sizeof(d));
    d.a = a.p; d.a_planes = a.fmt; d.a_plane_stride = a.plane_stride();
    d.N = a.N; d.H = a.H; d.W = a.W; d.a_pitch = a.pitch; d.a_coff = a.coff; d.Cin = a.C;
    d.mode = mode; d.ksize = ksize;
    d.w = c.wptr(w); d.w_planes = w.fmt; d.w_plane_stride = (long long)w.rows * w.K;
    if (w.slot_off != (size_t)-1) d.w_scale_inv = reinterpret_cast<const float*>(c.g->packed + w.slot_off) + 2;
    d.rows = rows; d.block_n = block_n; d.epi = epi;
}
void set_out(iper_conv_gemm_desc& d, const PlanesT& o) {
    d.out = o.p; d.out_planes = o.fmt; d.out_plane_stride = o.plane_stride(); d.out_pitch = o.pitch; d.out_coff = o.coff;
}
void set_x(iper_conv_gemm_desc& d, const PlanesT& x) {
    d.x = x.p; d.x_planes = x.fmt; d.x_plane_stride = x.plane_stride(); d.x_pitch = x.pitch; d.x_coff = x.coff;
}
// plain conv / transposed conv with the planes epilogue
int conv(Ctx& c, const std::string& name, const PlanesT& a, int mode, int ksize, const PlanesT& out, bool relu,
         const PlanesT* resid = nullptr, double* stats = nullptr, int force_pair = -1) {
    const Packed& w = c.P(name);
    const int rows = w.rows / (mode == IPER_CONVT_4S2 ? 4 : 1);
    if (c.ws.dry) return 0;
    iper_conv_gemm_desc d;
    fill(d, c, a, w, mode, ksize, rows, bn_for(rows), IPER_EPI_PLANES);
    d.bias = c.bias(w); d.relu = relu ? 1 : 0; d.stats_ws = stats;
    set_out(d, out);
    if (resid) set_x(d, *resid);
    d.cta_pair = force_pair >= 0 ? force_pair : cta_pair_for(c, rows, mode, ksize, a);
    return iper_conv_gemm(&d, (iper_stream_t)c.st);
}
int stem(Ctx& c, const std::string& name, const float* in_nchw, int N, int S, const PlanesT& out, double* stats) {
    const Packed& w = c.P(name);
    if (c.ws.dry) return 0;
    const float* wsi = w.slot_off != (size_t)-1 ? reinterpret_cast<const float*>(c.g->packed + w.slot_off) + 2 : nullptr;
    return iper_conv_stem_tc(in_nchw, N, 6, S, S, c.wptr(w), w.fmt, (long long)w.rows * w.K, wsi, c.bias(w), out.p, out.fmt,
                             out.plane_stride(), out.pitch, out.coff, stats, (iper_stream_t)c.st);
}
}  // namespace

// per-source cache layout: 3 + n_res stages, stage s = fp32 (ns, h, w, 2C+64) source maps
static void stage_dims(const iper_gen* g, int stage, int S, int& h, int& C) {
    if (stage < 3) { h = S >> (stage + 1); C = g->nf[stage]; } else { h = S >> 3; C = g->nf[2]; }
}
extern "C" size_t iper_gen_src_cache_bytes(iper_gen* g, int ns, int S) {
    if (!g || ns <= 0 || S <= 0) return 0;
    size_t tot = 0;
    for (int s = 0; s < 3 + g->n_res; s++) {
        int h, C; stage_dims(g, s, S, h, C);
        tot += (((size_t)ns * h * h * (2 * C + 64) * sizeof(float)) + 255) & ~(size_t)255;
    }
    return tot;
}
static float* stage_ptr(const iper_gen* g, void* cache, int stage, int ns, int S) {
    size_t off = 0;
    for (int s = 0; s < stage; s++) {
        int h, C; stage_dims(g, s, S, h, C);
        off += (((size_t)ns * h * h * (2 * C + 64) * sizeof(float)) + 255) & ~(size_t)255;
    }
    return (float*)((uint8_t*)cache + off);
}

static int run_forward_src(iper_gen* g, Ctx& c, const float* src_inputs, int ns, int S, void* cache) {
    const int P = g->precision;
    std::vector<PlanesT> feats;
    PlanesT x = c.planes(P, ns, S / 2, S / 2, 64);
    if (int rc = stem(c, "src_net.encoders.layers.0.0", src_inputs, ns, S, x, nullptr)) return rc;
    feats.push_back(x);
    for (int i = 1; i < 3; i++) {
        PlanesT y = c.planes(P, ns, x.H / 2, x.W / 2, g->nf[i]);
        if (int rc = conv(c, "src_net.encoders.layers." + std::to_string(i) + ".0", x, IPER_CONV_S2, 3, y, true)) return rc;
        x = y; feats.push_back(x);
    }
    for (int i = 0; i < g->n_res; i++) {
        PlanesT y = c.planes(P, ns, x.H, x.W, 256), z = c.planes(P, ns, x.H, x.W, 256);
        const std::string pre = "src_net.res_blocks." + std::to_string(i);
        if (int rc = conv(c, pre + ".main.0", x, IPER_CONV_S1, 3, y, true)) return rc;
        if (int rc = conv(c, pre + ".main.2", y, IPER_CONV_S1, 3, z, false, &x)) return rc;
        x = z; feats.push_back(x);
    }
    for (int s = 0; s < 3 + g->n_res; s++) {                   // source maps [(Wq^T Wk) x | Wv x | (Wk^T bq).x | pad]
        const std::string pre = s < 3 ? "enc_attlwbs." + std::to_string(s) : "res_attlwbs." + std::to_string(s - 3);
        const Packed& w = c.P(pre + ".kv");
        if (c.ws.dry) continue;
        iper_conv_gemm_desc d;
        fill(d, c, feats[s], w, IPER_CONV_S1, 1, w.rows, 64, IPER_EPI_F32);
        d.out = stage_ptr(g, cache, s, ns, S); d.out_planes = 1; d.out_pitch = w.rows;
        if (int rc = iper_conv_gemm(&d, (iper_stream_t)c.st)) return rc;
    }
    return 0;
}

static int run_forward_tsf(iper_gen* g, Ctx& c, const float* tsf_inputs, const float* Tst, const void* cache, int ns, int B, int S,
                           const float* bg, int bg_batched, float* img, float* mask, float* pred) {
    const int P = g->precision, *nf = g->nf;
    PlanesT cat0 = c.planes(P, B, S / 2, S / 2, nf[0] + nf[1]), cat1 = c.planes(P, B, S / 4, S / 4, nf[1] + nf[2]);
    // flow pyramid: one resize per resolution (LWB.resize_trans, attlwb_spade_resunet.py:175-182)
    const float* flows[3];
    for (int i = 0; i < 3; i++) {
        const int h = S >> (i + 1);
        float* f = (float*)c.ws.take(sizeof(float) * (size_t)B * ns * h * h * 2);
        flows[i] = f;
        if (!c.ws.dry)
            if (int rc = iper_flow_resize(Tst, B * ns, S, h, h, f, (iper_stream_t)c.st)) return rc;
    }
    auto att_block = [&](const std::string& pre, int stage, const PlanesT& x, const PlanesT* dst, double* stats, PlanesT& out) -> int {
        const int C = x.C, h = x.H;
        float* mr = (float*)c.ws.take(sizeof(float) * (size_t)B * C * 2);
        PlanesT a = c.planes(std::min(P, g->fmt_of("spade_shared")), B, h, h, C);
        PlanesT actv = c.planes(std::min(P, g->fmt_of("spade_gb")), B, h, h, 128);
        out = dst ? *dst : c.planes(P, B, h, h, C);
        if (c.ws.dry) return 0;
        if (int rc = iper_instnorm_finalize(stats, B, C, h * h, 1e-5f, mr, (iper_stream_t)c.st)) return rc;
        const int fi = stage < 3 ? stage : 2;
        const Packed& bv = c.P(pre + ".bv");
        if (int rc = iper_warp_attention(x.p, x.fmt, x.plane_stride(), x.pitch, x.coff, stage_ptr(g, (void*)cache, stage, ns, S),
                                         c.bias(bv), flows[fi], B, ns, h, h, C, a.p, a.fmt, a.plane_stride(), a.pitch, a.coff,
                                         (iper_stream_t)c.st)) return rc;
        if (int rc = conv(c, pre + ".spade.mlp_shared.0", a, IPER_CONV_S1, 3, actv, true)) return rc;
        const Packed& gb = c.P(pre + ".spade.gb");
        iper_conv_gemm_desc d;
        fill(d, c, actv, gb, IPER_CONV_S1, 3, 2 * C, bn_for(2 * C), IPER_EPI_SPADE);
        d.bias = c.bias(gb); d.mean_rstd = mr; d.spade_C = C;
        set_out(d, out); set_x(d, x);
        d.cta_pair = cta_pair_for(c, 2 * C, IPER_CONV_S1, 3, actv);
        return iper_conv_gemm(&d, (iper_stream_t)c.st);
    };
    auto stats_ws = [&](int C) { return (double*)c.ws.take(sizeof(double) * (size_t)B * C * 2); };

    // 1. encoder (:507-519)
    PlanesT x = c.planes(P, B, S / 2, S / 2, nf[0]);
    double* ws = stats_ws(nf[0]);
    if (int rc = stem(c, "tsf_net_enc.layers.0.0", tsf_inputs, B, S, x, ws)) return rc;
    PlanesT w0 = cat0.window(0, nf[0]), w1 = cat1.window(0, nf[1]);
    const PlanesT* enc_dst[3] = {&w0, &w1, nullptr};
    for (int i = 0; i < 3; i++) {
        if (i > 0) {
            PlanesT y = c.planes(P, B, x.H / 2, x.W / 2, nf[i]);
            ws = stats_ws(nf[i]);
            if (int rc = conv(c, "tsf_net_enc.layers." + std::to_string(i) + ".0", x, IPER_CONV_S2, 3, y, true, nullptr, ws)) return rc;
            x = y;
        }
        PlanesT o;
        if (int rc = att_block("enc_attlwbs." + std::to_string(i), i, x, enc_dst[i], ws, o)) return rc;
        x = o;
    }
    // 2. residual blocks (:522-529)
    for (int i = 0; i < g->n_res; i++) {
        PlanesT y = c.planes(P, B, x.H, x.W, 256), z = c.planes(P, B, x.H, x.W, 256);
        ws = stats_ws(256);
        const std::string pre = "res_blocks." + std::to_string(i);
        if (int rc = conv(c, pre + ".main.0", x, IPER_CONV_S1, 3, y, true)) return rc;
        if (int rc = conv(c, pre + ".main.2", y, IPER_CONV_S1, 3, z, false, &x, ws)) return rc;
        PlanesT o;
        if (int rc = att_block("res_attlwbs." + std::to_string(i), 3 + i, z, nullptr, ws, o)) return rc;
        x = o;
    }
    // 3. SkipDecoder (:348-357)
    if (int rc = conv(c, "tsf_net_dec.upconvs.0.0", x, IPER_CONVT_4S2, 4, cat1.window(nf[1], nf[2]), true)) return rc;
    PlanesT s0 = c.planes(P, B, S / 4, S / 4, nf[2]);
    if (int rc = conv(c, "tsf_net_dec.skippers.0.0", cat1, IPER_CONV_S1, 3, s0, true)) return rc;
    if (int rc = conv(c, "tsf_net_dec.upconvs.1.0", s0, IPER_CONVT_4S2, 4, cat0.window(nf[0], nf[1]), true)) return rc;
    PlanesT s1 = c.planes(P, B, S / 2, S / 2, nf[1]);
    if (int rc = conv(c, "tsf_net_dec.skippers.1.0", cat0, IPER_CONV_S1, 3, s1, true)) return rc;
    PlanesT d2 = c.planes(P, B, S, S, nf[0]);
    if (int rc = conv(c, "tsf_net_dec.upconvs.2.0", s1, IPER_CONVT_4S2, 4, d2, true)) return rc;
    // 4. heads (:533) + composite (imitator.py:393)
    if (c.ws.dry) return 0;
    const Packed& wh = c.P("tsf_heads");
    iper_conv_gemm_desc d;
    fill(d, c, d2, wh, IPER_CONV_ROW5, 5, 32, 32, IPER_EPI_HEADS);
    d.img = img; d.mask = mask; d.pred = pred; d.bg = bg; d.bg_batch_stride = bg_batched ? (long long)3 * S * S : 0;
    d.cta_pair = cta_pair_for(c, 32, IPER_CONV_ROW5, 5, d2);
    return iper_conv_gemm(&d, (iper_stream_t)c.st);
}

extern "C" size_t iper_gen_src_workspace_bytes(iper_gen* g, int ns, int S) {
    if (!g || !g->packed_ok || ns <= 0 || S <= 0 || S % 8) return 0;
    Ctx c{g, Bump{nullptr, 0, 0, true}, nullptr};
    if (run_forward_src(g, c, nullptr, ns, S, nullptr)) return 0;
    return c.ws.off + 256;
}
extern "C" size_t iper_gen_tsf_workspace_bytes(iper_gen* g, int ns, int B, int S) {
    if (!g || !g->packed_ok || ns <= 0 || B <= 0 || S <= 0 || S % 8) return 0;
    Ctx c{g, Bump{nullptr, 0, 0, true}, nullptr};
    if (run_forward_tsf(g, c, nullptr, nullptr, nullptr, ns, B, S, nullptr, 0, nullptr, nullptr, nullptr)) return 0;
    return c.ws.off + 256;
}

extern "C" int iper_gen_forward_src(iper_gen* g, const float* src_inputs, int ns, int S, void* src_cache, size_t src_cache_bytes,
                                    void* workspace, size_t workspace_bytes, iper_stream_t stream) {
    IPER_REQUIRE(g && g->packed_ok, "iper_gen_forward_src: weights are not packed (iper_gen_pack)");
    IPER_REQUIRE(src_inputs && src_cache && workspace, "iper_gen_forward_src: null pointer");
    IPER_REQUIRE(ns >= 1 && ns <= 8 && S >= 64 && S % 8 == 0, "iper_gen_forward_src: ns=%d S=%d", ns, S);
    IPER_REQUIRE(src_cache_bytes >= iper_gen_src_cache_bytes(g, ns, S), "iper_gen_forward_src: source cache too small");
    IPER_REQUIRE(workspace_bytes >= iper_gen_src_workspace_bytes(g, ns, S) && ((uintptr_t)workspace & 255) == 0,
                 "iper_gen_forward_src: workspace too small or not 256-byte aligned");
    Ctx c{g, Bump{(uint8_t*)workspace, workspace_bytes, 0, false}, (cudaStream_t)stream};
    return run_forward_src(g, c, src_inputs, ns, S, src_cache);
}

extern "C" int iper_gen_forward_tsf(iper_gen* g, const float* tsf_inputs, const float* Tst, const void* src_cache, int ns, int B,
                                    int S, const float* bg, int bg_batched, float* img, float* mask, float* pred, void* workspace,
                                    size_t workspace_bytes, iper_stream_t stream) {
    IPER_REQUIRE(g && g->packed_ok, "iper_gen_forward_tsf: weights are not packed (iper_gen_pack)");
    IPER_REQUIRE(tsf_inputs && Tst && src_cache && workspace && (img || mask || pred), "iper_gen_forward_tsf: null pointer");
    IPER_REQUIRE(!pred || bg, "iper_gen_forward_tsf: pred needs bg");
    IPER_REQUIRE(ns >= 1 && ns <= 8 && B >= 1 && S >= 64 && S % 8 == 0, "iper_gen_forward_tsf: ns=%d B=%d S=%d", ns, B, S);
    IPER_REQUIRE(workspace_bytes >= iper_gen_tsf_workspace_bytes(g, ns, B, S) && ((uintptr_t)workspace & 255) == 0,
                 "iper_gen_forward_tsf: workspace too small or not 256-byte aligned");
    Ctx c{g, Bump{(uint8_t*)workspace, workspace_bytes, 0, false}, (cudaStream_t)stream};
    return run_forward_tsf(g, c, tsf_inputs, Tst, src_cache, ns, B, S, bg, bg_batched, img, mask, pred);
}
