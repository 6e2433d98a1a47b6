#!/usr/bin/env python
"""CPU emulation of per-layer operand precision for the AttLWB-SPADE generator (design tool, not shipped code).

The tcgen05 convs take fp16 operands; a layer can run
  "x3"  activations and weights as hi+lo fp16 planes, 3 MMAs per K step  (~22-bit operands, error ~1e-7 relative)
  "x2a" split activations x fp16 weights, 2 MMAs          (weight rounding only)
  "x2w" fp16 activations x split weights, 2 MMAs          (activation rounding only)
  "x1"  fp16 x fp16, 1 MMA
This script evaluates a PLAN (layer group -> mode) against the exact fp32 oracle on the golden inputs and prints the
max-abs error of (tsf_img, tsf_mask) per case, so the plan shipped in ipercore_b200/generator.py (PRECISION_PLANS) can be
chosen with a known margin to BASELINE.json's 1e-3 tolerance.  `kv16` additionally stores the per-source attention maps
in fp16 (what iper_warp_attention gathers).

    python tools/precision_plan.py                 # the candidate plans on S=256 / S=512 goldens + two more seeds
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)

from oracle import generator_ref as G, weights  # noqa: E402

GROUPS = ("enc", "res", "spade_shared", "spade_gb", "convT", "skip", "heads")


def group_of(name):
    if "mlp_shared" in name:
        return "spade_shared"
    if "mlp_gamma" in name or "mlp_beta" in name:
        return "spade_gb"
    if name.startswith("tsf_net_enc"):
        return "enc"
    if name.startswith("res_blocks"):
        return "res"
    if "upconvs" in name:
        return "convT"
    if "skippers" in name:
        return "skip"
    if "img_reg" in name or "att_reg" in name:
        return "heads"
    return None          # fq/fk/fv (hoisted, exact here) and src_net (one-time, always x3)


def h(t):
    return t.half().float()


def split(t):            # value of the hi+lo planes
    hi = t.half().float()
    return hi + (t - hi).half().float()


def operands(mode, x, w):
    if mode == "x3":
        return split(x), split(w)
    if mode == "x2a":
        return split(x), h(w)
    if mode == "x2w":
        return h(x), split(w)
    if mode == "x1":
        return h(x), h(w)
    raise ValueError(mode)


class Emu:
    def __init__(self, plan, kv16=False, stem="exact"):
        self.plan, self.kv16 = plan, kv16

    def conv(self, sd, name, x, stride=1, padding=0):
        g = group_of(name)
        if g is None or (g == "enc" and name.endswith("layers.0.0")):      # the K=54 stem runs on CUDA cores in fp32
            return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)
        xq, wq = operands(self.plan.get(g, "x3"), x, sd[name + ".weight"])
        return F.conv2d(xq, wq, sd.get(name + ".bias"), stride=stride, padding=padding)

    def convT(self, sd, name, x):
        xq, wq = operands(self.plan.get("convT", "x3"), x, sd[name + ".weight"])
        return F.conv_transpose2d(xq, wq, sd.get(name + ".bias"), stride=2, padding=1)


def forward_tsf_emulated(sd, tsf_inputs, src_enc, src_res, Tst, plan, kv16):
    emu = Emu(plan, kv16)
    orig = (G._conv, G._convT, G.self_attention_lwb)
    G._conv, G._convT = emu.conv, emu.convT
    if kv16:
        def att(sd_, prefix, tsf_x, src_x, T):
            # the engine's algebra: K_s.q = warp((Wq^T Wk) x_s).x_t + warp((Wk^T bq).x_s) (+ const), V_s = warp(Wv x_s) + bv,
            # with the three source maps stored in fp16
            bs, ns, H, W, _ = T.shape
            C = tsf_x.shape[1]
            Wq, Wk, Wv = (sd_[prefix + k + ".weight"].reshape(C, C) for k in (".fq", ".fk", ".fv"))
            bq, bv = sd_[prefix + ".fq.bias"], sd_[prefix + ".fv.bias"]
            M = (Wq.double().t() @ Wk.double()).float()
            kmap = h(torch.einsum("oc,nchw->nohw", M, src_x))
            vmap = h(torch.einsum("oc,nchw->nohw", Wv, src_x))
            k0 = h(torch.einsum("c,nchw->nhw", (Wk.double().t() @ bq.double()).float(), src_x))[:, None]
            Tr = T.reshape(bs * ns, H, W, 2)
            wk = G.lwb_transform(kmap, Tr); wv = G.lwb_transform(vmap, Tr) + bv[None, :, None, None]
            w0 = G.lwb_transform(k0, Tr)
            hh, ww = tsf_x.shape[-2:]
            logits = ((wk.view(bs, ns, C, hh, ww) * tsf_x[:, None]).sum(2, keepdim=True) + w0.view(bs, ns, 1, hh, ww)) / np.sqrt(C)
            alpha = torch.softmax(logits, dim=1)
            xatt = (alpha * wv.view(bs, ns, C, hh, ww)).sum(1)
            return G.spade(sd_, prefix + ".spade", tsf_x, xatt)
        G.self_attention_lwb = att
    try:
        with torch.no_grad():
            return G.forward_tsf(sd, tsf_inputs, src_enc, src_res, Tst)
    finally:
        G._conv, G._convT, G.self_attention_lwb = orig


PLANS = {
    "fp16x2 (all x3)": {},
    "fp16 (all x1)": {g: "x1" for g in GROUPS},
    "spade x1": {"spade_shared": "x1", "spade_gb": "x1"},
    "spade_gb x1": {"spade_gb": "x1"},
    "spade_shared x1": {"spade_shared": "x1"},
    "spade_gb x1 + shared x2a": {"spade_gb": "x1", "spade_shared": "x2a"},
    "spade_gb x1 + shared x2w": {"spade_gb": "x1", "spade_shared": "x2w"},
    "spade x2a": {"spade_shared": "x2a", "spade_gb": "x2a"},
    "spade x2w": {"spade_shared": "x2w", "spade_gb": "x2w"},
    "spade x1 + res x2a": {"spade_shared": "x1", "spade_gb": "x1", "res": "x2a"},
    "spade x1 + res x2w": {"spade_shared": "x1", "spade_gb": "x1", "res": "x2w"},
    "res x1": {"res": "x1"},
    "res x2a": {"res": "x2a"},
    "res x2w": {"res": "x2w"},
    "heads x1": {"heads": "x1"},
    "convT x1": {"convT": "x1"},
    "skip x1": {"skip": "x1"},
    "enc x1": {"enc": "x1"},
}


def cases(sizes=(256, 512), extra_seeds=(7, 23)):
    import make_golden
    from oracle import flow_ref, synth
    out = []
    for S in sizes:
        out.append(("S%d golden" % S, 0, make_golden.gen_inputs(S)))
    tpl = synth.load_template()
    for sd_seed in extra_seeds:        # other weights, other pose, real raster-derived tsf_inputs
        S = 256
        cams, verts = synth.pose_sweep(tpl, 1, total=11, start=sd_seed % 11)
        scams, sverts = synth.source_views(tpl, 2)
        src_f2pts, sfim, _ = flow_ref.render_fim_wim(scams, sverts, tpl["faces"], S)
        fi = flow_ref.frame_inputs(cams, verts, tpl["faces"], tpl["map_fn"], tpl["f_uvs2img"],
                                   synth.smooth_image((1, 3, S, S), seed=sd_seed), src_f2pts, S)
        src_inputs = np.concatenate([synth.smooth_image((2, 3, S, S), seed=sd_seed + 1),
                                     flow_ref.encode_fim(sfim, tpl["map_fn"])], 1)[None]
        out.append(("S256 weights seed %d, rasterised pose" % sd_seed, sd_seed,
                    dict(src_inputs=src_inputs, tsf_inputs=fi["tsf_inputs"], Tst=fi["Tst"])))
    return out


def main():
    torch.set_num_threads(int(os.environ.get("OMP_NUM_THREADS", "8")))
    only = sys.argv[1:]
    cs = cases()
    prepared = []
    for tag, seed, inp in cs:
        sd = weights.synth_state_dict(seed)
        t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in inp.items()}
        with torch.no_grad():
            se, sr = G.forward_src(sd, t["src_inputs"])
            img, mask = G.forward_tsf(sd, t["tsf_inputs"], se, sr, t["Tst"])
        prepared.append((tag, sd, t, se, sr, img, mask))
    print("%-34s %s" % ("plan", "  ".join("%-22s" % c[0][:22] for c in cs)))
    for name, plan in PLANS.items():
        if only and not any(o in name for o in only):
            continue
        for kv16 in (False, True):
            errs = []
            t0 = time.time()
            for tag, sd, t, se, sr, img, mask in prepared:
                i2, m2 = forward_tsf_emulated(sd, t["tsf_inputs"], se, sr, t["Tst"], plan, kv16)
                errs.append(max(float((i2 - img).abs().max()), float((m2 - mask).abs().max())))
            print("%-34s %s   (worst %.2e, margin %.1fx, %.0fs)" % (name + (" +kv16" if kv16 else ""),
                  "  ".join("%-22.2e" % e for e in errs), max(errs), 1e-3 / max(errs), time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
