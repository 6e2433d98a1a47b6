"""Deterministic, torch-RNG-independent synthetic weights for the AttLWB-SPADE generator (test infrastructure).

The real checkpoint (assets/checkpoints/neural_renders/AttLWB-SPADE_id_G_2020-05-18.pth,
/root/reference/assets/configs/deploy.toml:65) is not available offline, so parity runs use weights that are a pure
function of (key name, shape, seed).  The same state_dict is loaded into the reference module (make_golden.py),
into the oracle restatement and into the CUDA engine, which also exercises the 221-key checkpoint layout
(SURVEY.md §8a "Checkpoint layout that must load unchanged").
"""
import hashlib
import re
from collections import OrderedDict

import numpy as np


def generator_param_shapes(cfg=None):
    """Ordered (name, shape) list of the reference AttentionLWBGenerator.state_dict()
    (/root/reference/iPERCore/models/networks/generators/attlwb_spade_resunet.py:538-613, bg_inpaintor.py:24-60)."""
    bg_f = [64, 128, 128, 256]
    src_f = [64, 128, 256]
    tsf_f = [64, 128, 256]
    n_res = 6
    out = []

    def conv(name, co, ci, k, bias=True):
        out.append((name + ".weight", (co, ci, k, k)))
        if bias:
            out.append((name + ".bias", (co,)))

    def convT(name, ci, co, k, bias=True):
        out.append((name + ".weight", (ci, co, k, k)))
        if bias:
            out.append((name + ".bias", (co,)))

    # bg_net (ResNetInpaintor): indices follow nn.Sequential positions (bg_inpaintor.py:30-57)
    conv("bg_net.main.0", bg_f[0], 4, 7)
    idx = 3
    for i in range(1, 4):
        conv("bg_net.main.%d" % idx, bg_f[i], bg_f[i - 1], 3)
        idx += 3
    for _ in range(n_res):
        conv("bg_net.main.%d.main.0" % idx, 256, 256, 3)
        conv("bg_net.main.%d.main.3" % idx, 256, 256, 3)
        idx += 1
    for i in range(3, 0, -1):
        convT("bg_net.main.%d" % idx, bg_f[i], bg_f[i - 1], 4, bias=False)
        idx += 3
    conv("bg_net.main.%d" % idx, 3, bg_f[0], 7, bias=False)
    # src_net (ResAutoEncoder)
    cin = 6
    for i, c in enumerate(src_f):
        conv("src_net.encoders.layers.%d.0" % i, c, cin, 3)
        cin = c
    for i in range(n_res):
        conv("src_net.res_blocks.%d.main.0" % i, 256, 256, 3)
        conv("src_net.res_blocks.%d.main.2" % i, 256, 256, 3)
    dec_f = list(reversed(src_f))
    cin = 256
    for i, c in enumerate(dec_f):
        convT("src_net.decoders.layers.%d.0" % i, cin, c, 4)
        cin = c
    conv("src_net.img_reg.0", 3, 64, 5, bias=False)
    conv("src_net.att_reg.0", 1, 64, 5, bias=False)
    # tsf_net
    cin = 6
    for i, c in enumerate(tsf_f):
        conv("tsf_net_enc.layers.%d.0" % i, c, cin, 3, bias=False)
        cin = c
    # SkipDecoder registers skippers before upconvs (attlwb_spade_resunet.py:342-343)
    dec_f = list(reversed(tsf_f))
    for i in range(2):
        s_in = tsf_f[3 - 2 - i] + dec_f[i]
        conv("tsf_net_dec.skippers.%d.0" % i, dec_f[i], s_in, 3)
    cin = 256
    for i, c in enumerate(dec_f):
        convT("tsf_net_dec.upconvs.%d.0" % i, cin, c, 4)
        cin = c

    def attlwb(prefix, c):
        conv(prefix + ".fq", c, c, 1)
        conv(prefix + ".fk", c, c, 1)
        conv(prefix + ".fv", c, c, 1)
        conv(prefix + ".spade.mlp_shared.0", 128, c, 3)
        conv(prefix + ".spade.mlp_gamma", c, 128, 3)
        conv(prefix + ".spade.mlp_beta", c, 128, 3)

    for i, c in enumerate(tsf_f):
        attlwb("enc_attlwbs.%d" % i, c)
    for i in range(n_res):
        attlwb("res_attlwbs.%d" % i, 256)
    for i in range(n_res):
        conv("res_blocks.%d.main.0" % i, 256, 256, 3)
        conv("res_blocks.%d.main.2" % i, 256, 256, 3)
    conv("tsf_img_reg.0", 3, 64, 5, bias=False)
    conv("tsf_att_reg.0", 1, 64, 5, bias=False)
    return out


def _layer_gain(name):
    """Per-layer-type gain on top of He-uniform so the random network is well conditioned (activations O(1),
    attention logits O(1), unsaturated heads) — a chaotic random net would make any tolerance meaningless."""
    if re.search(r"res_blocks\.\d+\.main\.2|bg_net\.main\.\d+\.main\.3", name):
        return 0.3                      # residual branch output
    if ".fq." in name or ".fk." in name:
        return 0.35
    if ".fv." in name:
        return 0.7
    if "mlp_gamma" in name or "mlp_beta" in name:
        return 0.35
    if "img_reg" in name or "att_reg" in name or name == "bg_net.main.27.weight":
        return 0.45
    return 1.0


def _rng_for(name, seed):
    h = hashlib.sha256(("%d:%s" % (seed, name)).encode()).digest()
    return np.random.Generator(np.random.PCG64(int.from_bytes(h[:8], "little")))


def synth_state_dict(seed=0, gain=1.0, as_torch=True):
    """Uniform(-b, b) weights with b = gain*sqrt(6/fan_in)*... chosen so activations keep O(1) scale through
    ReLU stacks (He-uniform); biases Uniform(-0.1, 0.1).  Pure function of (name, shape, seed)."""
    sd = OrderedDict()
    for name, shape in generator_param_shapes():
        rng = _rng_for(name, seed)
        if name.endswith(".weight"):
            is_T = (".decoders." in name or ".upconvs." in name
                    or name in ("bg_net.main.18.weight", "bg_net.main.21.weight", "bg_net.main.24.weight"))
            k = shape[2] * shape[3]
            # transposed conv 4x4 s2: each output sees k/4 taps of Cin=shape[0]
            fan_in = shape[0] * k // 4 if is_T else shape[1] * k
            b = gain * _layer_gain(name) * np.sqrt(6.0 / fan_in)
            w = rng.uniform(-b, b, size=shape).astype(np.float32)
        else:
            w = rng.uniform(-0.1, 0.1, size=shape).astype(np.float32)
        sd[name] = w
    if as_torch:
        import torch
        sd = OrderedDict((k, torch.from_numpy(v)) for k, v in sd.items())
    return sd
