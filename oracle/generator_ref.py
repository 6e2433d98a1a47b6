"""Plain-torch fp32 restatement of the AttLWB-SPADE generator (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Functional: takes the reference's flat ``state_dict`` (221 keys) and NCHW fp32 tensors.  Every function cites the
reference lines it follows (paths relative to /root/reference/iPERCore/models/networks/generators/).
Pinned against the reference module itself by tests/golden/make_golden.py (gen_*.npz fixtures).
"""
import math

import torch
import torch.nn.functional as F


def _conv(sd, name, x, stride=1, padding=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _convT(sd, name, x):
    return F.conv_transpose2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=2, padding=1)


def residual_block(sd, prefix, x):
    """attlwb_spade_resunet.py:14-25: x + conv3x3(ReLU(conv3x3(x))), bias, no norm."""
    y = F.relu(_conv(sd, prefix + ".main.0", x, padding=1))
    return x + _conv(sd, prefix + ".main.2", y, padding=1)


def resize_trans(T, h, w):
    """LWB.resize_trans (attlwb_spade_resunet.py:175-182): bilinear, align_corners=True, on the flow."""
    Ts = T.permute(0, 3, 1, 2)
    Ts = F.interpolate(Ts, size=(h, w), mode="bilinear", align_corners=True)
    return Ts.permute(0, 2, 3, 1)


def lwb_transform(x, T):
    """LWB.transform (attlwb_spade_resunet.py:184-191)."""
    h, w = x.shape[-2:]
    if T.shape[1] != h or T.shape[2] != w:
        T = resize_trans(T, h, w)
    return F.grid_sample(x, T, mode="bilinear", padding_mode="zeros", align_corners=False)


def attention(q, K, V):
    """SelfAttentionBlock (attlwb_spade_resunet.py:102-139): per-pixel softmax over the ns sources."""
    dk = K.shape[2]
    logits = (K * q.unsqueeze(1)).sum(dim=2, keepdim=True) / math.sqrt(dk)   # (N, ns, 1, H, W)
    alpha = torch.softmax(logits, dim=1)
    return (alpha * V).sum(dim=1)


def spade(sd, prefix, x, cond):
    """SPADE (attlwb_spade_resunet.py:80-93): IN(x) * (1 + gamma(cond)) + beta(cond)."""
    normalized = F.instance_norm(x, eps=1e-5)
    actv = F.relu(_conv(sd, prefix + ".mlp_shared.0", cond, padding=1))
    gamma = _conv(sd, prefix + ".mlp_gamma", actv, padding=1)
    beta = _conv(sd, prefix + ".mlp_beta", actv, padding=1)
    return normalized * (1 + gamma) + beta


def self_attention_lwb(sd, prefix, tsf_x, src_x, Tst, temp_x=None, Ttt=None):
    """SelfAttentionLWB.forward (attlwb_spade_resunet.py:208-252); temp_x / Ttt = the temporal branch (:230-246)."""
    bs, ns, H, W, _ = Tst.shape
    h, w = tsf_x.shape[-2:]
    src_warp = lwb_transform(src_x, Tst.reshape(bs * ns, H, W, 2))
    k = _conv(sd, prefix + ".fk", src_warp).view(bs, ns, -1, h, w)
    v = _conv(sd, prefix + ".fv", src_warp).view(bs, ns, -1, h, w)
    if temp_x is not None and Ttt is not None:
        nt = Ttt.shape[1]
        temp_warp = lwb_transform(temp_x, Ttt.reshape(bs * nt, H, W, 2))
        k = torch.cat([k, _conv(sd, prefix + ".fk", temp_warp).view(bs, nt, -1, h, w)], dim=1)
        v = torch.cat([v, _conv(sd, prefix + ".fv", temp_warp).view(bs, nt, -1, h, w)], dim=1)
    q = _conv(sd, prefix + ".fq", tsf_x)
    x = attention(q, k, v)
    return spade(sd, prefix + ".spade", tsf_x, x)


def forward_src(sd, src_inputs, n_res=6):
    """BaseAttentionLWBGenerator.forward_src(only_enc=True) (attlwb_spade_resunet.py:450-478)."""
    bs, ns, _, h, w = src_inputs.shape
    x = src_inputs.reshape(bs * ns, -1, h, w)
    enc_outs = []
    for i in range(3):
        x = F.relu(_conv(sd, "src_net.encoders.layers.%d.0" % i, x, stride=2, padding=1))
        enc_outs.append(x)
    res_outs = []
    for i in range(n_res):
        x = residual_block(sd, "src_net.res_blocks.%d" % i, x)
        res_outs.append(x)
    return enc_outs, res_outs


def forward_tsf(sd, tsf_inputs, src_enc_outs, src_res_outs, Tst, n_res=6, taps=None, temp_enc_outs=None, temp_res_outs=None,
                Ttt=None):
    """BaseAttentionLWBGenerator.forward_tsf (attlwb_spade_resunet.py:480-535); temp_* / Ttt = temporal attention inputs.

    ``taps`` (optional dict) collects intermediates for layer-wise debugging of the CUDA path."""
    x = tsf_inputs
    enc_outs = []
    for i in range(3):
        x = F.relu(_conv(sd, "tsf_net_enc.layers.%d.0" % i, x, stride=2, padding=1))
        if taps is not None:
            taps["enc%d_conv" % i] = x
        x = self_attention_lwb(sd, "enc_attlwbs.%d" % i, x, src_enc_outs[i], Tst,
                               None if temp_enc_outs is None else temp_enc_outs[i], Ttt)
        if taps is not None:
            taps["enc%d" % i] = x
        enc_outs.append(x)
    for i in range(n_res):
        x = residual_block(sd, "res_blocks.%d" % i, x)
        x = self_attention_lwb(sd, "res_attlwbs.%d" % i, x, src_res_outs[i], Tst,
                               None if temp_res_outs is None else temp_res_outs[i], Ttt)
        if taps is not None:
            taps["res%d" % i] = x
    # SkipDecoder.forward (attlwb_spade_resunet.py:348-357)
    d = x
    for i in range(3):
        d = F.relu(_convT(sd, "tsf_net_dec.upconvs.%d.0" % i, d))
        if i != 2:
            d = torch.cat([enc_outs[1 - i], d], dim=1)
            d = F.relu(_conv(sd, "tsf_net_dec.skippers.%d.0" % i, d, padding=1))
        if taps is not None:
            taps["dec%d" % i] = d
    img = torch.tanh(_conv(sd, "tsf_img_reg.0", d, padding=2))
    mask = torch.sigmoid(_conv(sd, "tsf_att_reg.0", d, padding=2))
    return img, mask


def forward_bg(sd, bg_inputs, n_res=6):
    """AttentionLWBGenerator.forward_bg -> ResNetInpaintor (bg_inpaintor.py:24-60)."""
    bs, ns, _, h, w = bg_inputs.shape
    x = bg_inputs.reshape(bs * ns, -1, h, w)
    inorm = lambda t: F.instance_norm(t, eps=1e-5)
    x = F.relu(inorm(_conv(sd, "bg_net.main.0", x, padding=3)))
    idx = 3
    for _ in range(3):
        x = F.relu(inorm(_conv(sd, "bg_net.main.%d" % idx, x, stride=2, padding=1)))
        idx += 3
    for _ in range(n_res):
        y = F.relu(inorm(_conv(sd, "bg_net.main.%d.main.0" % idx, x, padding=1)))
        x = x + inorm(_conv(sd, "bg_net.main.%d.main.3" % idx, y, padding=1))
        idx += 1
    for _ in range(3):
        x = F.relu(inorm(_convT(sd, "bg_net.main.%d" % idx, x)))
        idx += 3
    x = torch.tanh(_conv(sd, "bg_net.main.%d" % idx, x, padding=3))
    return x.view(bs, ns, 3, h, w)


def composite(tsf_img, tsf_mask, bg_img):
    """Imitator.forward (models/imitator.py:393): pred = mask*bg + (1-mask)*img."""
    return tsf_mask * bg_img + (1 - tsf_mask) * tsf_img
