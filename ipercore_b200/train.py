"""Training step of the Liquid Warping GAN on B200 (BASELINE.json configs[4]: 512x512, G + D + VGG
perceptual, bf16, NCCL gradient all-reduce) — see DESIGN.md §8b for what is and is not on hand-written kernels.

On the kernels of csrc/train.cu / csrc/train_ops.cu (everything NHWC bf16 = torch channels_last, no layout conversions):
  * every convolution of the generator, the discriminator and VGG19 — stride 1 (1x1 fq / fk / fv, 3x3, 5x5 heads, 7x7 BGNet
    ends, D's 4x4 pad-1 layers), stride 2 (encoders, D) and ConvTranspose2d(4, 2, 1) (decoders), the 1/3/4/6-channel ends
    zero-padded to 64 — forward with fused bias / residual / ReLU, data gradient, weight gradient (MN-major tcgen05 operands
    straight from NHWC) and bias gradient, the last two accumulated straight into the flat fp32 gradient buffer the all-reduce
    and the optimizer work on (``_Conv``);
  * LWB.transform (grid_sample) and its gradient (``_Warp``), the softmax over sources + weighted sum of SelfAttentionLWB
    (``_AttCombine``), InstanceNorm / SPADE modulation / ReLU / LeakyReLU with their backward (``_Norm``);
  * Adam for G and D as one pass each over flat fp32 parameter / moment buffers that also rewrites the bf16 forward and dgrad
    packings of every convolution weight (``ParamStore``);
  * the whole step (forward, both backward passes, bucketed NCCL all-reduces, both optimizer passes) replayed as ONE CUDA graph
    (``LWGTrainStep(graph=True)``): at batch 1 the step is ~1370 launches of 5-50 us each, so the graph is made wide — weight /
    bias gradients on a side stream next to the data gradient, BGNet / the VGG target features / the source decoder as branches.
Still ATen: channel-padding copies of the tiny ends, max-pooling, tanh / sigmoid, the loss reductions.

Mirrors ``LWGTrainer`` (iPERCore/tools/trainers/lwg_trainer.py: forward :699-731, optimize_G :733-795, optimize_D :797-834,
optimize_parameters :326-352) with the generator's training-shape forward
(attlwb_spade_resunet.py:633-699 ``forward(..., only_tsf=False)``, ``forward_src(only_enc=False)`` :471-478),
``PatchDiscriminator`` (discriminators/patch_dis.py:8-70 behind "patch_global", multi_scale_dis.py:47-107), ``VGGLoss``
(criterions/vggloss.py:261-292), ``LSGANLoss`` / ``TVLoss`` (ganloss.py, generals.py) and the DDP gradient all-reduce of
iPERCore/services/train.py:89-95 done as bucketed flat all-reduces overlapped with backward.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ._lib import check, lib
from .ops import _stream

BF16 = torch.bfloat16
CL = torch.channels_last


USE_KERNELS = True          # tests flip this to compare against the all-torch formulation
USE_STREAMS = True          # weight / bias gradients on a side stream, parallel to the data gradient
USE_THIN = False            # opt-in: weight gradients of the layers with <= 4 channels on one side on the CUDA-core kernel instead of
                            # zero-padded MMAs.  Correct (tests), but measured SLOWER inside the wide graph (16.4 vs 15.5 ms per step):
                            # its CUDA-core work competes with the epilogues of the branches that run next to it, whereas the
                            # tensor-core version idles the SIMT pipes.
USE_BRANCHES = True         # independent sub-networks (BGNet, VGG of the target) on their own streams = parallel branches of the graph
_SIDE, _BRANCH = {}, {}


def _side_stream(device):
    """The side stream paired with the CURRENT stream (each branch of the step has its own)."""
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(dev).cuda_stream)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _note(_SIDE[key])


def _branch_stream(device, idx):
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), idx)
    if key not in _BRANCH:
        _BRANCH[key] = torch.cuda.Stream(device=dev)
    return _note(_BRANCH[key])


_USED = []          # streams the CURRENT step has forked work onto (cleared at the top of every step; inside a capture they are all capturing)


def _note(stream):
    if all(stream != s for s in _USED):
        _USED.append(stream)
    return stream


def _all_streams(device):
    dev = torch.device(device)
    d = dev.index if dev.index is not None else torch.cuda.current_device()
    return [s for s in _USED if s.device.index == d]


class _Branch:
    """``with _Branch(device, idx) as b: y = f(x)`` runs f on branch stream idx, forked from the current stream; ``b.join(y...)`` makes
    the current stream wait for it (autograd replays the backward of those ops on the same stream, i.e. in parallel too)."""

    def __init__(self, device, idx):
        self.on = bool(USE_BRANCHES and USE_KERNELS and torch.device(device).type == "cuda")
        self.main = torch.cuda.current_stream(device) if self.on else None
        self.s = _branch_stream(device, idx) if self.on else None
        self.ctx = None

    def __enter__(self):
        if self.on:
            self.s.wait_stream(self.main)
            self.ctx = torch.cuda.stream(self.s)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.ctx.__exit__(*exc)
        return False

    def join(self, *tensors):
        if self.on:
            self.main.wait_stream(self.s)
            for t in tensors:
                t.record_stream(self.main)


def _pad64(c):
    return (c + 63) // 64 * 64


S1, S2, CT = 1, 2, 3        # kinds of convolution the kernels run: stride 1, stride 2, ConvTranspose2d(k, 2, 1)


def _kind(x, weight, stride=1, padding=None, transposed=False):
    """0 = not on the kernels, else S1 / S2 / CT.  Every tensor a kernel reads must hold one 16x8 TMA box at its (possibly halved)
    resolution; channel counts that are not multiples of 64 (the 1/3/4/6-channel ends) are zero-padded by the caller below."""
    kh, kw = weight.shape[2:]
    h, w = x.shape[-2:]
    if padding is None:
        padding = kh // 2
    if not (USE_KERNELS and x.is_cuda and kh == kw and 1 <= kh <= 7 and 0 <= padding < kh):
        return 0
    if transposed:
        return CT if (stride == 2 and kh == 4 and padding == 1 and h >= 8 and w >= 16) else 0
    if stride == 1:
        ho, wo = h + 2 * padding - kh + 1, w + 2 * padding - kw + 1
        return S1 if (h >= 8 and w >= 16 and ho >= 8 and wo >= 16) else 0            # dgrad reads dY (ho x wo)
    if stride == 2:
        return S2 if (kh in (3, 4) and padding == 1 and h % 2 == 0 and w % 2 == 0 and h >= 16 and w >= 32) else 0
    return 0


def _eligible(x, weight, stride=1, padding=None, transposed=False):
    return _kind(x, weight, stride, padding, transposed) != 0


def _phase_taps(k, pad):
    """Taps of the four output phases (py, px) of a stride-2 transposition: those with (ky + pad + py), (kx + pad + px) even."""
    out = []
    for py in (0, 1):
        for px in (0, 1):
            out.append([(ky, kx) for ky in range(k) if (ky + pad + py) % 2 == 0 for kx in range(k) if (kx + pad + px) % 2 == 0])
    return out


def pack_weight(weight, kind=S1, pad=None):
    """4-D weight -> (forward packing, dgrad packing) in bf16, K-major rows, channels zero-padded to multiples of 64 — the torch
    formulation of csrc/train.cu adam_pack_kernel's repacking (frozen weights and tests; trainable weights are repacked by the
    fused Adam pass).  With (co, ci) = the first two dims (a transposed convolution's weight is (ci_T, co_T, k, k)):
      plain  (co_pad, K = (tap, ci))                 forward of S1 / S2, data gradient of CT
      rev    (ci_pad, K = (taps-1-tap, co))          data gradient of S1
      phases 4 x (ci_pad, K = (tap in phase, co))    data gradient of S2, forward of CT"""
    co, ci, k, _ = weight.shape
    pad = k // 2 if pad is None else pad
    cop, cip = _pad64(co), _pad64(ci)
    w = weight.detach()
    plain = torch.zeros((cop, k * k, cip), dtype=BF16, device=w.device)
    plain[:co, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, k * k, ci)
    if kind == S1:
        rev = torch.zeros((cip, k * k, cop), dtype=BF16, device=w.device)
        rev[:ci, :, :co] = w.flip(2, 3).permute(1, 2, 3, 0).reshape(ci, k * k, co)
        return plain.view(cop, -1), rev.view(cip, -1)
    blocks = []
    for taps in _phase_taps(k, pad):
        b = torch.zeros((cip, len(taps), cop), dtype=BF16, device=w.device)
        for t, (ky, kx) in enumerate(taps):
            b[:ci, t, :co] = w[:, :, ky, kx].t()
        blocks.append(b.reshape(-1))
    phases = torch.cat(blocks)
    return (plain.view(cop, -1), phases) if kind == S2 else (phases, plain.view(cop, -1))


def _packs(weight, kind, pad):
    pk = getattr(weight, "_iper_pack", None)            # maintained by ParamStore (trainable) ...
    if pk is None:
        if weight.requires_grad:
            return pack_weight(weight, kind, pad)
        pk = weight._iper_pack = pack_weight(weight, kind, pad)    # ... or packed once (frozen weights: VGG)
    return pk


class _PadCL(torch.autograd.Function):
    """(N,C,H,W) fp32 | bf16, any strides -> bf16 channels_last with the channels zero-padded to cpad, in one kernel (instead of a
    cast, an F.pad and a layout copy: 45 us each on a 512x512 map).  Backward = the slice of the first C channels."""

    @staticmethod
    def forward(ctx, x, cpad):
        n, c, h, w = x.shape
        out = torch.empty((n, cpad, h, w), dtype=BF16, device=x.device, memory_format=CL)
        sn, sc, sh, sw = x.stride()
        check(lib.iper_pad_nhwc_bf16(x.data_ptr(), int(x.dtype == BF16), n, c, h, w, sn, sc, sh, sw, cpad, out.data_ptr(), _stream()), "pad_nhwc_bf16")
        ctx.c, ctx.dt = c, x.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        return g[:, :ctx.c].to(ctx.dt), None


def _to_cl(x, cpad):
    """-> bf16 channels_last with the channel count zero-padded to cpad."""
    if x.shape[1] != cpad and USE_KERNELS and x.is_cuda and x.dtype in (BF16, torch.float32):
        return _PadCL.apply(x, cpad)
    x = x.to(BF16)
    if x.shape[1] != cpad:
        x = F.pad(x, (0, 0, 0, 0, 0, cpad - x.shape[1]))
    return x.contiguous(memory_format=CL)


def _conv_call(x_cl, w_packed, cout, ksize, stride=1, pad=None, bias=None, relu=False, add=None):
    n, cin, h, w = x_cl.shape
    pad = ksize // 2 if pad is None else pad
    ho, wo = (h + 2 * pad - ksize) // stride + 1, (w + 2 * pad - ksize) // stride + 1
    y = torch.empty((n, cout, ho, wo), dtype=BF16, device=x_cl.device, memory_format=CL)
    check(lib.iper_conv_bf16(x_cl.data_ptr(), n, h, w, cin, w_packed.data_ptr(), cout, ksize, stride, pad, 0 if bias is None else bias.data_ptr(),
                             int(relu), 0 if add is None else add.data_ptr(), y.data_ptr(), _stream()), "conv_bf16")
    return y


def _convT_call(x_cl, w_phases, cout, ksize, pad, bias=None, relu=False):
    n, cin, h, w = x_cl.shape
    y = torch.empty((n, cout, 2 * h, 2 * w), dtype=BF16, device=x_cl.device, memory_format=CL)
    check(lib.iper_conv_transposed_bf16(x_cl.data_ptr(), n, h, w, cin, w_phases.data_ptr(), cout, ksize, pad,
                                        0 if bias is None else bias.data_ptr(), int(relu), y.data_ptr(), _stream()), "conv_transposed_bf16")
    return y


class _Conv(torch.autograd.Function):
    """y = [relu]( conv(x, weight, bias) [+ add] ) in bf16 on the tcgen05 kernels — stride-1 (any k <= 7 / padding), stride-2
    (k = 3 | 4, pad 1) and ConvTranspose2d(4, 2, 1); x any layout, y channels_last.  Weight / bias gradients go straight into the
    parameter's flat fp32 gradient when the parameter has a sink (ParamStore), otherwise they are returned to autograd."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, add, kind, pad):
        k = weight.shape[2]
        cin, cout = (weight.shape[0], weight.shape[1]) if kind == CT else (weight.shape[1], weight.shape[0])
        cinp, coutp = _pad64(cin), _pad64(cout)
        x_cl = _to_cl(x, cinp)
        wf, _ = _packs(weight, kind, pad)
        b = None
        if bias is not None:
            b = bias.detach().float()
            b = (F.pad(b, (0, coutp - cout)) if cout != coutp else b).contiguous()
        if kind == CT:
            y = _convT_call(x_cl, wf, coutp, k, pad, b, relu)
        else:
            a = None if add is None else _to_cl(add, coutp)
            y = _conv_call(x_cl, wf, coutp, k, 2 if kind == S2 else 1, pad, b, relu, a)
        ctx.save_for_backward(x_cl, weight, y if relu else None)
        ctx.cfg = (bias is not None, relu, add is not None, kind, pad, cin, cout)
        ctx.bias_ref = bias
        if weight.requires_grad and torch.is_grad_enabled():
            weight._iper_uses = getattr(weight, "_iper_uses", 0) + 1
        return y if cout == coutp else y[:, :cout]

    @staticmethod
    def backward(ctx, dy):
        x_cl, weight, y = ctx.saved_tensors
        has_bias, relu, has_add, kind, pad, cin, cout = ctx.cfg
        k = weight.shape[2]
        cinp, coutp = _pad64(cin), _pad64(cout)
        n, _, h, w = x_cl.shape
        dy_cl = _to_cl(dy, coutp)
        if relu:
            dy_cl = torch.ops.aten.threshold_backward(dy_cl, y, 0).contiguous(memory_format=CL)
        dx = dw = db = dadd = None
        if ctx.needs_input_grad[0]:
            _, wd = _packs(weight, kind, pad)
        want_w, want_b = ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]
        weight._iper_uses = getattr(weight, "_iper_uses", 1) - 1
        # the weight / bias gradients are independent of the data gradient: they run on a side stream (in the captured step: a
        # parallel branch of the graph) and are joined before this node returns
        main, side = torch.cuda.current_stream(), (_side_stream(x_cl.device) if USE_STREAMS and (want_w or want_b) else None)
        if side is not None:
            side.wait_stream(main)
        with torch.cuda.stream(side if side is not None else main):
            st = _stream()
            if want_w:        # gradient layout (d0, tap, d1) of the (d0, d1, k, k) tensor: a warp's atomics fall on consecutive floats
                d0, d1 = weight.shape[:2]
                sink = getattr(weight, "_iper_sink", None)
                g = sink.grad if sink is not None else torch.zeros((d0, k * k, d1), dtype=torch.float32, device=x_cl.device)
                same = kind == S1 and k % 2 == 1 and pad == k // 2
                if USE_THIN and same and cout <= 4 and cin % 64 == 0:         # thin = dY (heads, BGNet's 64 -> 3 end): CUDA-core kernel
                    check(lib.iper_thin_wgrad_bf16(x_cl.data_ptr(), dy_cl.data_ptr(), n, h, w, cin, coutp, cout, k, pad, 1, g.data_ptr(),
                                                   k * k * d1, 1, d1, st), "thin_wgrad_bf16")
                elif USE_THIN and same and cin <= 4 and cout % 64 == 0:       # thin = X (BGNet's 4 -> 64 stem)
                    check(lib.iper_thin_wgrad_bf16(dy_cl.data_ptr(), x_cl.data_ptr(), n, h, w, cout, cinp, cin, k, pad, -1, g.data_ptr(),
                                                   1, k * k * d1, d1, st), "thin_wgrad_bf16")
                elif kind == CT:  # dWt[ci, co, tap] = sum_p x[p, ci] dY[2p + tap - pad, co]: the strided read runs over dY
                    check(lib.iper_conv_wgrad_bf16(dy_cl.data_ptr(), x_cl.data_ptr(), n, 2 * h, 2 * w, coutp, cinp, k, 2, pad, g.data_ptr(),
                                                   k * k * d1, 1, d1, cin, cout, st), "conv_wgrad_bf16")
                else:
                    check(lib.iper_conv_wgrad_bf16(x_cl.data_ptr(), dy_cl.data_ptr(), n, h, w, cinp, coutp, k, 2 if kind == S2 else 1, pad,
                                                   g.data_ptr(), k * k * d1, 1, d1, cout, cin, st), "conv_wgrad_bf16")
                if sink is None:
                    g.record_stream(main)            # allocated under the side stream, consumed by autograd on the main one
                    dw = g.view(d0, k, k, d1).permute(0, 3, 1, 2).to(weight.dtype)
            if want_b:
                bias = ctx.bias_ref
                bsink = getattr(bias, "_iper_sink", None)
                gb = bsink.grad if bsink is not None else torch.zeros((cout,), dtype=torch.float32, device=x_cl.device)
                check(lib.iper_bias_grad_bf16(dy_cl.data_ptr(), dy_cl.shape[0] * dy_cl.shape[2] * dy_cl.shape[3], cout, coutp, gb.data_ptr(), st),
                      "bias_grad_bf16")
                if bsink is None:
                    gb.record_stream(main)
                    db = gb.to(bias.dtype)
        if ctx.needs_input_grad[0]:
            if kind == S1:
                dx = _conv_call(dy_cl, wd, cinp, k, 1, k - 1 - pad)
            elif kind == S2:
                dx = _convT_call(dy_cl, wd, cinp, k, pad)
            else:
                dx = _conv_call(dy_cl, wd, cinp, k, 2, pad)
            if cin != cinp:
                dx = dx[:, :cin]
        if side is not None:
            main.wait_stream(side)
        if weight._iper_uses == 0:
            if want_w and getattr(weight, "_iper_sink", None) is not None:
                weight._iper_sink.ready()
            if want_b and getattr(ctx.bias_ref, "_iper_sink", None) is not None:
                ctx.bias_ref._iper_sink.ready()
        if has_add and ctx.needs_input_grad[4]:
            dadd = dy_cl if cout == coutp else dy_cl[:, :cout]
        return dx, dw, db, None, dadd, None, None


def _torch_conv(x, weight, bias, stride=1, padding=0, transposed=False):
    """The PyTorch side of the step (strided / transposed convolutions).  bf16 channels_last; a 6-channel input (the first
    encoder layers) is zero-padded to 8 channels: cuDNN has tensor-core engines only for 8-aligned channels, and the fp32 SIMT
    weight-gradient kernel it falls back to otherwise costs 1.15 ms per layer at 512x512."""
    if not USE_KERNELS:
        tiny = min(weight.shape[0], weight.shape[1]) < 8
        dt = torch.float32 if tiny else x.dtype
        fn = F.conv_transpose2d if transposed else F.conv2d
        return fn(x.to(dt), weight.to(dt), None if bias is None else bias.to(dt), stride=stride, padding=padding)
    x, w = x.to(BF16), weight.to(BF16)
    cin_dim = 0 if transposed else 1
    if w.shape[cin_dim] % 8:
        p8 = 8 - w.shape[cin_dim] % 8
        x = F.pad(x, (0, 0, 0, 0, 0, p8))
        w = F.pad(w, (0, 0, 0, 0, 0, p8) if cin_dim == 1 else (0, 0, 0, 0, 0, 0, 0, p8))
    fn = F.conv_transpose2d if transposed else F.conv2d
    return fn(x.contiguous(memory_format=CL), w.contiguous(memory_format=CL), None if bias is None else bias.to(BF16), stride=stride, padding=padding)


def conv(x, weight, bias=None, relu=False, add=None, stride=1, padding=None, transposed=False):
    """Convolution [+ residual] [ReLU]: tcgen05 bf16 kernels (forward, dgrad, wgrad, bias grad) when the layer qualifies (see
    _kind), torch otherwise."""
    padding = weight.shape[-1] // 2 if padding is None else padding
    kind = _kind(x, weight, stride, padding, transposed)
    if kind and not (kind != S1 and add is not None):
        return _Conv.apply(x, weight, bias, relu, add, kind, padding)
    y = _torch_conv(x, weight, bias, stride=stride, padding=padding, transposed=transposed)
    if add is not None:
        y = y + add
    return F.relu(y) if relu else y


def conv3x3(x, weight, bias=None):
    return conv(x, weight, bias)


class _Warp(torch.autograd.Function):
    """LWB.transform (attlwb_spade_resunet.py:184-191): grid_sample(src, T) bilinear / zeros / align_corners=False on NHWC bf16;
    the gradient w.r.t. the source features is scattered with float4 atomics into an fp32 buffer.  T (M,h,w,2) fp32 carries none."""

    @staticmethod
    def forward(ctx, src, T):
        m, c, h, w = src.shape
        src_cl = _to_cl(src, c)
        T = T.float().contiguous()
        out = torch.empty_like(src_cl)
        check(lib.iper_warp_bf16(src_cl.data_ptr(), T.data_ptr(), m, h, w, c, out.data_ptr(), _stream()), "warp_bf16")
        ctx.save_for_backward(T)
        return out

    @staticmethod
    def backward(ctx, dout):
        (T,) = ctx.saved_tensors
        m, c, h, w = dout.shape
        d_cl = _to_cl(dout, c)
        ds = torch.empty((m, c, h, w), dtype=torch.float32, device=dout.device, memory_format=CL)
        check(lib.iper_warp_bwd_bf16(d_cl.data_ptr(), T.data_ptr(), m, h, w, c, ds.data_ptr(), _stream()), "warp_bwd_bf16")
        return ds.to(BF16), None


class _AttCombine(torch.autograd.Function):
    """a = sum_s softmax_s(K_s . q / sqrt(C)) V_s per pixel (attlwb_spade_resunet.py:121-139, 232-240); k, v (bs*ns,C,h,w), q (bs,C,h,w)."""

    @staticmethod
    def forward(ctx, k, v, q, ns):
        bs, c, h, w = q.shape
        k, v, q = _to_cl(k, c), _to_cl(v, c), _to_cl(q, c)
        a = torch.empty_like(q)
        alpha = torch.empty((bs, ns, h * w), dtype=torch.float32, device=q.device)
        check(lib.iper_att_combine_bf16(k.data_ptr(), v.data_ptr(), q.data_ptr(), bs, ns, h * w, c, a.data_ptr(), alpha.data_ptr(), _stream()),
              "att_combine_bf16")
        ctx.save_for_backward(k, v, q, alpha)
        ctx.ns = ns
        return a

    @staticmethod
    def backward(ctx, da):
        k, v, q, alpha = ctx.saved_tensors
        bs, c, h, w = q.shape
        da = _to_cl(da, c)
        dk, dv, dq = torch.empty_like(k), torch.empty_like(v), torch.empty_like(q)
        check(lib.iper_att_combine_bwd_bf16(da.data_ptr(), k.data_ptr(), v.data_ptr(), q.data_ptr(), alpha.data_ptr(), bs, ctx.ns, h * w, c,
                                            dk.data_ptr(), dv.data_ptr(), dq.data_ptr(), _stream()), "att_combine_bwd_bf16")
        return dk, dv, dq, None


class _Norm(torch.autograd.Function):
    """y = act( InstanceNorm(x) * (1 + gamma) + beta ): SPADE (attlwb_spade_resunet.py:80-93) with gamma / beta, plain
    InstanceNorm2d(affine=False) [+ ReLU / LeakyReLU] without (bg_inpaintor.py, patch_dis.py).  NHWC bf16, fp64 statistics."""

    @staticmethod
    def forward(ctx, x, gamma, beta, act, slope, eps):
        n, c, h, w = x.shape
        x = _to_cl(x, c)
        g = None if gamma is None else _to_cl(gamma, c)
        b = None if beta is None else _to_cl(beta, c)
        stats = torch.empty((n, c, 2), dtype=torch.float64, device=x.device)
        st = _stream()
        check(lib.iper_norm_stats_bf16(x.data_ptr(), n, h * w, c, stats.data_ptr(), st), "norm_stats_bf16")
        y = torch.empty_like(x)
        check(lib.iper_norm_apply_bf16(x.data_ptr(), stats.data_ptr(), 0 if g is None else g.data_ptr(), 0 if b is None else b.data_ptr(), n, h * w, c,
                                       eps, int(act), slope, y.data_ptr(), st), "norm_apply_bf16")
        ctx.save_for_backward(x, g, y if act else None, stats)
        ctx.cfg = (int(act), slope, eps)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, y, stats = ctx.saved_tensors
        act, slope, eps = ctx.cfg
        n, c, h, w = x.shape
        dy = _to_cl(dy, c)
        sums = torch.empty((n, c, 2), dtype=torch.float64, device=x.device)
        dx = torch.empty_like(x)
        dg = db = None
        if g is not None:
            dg, db = torch.empty_like(x), torch.empty_like(x)
        check(lib.iper_norm_bwd_bf16(dy.data_ptr(), x.data_ptr(), 0 if y is None else y.data_ptr(), 0 if g is None else g.data_ptr(), stats.data_ptr(),
                                     n, h * w, c, eps, act, slope, sums.data_ptr(), 0 if dg is None else dg.data_ptr(),
                                     0 if db is None else db.data_ptr(), dx.data_ptr(), _stream()), "norm_bwd_bf16")
        return dx, dg, db, None, None, None


def _norm_ok(x):
    return bool(USE_KERNELS and x.is_cuda and x.shape[1] in (64, 128, 256, 512))


def inorm(x, act=False, slope=0.0, gamma=None, beta=None, eps=1e-5):
    """act( IN(x) [* (1 + gamma) + beta] ) -> bf16."""
    if _norm_ok(x):
        return _Norm.apply(x, gamma, beta, act, slope, eps)
    y = F.instance_norm(x.float(), eps=eps)
    if gamma is not None:
        y = y * (1 + gamma.float()) + beta.float()
    if act:
        y = F.leaky_relu(y, slope) if slope else F.relu(y)
    return y.to(BF16)


# ----------------------------------------------------------------------------------------------------------------------
# generator, training shape (functional over the parameter holders of ipercore_b200.generator.AttentionLWBGenerator)
# ----------------------------------------------------------------------------------------------------------------------
class TrainableGenerator(nn.Module):
    """AttentionLWBGenerator.forward(bg, src, tsf, Tst, only_tsf=False) (attlwb_spade_resunet.py:633-699) with gradients.
    `net` holds the parameters under the reference's state_dict names (so ``personalized.pth`` saves / loads unchanged)."""

    def __init__(self, net):
        super().__init__()
        self.net = net
        for p in self.net.parameters():
            p.requires_grad_(True)
        self.n_res = net.n_res

    def _p(self, name):
        mod, attr = name.rsplit(".", 1)
        return getattr(self.net.get_submodule(mod), attr, None)

    def _c(self, name, x, stride=1, padding=0, relu=False, add=None):
        return conv(x, self._p(name + ".weight"), self._p(name + ".bias"), relu=relu, add=add, stride=stride, padding=padding)

    def _ct(self, name, x, relu=False):
        return conv(x, self._p(name + ".weight"), self._p(name + ".bias"), relu=relu, stride=2, padding=1, transposed=True)

    def _res(self, prefix, x, second):
        y = self._c("%s.main.0" % prefix, x, padding=1, relu=True)
        return self._c("%s.main.%d" % (prefix, second), y, padding=1, add=x)

    @staticmethod
    def native_weight(name, param):
        """(kind, pad) of the convolution a 4-D weight belongs to — which packings the fused Adam pass maintains for it — or None.
        Mirrors the layer calls below: stride-2 encoders (3x3, pad 1), transposed decoders (4x4, pad 1), everything else stride 1."""
        if param.dim() != 4 or param.shape[2] != param.shape[3] or not 1 <= param.shape[2] <= 7:
            return None
        k = param.shape[2]
        if any(t in name for t in ("decoders.", "upconvs.")) or any(name.endswith("bg_net.main.%d.weight" % i) for i in (18, 21, 24)):
            return (CT, 1) if k == 4 else None
        if any(t in name for t in ("encoders.", "tsf_net_enc.")) or any(name.endswith("bg_net.main.%d.weight" % i) for i in (3, 6, 9)):
            return (S2, 1) if k in (3, 4) else None
        return (S1, k // 2) if k % 2 else None

    # ---- BGNet (bg_inpaintor.py:24-60) ----
    def forward_bg(self, bg_inputs):
        bs, ns, _, h, w = bg_inputs.shape
        x = bg_inputs.reshape(bs * ns, -1, h, w).to(BF16)
        x = inorm(self._c("bg_net.main.0", x, padding=3), act=True)     # 4 -> 64 7x7 (zero-padded to 64 input channels on the kernels)
        idx = 3
        for _ in range(3):
            x = inorm(self._c("bg_net.main.%d" % idx, x, stride=2, padding=1), act=True); idx += 3
        for _ in range(self.n_res):
            y = inorm(self._c("bg_net.main.%d.main.0" % idx, x, padding=1), act=True)
            x = x + inorm(self._c("bg_net.main.%d.main.3" % idx, y, padding=1)); idx += 1
        for _ in range(3):
            x = inorm(self._ct("bg_net.main.%d" % idx, x), act=True); idx += 3
        return torch.tanh(self._c("bg_net.main.%d" % idx, x, padding=3).float()).reshape(bs, ns, 3, h, w)

    # ---- SIDNet (attlwb_spade_resunet.py:450-478, ResAutoEncoder) ----
    def forward_src(self, src_inputs, only_enc=False):
        bs, ns, _, h, w = src_inputs.shape
        x = src_inputs.reshape(bs * ns, -1, h, w).to(BF16)
        enc = []
        for i in range(3):
            x = self._c("src_net.encoders.layers.%d.0" % i, x, stride=2, padding=1, relu=True).to(BF16); enc.append(x)
        res = []
        for i in range(self.n_res):
            x = self._res("src_net.res_blocks.%d" % i, x, 2); res.append(x)
        if only_enc:
            return enc, res
        # the source decoder + heads only feed the reconstruction loss: TSFNet needs enc / res alone, so they run as a branch
        # parallel to forward_tsf (joined by the caller through self._src_branch)
        with _Branch(x.device, 3) as br:
            d = x
            for i in range(3):
                d = self._ct("src_net.decoders.layers.%d.0" % i, d, relu=True)
            img = torch.tanh(self._c("src_net.img_reg.0", d, padding=2).float()).reshape(bs, ns, 3, h, w)
            mask = torch.sigmoid(self._c("src_net.att_reg.0", d, padding=2).float()).reshape(bs, ns, 1, h, w)
        self._src_branch = (br, img, mask)
        return enc, res, img, mask

    # ---- SelfAttentionLWB (attlwb_spade_resunet.py:208-252) ----
    def _flow(self, Tst, h, w, cache):
        """LWB.resize_trans (attlwb_spade_resunet.py:175-182), once per resolution instead of once per block."""
        if (h, w) not in cache:
            bs, ns, H, W, _ = Tst.shape
            T = Tst.reshape(bs * ns, H, W, 2).float()
            if H != h or W != w:
                if USE_KERNELS and T.is_cuda and H == W:
                    from .ops import flow_resize
                    T = flow_resize(T, h, w)
                else:
                    T = F.interpolate(T.permute(0, 3, 1, 2), size=(h, w), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
            cache[(h, w)] = T.contiguous()
        return cache[(h, w)]

    def _att(self, prefix, tsf_x, src_x, Tst, cache):
        bs, ns = Tst.shape[:2]
        h, w = tsf_x.shape[-2:]
        c = tsf_x.shape[1]
        T = self._flow(Tst, h, w, cache)
        if USE_KERNELS and tsf_x.is_cuda and c in (64, 128, 256) and ns <= 8:
            warp = _Warp.apply(src_x, T)
            a = _AttCombine.apply(self._c(prefix + ".fk", warp), self._c(prefix + ".fv", warp), self._c(prefix + ".fq", tsf_x), ns)
        else:
            warp = F.grid_sample(src_x.float(), T.float(), mode="bilinear", padding_mode="zeros", align_corners=False).to(BF16)
            k = self._c(prefix + ".fk", warp).reshape(bs, ns, -1, h, w)
            v = self._c(prefix + ".fv", warp).reshape(bs, ns, -1, h, w)
            q = self._c(prefix + ".fq", tsf_x)
            logits = (k.float() * q.float().unsqueeze(1)).sum(dim=2, keepdim=True) / math.sqrt(k.shape[2])
            a = (torch.softmax(logits, dim=1) * v.float()).sum(dim=1).to(BF16)
        actv = self._c(prefix + ".spade.mlp_shared.0", a, padding=1, relu=True)
        gamma = self._c(prefix + ".spade.mlp_gamma", actv, padding=1)
        beta = self._c(prefix + ".spade.mlp_beta", actv, padding=1)
        return inorm(tsf_x, gamma=gamma, beta=beta)

    def forward_tsf(self, tsf_inputs, src_enc, src_res, Tst):
        x = tsf_inputs.to(BF16)
        enc, cache = [], {}
        for i in range(3):
            x = self._c("tsf_net_enc.layers.%d.0" % i, x, stride=2, padding=1, relu=True).to(BF16)
            x = self._att("enc_attlwbs.%d" % i, x, src_enc[i], Tst, cache); enc.append(x)
        for i in range(self.n_res):
            x = self._res("res_blocks.%d" % i, x, 2)
            x = self._att("res_attlwbs.%d" % i, x, src_res[i], Tst, cache)
        d = x
        for i in range(3):
            d = self._ct("tsf_net_dec.upconvs.%d.0" % i, d, relu=True)
            if i != 2:
                d = self._c("tsf_net_dec.skippers.%d.0" % i, torch.cat([enc[1 - i], d], dim=1), padding=1, relu=True)
        img = torch.tanh(self._c("tsf_img_reg.0", d, padding=2).float())
        mask = torch.sigmoid(self._c("tsf_att_reg.0", d, padding=2).float())
        return img, mask

    def forward(self, bg_inputs, src_inputs, tsf_inputs, Tst):
        """-> bg_img (bs,.,3,h,w), src_imgs (bs,ns,3,h,w), src_masks (bs,ns,1,h,w), tsf_imgs (bs,nt,3,h,w), tsf_masks (bs,nt,1,h,w)."""
        with _Branch(bg_inputs.device, 1) as br:          # BGNet does not depend on SIDNet / TSFNet: its own branch, forward and backward
            bg_img = self.forward_bg(bg_inputs)
        enc, res, src_imgs, src_masks = self.forward_src(src_inputs, only_enc=False)
        imgs, masks = [], []
        for t in range(tsf_inputs.shape[1]):
            i, m = self.forward_tsf(tsf_inputs[:, t], enc, res, Tst[:, t].contiguous())
            imgs.append(i); masks.append(m)
        br.join(bg_img)
        sb, self._src_branch = self._src_branch, None
        sb[0].join(sb[1], sb[2])
        return bg_img, src_imgs, src_masks, torch.stack(imgs, 1), torch.stack(masks, 1)


class PatchDiscriminator(nn.Module):
    """discriminators/patch_dis.py:8-70 as configured by AttLWB-SPADE.toml [Discriminator]: input 3 + 3 channels, ndf 64,
    4 stride-2 layers, instance norm, no sigmoid (LSGAN).  Same ``model.N`` parameter names as the reference."""

    def __init__(self, input_nc=6, ndf=64, n_layers=4, max_nf_mult=8):
        super().__init__()
        seq = [nn.Conv2d(input_nc, ndf, 4, 2, 1), nn.LeakyReLU(0.2, True)]
        mult = 1
        for n in range(1, n_layers):
            prev, mult = mult, min(2 ** n, max_nf_mult)
            seq += [nn.Conv2d(ndf * prev, ndf * mult, 4, 2, 1, bias=True), nn.InstanceNorm2d(ndf * mult, affine=False), nn.LeakyReLU(0.2, True)]
        prev, mult = mult, min(2 ** n_layers, max_nf_mult)
        seq += [nn.Conv2d(ndf * prev, ndf * mult, 4, 1, 1, bias=True), nn.InstanceNorm2d(ndf * mult, affine=False), nn.LeakyReLU(0.2, True)]
        seq += [nn.Conv2d(ndf * mult, 1, 4, 1, 1)]
        self.model = nn.Sequential(*seq)

    def forward(self, x):
        """The 4x4 convolutions (stride 2 / stride 1, pad 1; the 6-channel input and the 1-channel output zero-padded to 64) on the
        tcgen05 kernels, InstanceNorm2d + LeakyReLU(0.2) fused (``inorm``)."""
        if not (USE_KERNELS and x.is_cuda):
            return self.model(x)
        skip = False
        for m in self.model:
            if isinstance(m, nn.Conv2d):
                x = conv(x, m.weight, m.bias, stride=m.stride[0], padding=m.padding[0])
            elif isinstance(m, nn.InstanceNorm2d):
                x = inorm(x, act=True, slope=0.2, eps=m.eps)         # InstanceNorm2d + the LeakyReLU(0.2) that follows it
                skip = True
                continue
            elif not skip:
                x = m(x)
            skip = False
        return x.float()

    def native_weight(self):
        """{parameter: (kind, pad)} for ParamStore."""
        return {m.weight: (S2 if m.stride[0] == 2 else S1, m.padding[0]) for m in self.model if isinstance(m, nn.Conv2d)}


class VGG19Features(nn.Module):
    """criterions/vggloss.py VGG19 slices (relu1_1, relu2_1, relu3_1, relu4_1, relu5_1 of torchvision's layout), frozen."""
    CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512]
    CUTS = (2, 7, 12, 21, 30)

    def __init__(self):
        super().__init__()
        layers, cin = [], 3
        for v in self.CFG:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=False)]
                cin = v
        self.features = nn.Sequential(*layers)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, x):
        outs, lo = [], 0
        for hi in self.CUTS:
            for layer in self.features[lo:hi]:
                # frozen 3x3 convs: the same tcgen05 kernels (forward + data gradient; no weight gradient is requested)
                x = conv(x, layer.weight, layer.bias, relu=True).to(BF16) if isinstance(layer, nn.Conv2d) else \
                    (x if isinstance(layer, nn.ReLU) else layer(x))
            outs.append(x); lo = hi
        return outs


def vgg_target(vgg, y):
    """VGG features of the real image: no gradient, depends only on the batch -> computed on its own branch at the top of the step."""
    with torch.no_grad():
        return vgg(y.to(BF16))


def vgg_loss(vgg, x, y, weights=(1 / 32, 1 / 16, 1 / 8, 1 / 4, 1.0), fy=None):
    fx = vgg(x.to(BF16))
    if fy is None:
        fy = vgg_target(vgg, y)
    return sum(w * F.l1_loss(a.float(), b.float().detach()) for w, a, b in zip(weights, fx, fy))


def lsgan(outs, target):
    return torch.mean((outs.float() - target) ** 2)


def tv_loss(m):
    return torch.mean(torch.abs(m[:, :, :, :-1] - m[:, :, :, 1:])) + torch.mean(torch.abs(m[:, :, :-1, :] - m[:, :, 1:, :]))


class FlatGradBuckets:
    """Gradient all-reduce of iPERCore/services/train.py:89-95 (DDP) as a few flat NCCL all-reduces overlapped with backward:
    every parameter's .grad is a view into ONE flat fp32 buffer (reverse registration order = roughly the order the gradients
    become ready), buckets are contiguous slices of it; when the last gradient of a bucket has been accumulated its all-reduce
    starts asynchronously.  `offsets[param]` is the parameter's position — ParamStore lays parameters and Adam moments out the
    same way."""

    def __init__(self, params, n_buckets=4, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        params = [p for p in params if p.requires_grad][::-1]
        self.params = params
        total = sum(p.numel() for p in params)
        per = -(-total // max(1, n_buckets))
        self.buckets, cur, size = [], [], 0
        for p in params:
            cur.append(p); size += p.numel()
            if size >= per:
                self.buckets.append(cur); cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.storage = torch.zeros(total, dtype=torch.float32, device=params[0].device)
        self.flat, self.pending, self.handles, self.offsets, self.bucket_of = [], [], [], {}, {}
        off = 0
        for bi, b in enumerate(self.buckets):
            start = off
            for p in b:
                p.grad = self.storage[off:off + p.numel()].view_as(p)
                self.offsets[p] = off; self.bucket_of[p] = bi
                p.register_post_accumulate_grad_hook(self._hook(bi))
                off += p.numel()
            self.flat.append(self.storage[start:off]); self.pending.append(len(b))
        self.count = [0] * len(self.buckets)
        self.active = True              # False: gradients that are about to be discarded (D's, during the G step) are not reduced

    def _hook(self, bi):
        def fn(_p):
            if not self.active:
                return
            self.count[bi] += 1
            if self.count[bi] == self.pending[bi] and self.world > 1:
                if self.storage.is_cuda:                    # gradients of this bucket may have been produced on other branches
                    cur = torch.cuda.current_stream(self.storage.device)
                    for s_ in _all_streams(self.storage.device):
                        if s_ != cur:
                            cur.wait_stream(s_)
                self.handles.append(self.dist.all_reduce(self.flat[bi], op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))
        return fn

    def ready(self, p):
        """For gradients a kernel accumulated straight into the flat buffer (no autograd accumulation, so no hook)."""
        self._hook(self.bucket_of[p])(p)

    def zero(self):
        self.storage.zero_()
        self.count = [0] * len(self.buckets)
        self.handles = []

    def finish(self, average=True):
        """Wait for the in-flight all-reduces (buckets whose hooks never fired — unused parameters — are reduced now); average
        unless the caller folds 1/world into its optimizer pass."""
        if self.world > 1:
            for bi, f in enumerate(self.flat):
                if self.count[bi] != self.pending[bi]:
                    self.handles.append(self.dist.all_reduce(f, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))
            for h in self.handles:
                h.wait()
            if average:
                self.storage.div_(self.world)


class _Sink:
    def __init__(self, grad, buckets, param):
        self.grad, self._b, self._p = grad, buckets, param

    def ready(self):
        self._b.ready(self._p)


class ParamStore:
    """Flat fp32 master parameters + Adam moments laid out like the flat gradient of `buckets`, updated by ONE kernel per step
    (csrc/train.cu adam_pack_kernel = torch.optim.Adam(lr, betas, eps) of lwg_trainer.py's optimizers) that also rewrites the
    bf16 forward / dgrad packings of the convolution weights `native(name, param)` selects.  Parameters become views of the flat
    buffer (state_dict names and layouts unchanged: ``personalized.pth`` saves / loads as before; call ``repack()`` after loading)
    and get ``_iper_pack`` (the packings) and ``_iper_sink`` (their flat-gradient slice, written directly by the wgrad kernels)."""
    CHUNK = 4096

    def __init__(self, named_params, buckets, native=lambda name, p: None, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        import ctypes
        from ._lib import AdamSeg
        self.b, self.lr, self.betas, self.eps = buckets, lr, betas, eps
        dev = buckets.storage.device
        total = buckets.storage.numel()
        self.p = torch.empty(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)
        named = [(n, p) for n, p in named_params if p.requires_grad]
        segs, chunks, fwd_total, dg_total, packed = [], [], 0, 0, []
        for name, p in named:
            off = buckets.offsets[p]
            self.p[off:off + p.numel()].copy_(p.detach().reshape(-1))
            p.data = self.p[off:off + p.numel()].view_as(p)
            sg = AdamSeg(offset=off, numel=p.numel(), co=0, ci=0, taps=0, co_pad=0, ci_pad=0, reserved=0, fwd_offset=0, dgrad_offset=-1)
            kp = native(name, p)
            if kp:
                co, ci, k, _ = p.shape
                sg.co, sg.ci, sg.taps, sg.co_pad, sg.ci_pad = co, ci, k * k, _pad64(co), _pad64(ci)
                sg.reserved = kp[0] | (k << 8) | (kp[1] << 16)
                sg.fwd_offset, sg.dgrad_offset = fwd_total, dg_total
                fwd_total += sg.co_pad * sg.taps * sg.ci_pad; dg_total += sg.ci_pad * sg.taps * sg.co_pad
                packed.append((p, sg))
                # the gradient of a packed weight lives in the flat buffer as (co, tap, ci) — what the wgrad kernel writes fastest
                # and what the fused Adam pass reads; p.grad is the matching (co, ci, ky, kx) VIEW of it
                gs = buckets.storage[off:off + p.numel()].view(co, k * k, ci)
                p.grad = gs.view(co, k, k, ci).permute(0, 3, 1, 2)
                p._iper_sink = _Sink(gs, buckets, p)
            elif p.dim() == 1:
                p._iper_sink = _Sink(p.grad, buckets, p)         # biases of native layers are reduced by iper_bias_grad_bf16
            for c0 in range(0, p.numel(), self.CHUNK):
                chunks.append((len(segs), c0))
            segs.append(sg)
        self.pack_fwd = torch.zeros(max(fwd_total, 8), dtype=BF16, device=dev)
        self.pack_dgrad = torch.zeros(max(dg_total, 8), dtype=BF16, device=dev)
        for p, sg in packed:         # flat slices (the kernels take pointers; see pack_weight for the three layouts)
            size = sg.co_pad * sg.taps * sg.ci_pad
            p._iper_pack = (self.pack_fwd[sg.fwd_offset:sg.fwd_offset + size], self.pack_dgrad[sg.dgrad_offset:sg.dgrad_offset + size])
        arr = (AdamSeg * len(segs))(*segs)
        self.segs = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.chunks = torch.tensor(chunks, dtype=torch.int32).to(dev)
        self.n_chunks = len(chunks)
        self.repack()

    def _launch(self, update):
        check(lib.iper_adam_pack(self.p.data_ptr(), self.b.storage.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.segs.data_ptr(),
                                 self.chunks.data_ptr(), self.n_chunks, self.CHUNK, self.lr, self.betas[0], self.betas[1], self.eps,
                                 1.0 / self.b.world, self.step_t.data_ptr(), int(update), self.pack_fwd.data_ptr(),
                                 self.pack_dgrad.data_ptr(), _stream()), "adam_pack")

    def repack(self):
        self._launch(False)

    def step(self):
        """One Adam step on the (summed, not yet averaged) flat gradient."""
        self.step_t += 1
        self._launch(True)


class LWGTrainStep:
    """One optimisation step of LWGTrainer.optimize_parameters (lwg_trainer.py:326-352): G step then D step, Adam lr 1e-4,
    betas (0.9, 0.999) (deploy.toml [Train]); losses of optimize_G / optimize_D with use_face = false.  ``graph=True`` captures
    the whole step (forward, both backward passes, the bucketed all-reduces, both optimizer passes) in one CUDA graph after a
    few eager warm-up steps: at batch 1 the step is ~1400 small launches and otherwise bound by launch latency."""

    def __init__(self, net, device, lr=1e-4, lambdas=None, distributed=False, graph=False, fused_adam=True):
        self.dev = device
        self.G = TrainableGenerator(net).to(device)
        self.D = PatchDiscriminator().to(device)
        self.vgg = VGG19Features().to(device).eval()
        self.lam = dict(rec=10.0, tsf=10.0, mask=5.0, smooth=1.0, adv=1.0)
        self.lam.update(lambdas or {})
        self.fused = bool(fused_adam and USE_KERNELS)
        self.use_graph, self._graph, self._static, self._out, self._eager_steps = bool(graph), None, None, None, 0
        if self.fused:
            self.bk_G = FlatGradBuckets(list(self.G.parameters()))
            self.bk_D = FlatGradBuckets(list(self.D.parameters()), n_buckets=2)
            self.st_G = ParamStore(list(self.G.named_parameters()), self.bk_G, native=TrainableGenerator.native_weight, lr=lr)
            dmap = self.D.native_weight()
            self.st_D = ParamStore(list(self.D.named_parameters()), self.bk_D, native=lambda n, p: dmap.get(p), lr=lr)
            self.opt_G = self.opt_D = None
        else:
            self.opt_G = torch.optim.Adam(self.G.parameters(), lr=lr, betas=(0.9, 0.999), capturable=self.use_graph)
            self.opt_D = torch.optim.Adam(self.D.parameters(), lr=lr, betas=(0.9, 0.999), capturable=self.use_graph)
            self.bk_G = FlatGradBuckets(list(self.G.parameters())) if distributed else None
            self.bk_D = FlatGradBuckets(list(self.D.parameters()), n_buckets=2) if distributed else None

    def _zero(self, opt, bk):
        if bk is not None:
            bk.zero()
        else:
            opt.zero_grad(set_to_none=True)

    def _update(self, opt, bk, store):
        if store is not None:
            bk.finish(average=False)          # 1 / world is folded into the Adam pass
            store.step()
        else:
            if bk is not None:
                bk.finish()
            opt.step()

    def set_lr(self, lr):
        """Learning rate of both optimizers (lwg_trainer.py:300-324 decays them together).  The rate is a launch argument of the
        fused Adam pass, so a captured step is dropped and re-captured when it changes."""
        if self.fused:
            changed = lr != self.st_G.lr
            self.st_G.lr = self.st_D.lr = lr
            if changed:
                self._graph = None
        else:
            for opt in (self.opt_G, self.opt_D):
                for grp in opt.param_groups:
                    grp["lr"] = lr

    def step(self, batch, trainable=True):
        """batch: bg_inputs (bs,1,4,h,w), src_inputs (bs,ns,6,h,w), tsf_inputs (bs,nt,6,h,w), Tst (bs,nt,ns,h,w,2), real_src
        (bs,ns,3,h,w), real_tsf (bs,nt,3,h,w), real_bg (bs,3,h,w), body_mask (bs,ns+nt,1,h,w) -> dict of loss values.
        trainable=False (optimize_parameters' flag, lwg_trainer.py:326-352): only D is updated; such steps run eagerly."""
        if not self.use_graph or not trainable:
            return self._step(batch, trainable)
        if self._graph is None:
            if self._eager_steps < 2:         # eager warm-up: lazy initialisations (cuDNN plans, NCCL, function attributes)
                self._eager_steps += 1
                return self._step(batch)
            self._static = {k: v.clone() for k, v in batch.items()}
            torch.cuda.synchronize(self.dev)
            from ._lib import launch_count
            n0 = launch_count()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._out = self._step(self._static)
            self.launches_per_step = launch_count() - n0       # libiper_b200 launches inside one replay
        for k, v in batch.items():
            if v.data_ptr() != self._static[k].data_ptr():
                self._static[k].copy_(v, non_blocking=True)
        self._graph.replay()
        return self._out

    def _step(self, batch, trainable=True):
        b = batch
        bs, nt = b["tsf_inputs"].shape[:2]
        ns = b["src_inputs"].shape[1]
        h, w = b["tsf_inputs"].shape[-2:]
        st_G, st_D = (self.st_G, self.st_D) if self.fused else (None, None)
        for p in self.G.parameters():
            p._iper_uses = 0                  # forward uses of a weight still waiting for their backward (see _Conv)
        _USED.clear()
        r_tsf = b["real_tsf"].reshape(bs * nt, 3, h, w)
        with _Branch(self.dev, 2) as br_vgg:
            fy = vgg_target(self.vgg, r_tsf) if trainable else None
        # ---- forward (lwg_trainer.py:699-731) ----
        with torch.set_grad_enabled(bool(trainable)):
            fake_bg, src_color, src_mask, tsf_color, tsf_mask = self.G(b["bg_inputs"], b["src_inputs"], b["tsf_inputs"], b["Tst"])
        fake_src = src_mask * fake_bg + (1 - src_mask) * src_color
        fake_tsf = tsf_mask * fake_bg + (1 - tsf_mask) * tsf_color
        fake_masks = torch.cat([src_mask, tsf_mask], dim=1)
        # ---- G step (optimize_G :733-795) ----
        tsf_cond = b["tsf_inputs"][:, :, -3:].reshape(bs * nt, 3, h, w)
        f_tsf = fake_tsf.reshape(bs * nt, 3, h, w)
        for p in self.D.parameters():         # the adversarial term back-propagates THROUGH D: no gradients for D's own parameters
            p.requires_grad_(False)
        d_fake = self.D(torch.cat([f_tsf, tsf_cond], dim=1))
        for p in self.D.parameters():
            p.requires_grad_(True)
        l_adv = lsgan(d_fake, 0.0) * self.lam["adv"]
        l_rec = (F.l1_loss(fake_src, b["real_src"]) + F.l1_loss(fake_bg.reshape(-1, 3, h, w), b["real_bg"])) / 2 * self.lam["rec"]
        br_vgg.join(*(fy or ()))
        l_tsf = vgg_loss(self.vgg, f_tsf, r_tsf, fy=fy) * self.lam["tsf"]
        fm = fake_masks.reshape(bs * (ns + nt), 1, h, w)
        l_mask = F.l1_loss(fm, b["body_mask"].reshape(bs * (ns + nt), 1, h, w)) * self.lam["mask"]
        l_smooth = tv_loss(fm) * self.lam["smooth"]
        loss_G = l_rec + l_tsf + l_adv + l_mask + l_smooth
        if trainable:
            self._zero(self.opt_G, self.bk_G)
            if self.bk_D is not None:
                self.bk_D.active = False      # the adversarial term back-propagates through D: those gradients are discarded
            loss_G.backward()
            if self.bk_D is not None:
                self.bk_D.active = True
            self._update(self.opt_G, self.bk_G, st_G)
        # ---- D step (optimize_D :797-834) ----
        real_in = torch.cat([r_tsf, tsf_cond], dim=1)
        fake_in = torch.cat([f_tsf.detach(), tsf_cond], dim=1)
        d_out = self.D(torch.cat([real_in, fake_in], dim=0))      # one pass over both (instance norm is per sample: same result)
        nb = real_in.shape[0]
        loss_D = lsgan(d_out[:nb], 1.0) + lsgan(d_out[nb:], -1.0)
        self._zero(self.opt_D, self.bk_D)
        loss_D.backward()
        self._update(self.opt_D, self.bk_D, st_D)
        return dict(G=loss_G.detach(), D=loss_D.detach(), rec=l_rec.detach(), tsf=l_tsf.detach(), adv=l_adv.detach(),
                    mask=l_mask.detach(), smooth=l_smooth.detach())
