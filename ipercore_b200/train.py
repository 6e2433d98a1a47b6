"""Training step of the Liquid Warping GAN on B200 (BASELINE.json configs[4]: 512x512, G + D + VGG
perceptual, bf16, NCCL gradient all-reduce) — FIRST STAGE, see DESIGN.md for what is and is not on hand-written kernels.

What runs on the kernels of csrc/train.cu: every 3x3 / stride-1 convolution with 64-aligned channels of the generator —
forward, data gradient, weight gradient in bf16 on tcgen05 (``conv3x3``, an ``autograd.Function``).  At 512x512 those are the
12 ResidualBlock convs, the 27 SPADE convs and the 2 skip convs of TSFNet plus the 12 of SIDNet: ~60 % of the generator's
training FLOPs.  Everything else of the step (strided / transposed / 1x1 / 5x5 / 7x7 convs, grid_sample, instance norm,
discriminator, VGG19, losses, Adam) is plain PyTorch in bf16 autocast for now: it is the scaffolding that makes the step
complete and measurable, not the product.

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


def _eligible(x, weight):
    """forward / dgrad need 64-aligned channels and a map of at least 8x16; the weight gradient (trainable layers only) also
    needs W % 64 == 0 (64-pixel K rows) and one channel count that is a multiple of 128 (its 128-row operand)."""
    co, ci, kh, kw = weight.shape
    n, _, h, w = x.shape
    if not (USE_KERNELS and x.is_cuda and kh == 3 and kw == 3 and ci % 64 == 0 and co % 64 == 0 and h >= 8 and w >= 16):
        return False
    return (not weight.requires_grad) or (w % 64 == 0 and (ci % 128 == 0 or co % 128 == 0))


def _conv_fwd(x_cl, w_packed, cout, bias, relu=False):
    n, cin, h, w = x_cl.shape
    y = torch.empty((n, cout, h, w), dtype=BF16, device=x_cl.device, memory_format=CL)
    check(lib.iper_conv3x3_bf16(x_cl.data_ptr(), n, h, w, cin, w_packed.data_ptr(), cout, 0 if bias is None else bias.data_ptr(),
                                int(relu), y.data_ptr(), _stream()), "conv3x3_bf16")
    return y


class _Conv3x3(torch.autograd.Function):
    """y = conv2d(x, weight, bias, stride 1, padding 1) in bf16 on the tcgen05 kernels; x any layout, y channels_last."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x_cl = x.to(BF16).contiguous(memory_format=CL)
        co, ci = weight.shape[:2]
        wp = weight.detach().permute(0, 2, 3, 1).reshape(co, 9 * ci).to(BF16).contiguous()      # K = (tap, ci)
        b = None if bias is None else bias.detach().float().contiguous()
        y = _conv_fwd(x_cl, wp, co, b)
        ctx.save_for_backward(x_cl, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x_cl, weight = ctx.saved_tensors
        co, ci = weight.shape[:2]
        n, _, h, w = x_cl.shape
        dy_cl = dy.to(BF16).contiguous(memory_format=CL)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dX = conv(dY, w') with w'(ci, (2-ky, 2-kx), co) = w(co, ci, ky, kx)
            wd = weight.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(ci, 9 * co).to(BF16).contiguous()
            dx = _conv_fwd(dy_cl, wd, ci, None)
        if ctx.needs_input_grad[1]:
            x_nchw = x_cl.contiguous()                      # pixels contiguous per channel: K-major rows for the pixel contraction
            dy_nchw = dy_cl.contiguous()
            g = torch.empty((co, 9, ci), dtype=torch.float32, device=x_cl.device)
            ws = torch.empty((lib.iper_conv3x3_wgrad_workspace_bytes(n, h, w, ci),), dtype=torch.uint8, device=x_cl.device)
            check(lib.iper_conv3x3_wgrad_bf16(x_nchw.data_ptr(), dy_nchw.data_ptr(), n, h, w, ci, co, g.data_ptr(), ws.data_ptr(),
                                              ws.numel(), _stream()), "conv3x3_wgrad_bf16")
            dw = g.view(co, 3, 3, ci).permute(0, 3, 1, 2).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy_cl.float().sum(dim=(0, 2, 3))
        return dx, dw, db


def _torch_conv(x, weight, bias, stride=1, padding=0, transposed=False):
    """The PyTorch side of the step.  Layers with fewer than 8 input or output channels (6-channel stems, the 3- / 1-channel
    heads, the discriminator's ends) run in fp32: cuDNN has no bf16 channels_last engine for some of their gradients
    ("GET was unable to find an engine"), and they are a rounding error of the step's FLOPs."""
    tiny = min(weight.shape[0], weight.shape[1]) < 8
    dt = torch.float32 if tiny else x.dtype
    fn = F.conv_transpose2d if transposed else F.conv2d
    y = fn(x.to(dt), weight.to(dt), None if bias is None else bias.to(dt), stride=stride, padding=padding)
    return y if tiny else y


def conv3x3(x, weight, bias=None):
    """3x3 / s1 / p1 convolution: tcgen05 bf16 kernels (forward, dgrad, wgrad) when the layer qualifies, torch otherwise."""
    if _eligible(x, weight):
        return _Conv3x3.apply(x, weight, bias)
    return _torch_conv(x, weight, bias, padding=1)


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

    def _c(self, name, x, stride=1, padding=0):
        w, b = self._p(name + ".weight"), self._p(name + ".bias")
        if w.shape[-1] == 3 and stride == 1 and padding == 1:
            return conv3x3(x, w, b)
        return _torch_conv(x, w, b, stride=stride, padding=padding)

    def _ct(self, name, x):
        w, b = self._p(name + ".weight"), self._p(name + ".bias")
        return _torch_conv(x, w, b, stride=2, padding=1, transposed=True)

    def _res(self, prefix, x, second):
        y = F.relu(self._c("%s.main.0" % prefix, x, padding=1))
        return x + self._c("%s.main.%d" % (prefix, second), y, padding=1)

    # ---- BGNet (bg_inpaintor.py:24-60) ----
    def forward_bg(self, bg_inputs):
        bs, ns, _, h, w = bg_inputs.shape
        x = bg_inputs.reshape(bs * ns, -1, h, w).to(BF16)
        inorm = lambda t: F.instance_norm(t.float(), eps=1e-5).to(BF16)
        x = F.relu(inorm(self._c("bg_net.main.0", x, padding=3)))       # 4 -> 64 (fp32 conv, see _torch_conv); inorm returns bf16
        idx = 3
        for _ in range(3):
            x = F.relu(inorm(self._c("bg_net.main.%d" % idx, x, stride=2, padding=1))); idx += 3
        for _ in range(self.n_res):
            y = F.relu(inorm(self._c("bg_net.main.%d.main.0" % idx, x, padding=1)))
            x = x + inorm(self._c("bg_net.main.%d.main.3" % idx, y, padding=1)); idx += 1
        for _ in range(3):
            x = F.relu(inorm(self._ct("bg_net.main.%d" % idx, x))); idx += 3
        return torch.tanh(self._c("bg_net.main.%d" % idx, x, padding=3).float()).view(bs, ns, 3, h, w)

    # ---- SIDNet (attlwb_spade_resunet.py:450-478, ResAutoEncoder) ----
    def forward_src(self, src_inputs, only_enc=False):
        bs, ns, _, h, w = src_inputs.shape
        x = src_inputs.reshape(bs * ns, -1, h, w).to(BF16)
        enc = []
        for i in range(3):
            x = F.relu(self._c("src_net.encoders.layers.%d.0" % i, x, stride=2, padding=1)).to(BF16); enc.append(x)
        res = []
        for i in range(self.n_res):
            x = self._res("src_net.res_blocks.%d" % i, x, 2); res.append(x)
        if only_enc:
            return enc, res
        d = x
        for i in range(3):
            d = F.relu(self._ct("src_net.decoders.layers.%d.0" % i, d))
        img = torch.tanh(self._c("src_net.img_reg.0", d, padding=2).float()).view(bs, ns, 3, h, w)
        mask = torch.sigmoid(self._c("src_net.att_reg.0", d, padding=2).float()).view(bs, ns, 1, h, w)
        return enc, res, img, mask

    # ---- SelfAttentionLWB (attlwb_spade_resunet.py:208-252) ----
    def _att(self, prefix, tsf_x, src_x, Tst):
        bs, ns, H, W, _ = Tst.shape
        h, w = tsf_x.shape[-2:]
        T = Tst.reshape(bs * ns, H, W, 2)
        if H != h or W != w:
            T = F.interpolate(T.permute(0, 3, 1, 2), size=(h, w), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
        warp = F.grid_sample(src_x.float(), T.float(), mode="bilinear", padding_mode="zeros", align_corners=False).to(BF16)
        k = self._c(prefix + ".fk", warp).view(bs, ns, -1, h, w)
        v = self._c(prefix + ".fv", warp).view(bs, ns, -1, h, w)
        q = self._c(prefix + ".fq", tsf_x)
        logits = (k.float() * q.float().unsqueeze(1)).sum(dim=2, keepdim=True) / math.sqrt(k.shape[2])
        a = (torch.softmax(logits, dim=1) * v.float()).sum(dim=1).to(BF16)
        normalized = F.instance_norm(tsf_x.float(), eps=1e-5)
        actv = F.relu(self._c(prefix + ".spade.mlp_shared.0", a, padding=1))
        gamma = self._c(prefix + ".spade.mlp_gamma", actv, padding=1).float()
        beta = self._c(prefix + ".spade.mlp_beta", actv, padding=1).float()
        return (normalized * (1 + gamma) + beta).to(BF16)

    def forward_tsf(self, tsf_inputs, src_enc, src_res, Tst):
        x = tsf_inputs.to(BF16)
        enc = []
        for i in range(3):
            x = F.relu(self._c("tsf_net_enc.layers.%d.0" % i, x, stride=2, padding=1)).to(BF16)
            x = self._att("enc_attlwbs.%d" % i, x, src_enc[i], Tst); enc.append(x)
        for i in range(self.n_res):
            x = self._res("res_blocks.%d" % i, x, 2)
            x = self._att("res_attlwbs.%d" % i, x, src_res[i], Tst)
        d = x
        for i in range(3):
            d = F.relu(self._ct("tsf_net_dec.upconvs.%d.0" % i, d))
            if i != 2:
                d = F.relu(self._c("tsf_net_dec.skippers.%d.0" % i, torch.cat([enc[1 - i], d], dim=1), padding=1))
        img = torch.tanh(self._c("tsf_img_reg.0", d, padding=2).float())
        mask = torch.sigmoid(self._c("tsf_att_reg.0", d, padding=2).float())
        return img, mask

    def forward(self, bg_inputs, src_inputs, tsf_inputs, Tst):
        """-> bg_img (bs,.,3,h,w), src_imgs (bs,ns,3,h,w), src_masks (bs,ns,1,h,w), tsf_imgs (bs,nt,3,h,w), tsf_masks (bs,nt,1,h,w)."""
        bg_img = self.forward_bg(bg_inputs)
        enc, res, src_imgs, src_masks = self.forward_src(src_inputs, only_enc=False)
        imgs, masks = [], []
        for t in range(tsf_inputs.shape[1]):
            i, m = self.forward_tsf(tsf_inputs[:, t], enc, res, Tst[:, t].contiguous())
            imgs.append(i); masks.append(m)
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
        return self.model(x)


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
                x = F.relu(conv3x3(x, layer.weight, layer.bias)).to(BF16) if isinstance(layer, nn.Conv2d) else \
                    (x if isinstance(layer, nn.ReLU) else layer(x))
            outs.append(x); lo = hi
        return outs


def vgg_loss(vgg, x, y, weights=(1 / 32, 1 / 16, 1 / 8, 1 / 4, 1.0)):
    fx = vgg(x.to(BF16))
    with torch.no_grad():
        fy = vgg(y.to(BF16))
    return sum(w * F.l1_loss(a.float(), b.float().detach()) for w, a, b in zip(weights, fx, fy))


def lsgan(outs, target):
    return torch.mean((outs.float() - target) ** 2)


def tv_loss(m):
    return torch.mean(torch.abs(m[:, :, :, :-1] - m[:, :, :, 1:])) + torch.mean(torch.abs(m[:, :, :-1, :] - m[:, :, 1:, :]))


class FlatGradBuckets:
    """Gradient all-reduce of iPERCore/services/train.py:89-95 (DDP) as a few flat NCCL all-reduces overlapped with backward:
    parameter .grad tensors are views into one flat buffer per bucket (reverse registration order = roughly the order the
    gradients become ready); when the last gradient of a bucket has been accumulated its all-reduce starts asynchronously."""

    def __init__(self, params, n_buckets=4, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        params = [p for p in params if p.requires_grad][::-1]
        total = sum(p.numel() for p in params)
        per = -(-total // max(1, n_buckets))
        self.buckets, cur, size = [], [], 0
        for p in params:
            cur.append(p); size += p.numel()
            if size >= per:
                self.buckets.append(cur); cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat, self.pending, self.handles = [], [], []
        for bi, b in enumerate(self.buckets):
            flat = torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device)
            off = 0
            for p in b:
                p.grad = flat[off:off + p.numel()].view_as(p); off += p.numel()
                p.register_post_accumulate_grad_hook(self._hook(bi))
            self.flat.append(flat); self.pending.append(len(b))
        self.count = [0] * len(self.buckets)

    def _hook(self, bi):
        def fn(_p):
            self.count[bi] += 1
            if self.count[bi] == self.pending[bi] and self.world > 1:
                self.handles.append(self.dist.all_reduce(self.flat[bi], op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))
        return fn

    def zero(self):
        for f in self.flat:
            f.zero_()
        self.count = [0] * len(self.buckets)
        self.handles = []

    def finish(self):
        """Wait for the in-flight all-reduces (buckets whose hooks never fired — unused parameters — are reduced now) and average."""
        if self.world > 1:
            for bi, f in enumerate(self.flat):
                if self.count[bi] != self.pending[bi]:
                    self.handles.append(self.dist.all_reduce(f, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))
            for h in self.handles:
                h.wait()
            for f in self.flat:
                f.div_(self.world)


class LWGTrainStep:
    """One optimisation step of LWGTrainer.optimize_parameters (lwg_trainer.py:326-352): G step then D step, Adam lr 1e-4,
    betas (0.9, 0.999) (deploy.toml [Train]); losses of optimize_G / optimize_D with use_face = false."""

    def __init__(self, net, device, lr=1e-4, lambdas=None, distributed=False):
        self.dev = device
        self.G = TrainableGenerator(net).to(device)
        self.D = PatchDiscriminator().to(device)
        self.vgg = VGG19Features().to(device).eval()
        self.lam = dict(rec=10.0, tsf=10.0, mask=5.0, smooth=1.0, adv=1.0)
        self.lam.update(lambdas or {})
        self.opt_G = torch.optim.Adam(self.G.parameters(), lr=lr, betas=(0.9, 0.999))
        self.opt_D = torch.optim.Adam(self.D.parameters(), lr=lr, betas=(0.9, 0.999))
        self.bk_G = FlatGradBuckets(list(self.G.parameters())) if distributed else None
        self.bk_D = FlatGradBuckets(list(self.D.parameters()), n_buckets=2) if distributed else None

    def _zero(self, opt, bk):
        if bk is not None:
            bk.zero()
        else:
            opt.zero_grad(set_to_none=True)

    def step(self, batch):
        """batch: bg_inputs (bs,1,4,h,w), src_inputs (bs,ns,6,h,w), tsf_inputs (bs,nt,6,h,w), Tst (bs,nt,ns,h,w,2), real_src
        (bs,ns,3,h,w), real_tsf (bs,nt,3,h,w), real_bg (bs,3,h,w), body_mask (bs,ns+nt,1,h,w) -> dict of loss values."""
        b = batch
        bs, nt = b["tsf_inputs"].shape[:2]
        ns = b["src_inputs"].shape[1]
        h, w = b["tsf_inputs"].shape[-2:]
        # ---- forward (lwg_trainer.py:699-731) ----
        fake_bg, src_color, src_mask, tsf_color, tsf_mask = self.G(b["bg_inputs"], b["src_inputs"], b["tsf_inputs"], b["Tst"])
        fake_src = src_mask * fake_bg + (1 - src_mask) * src_color
        fake_tsf = tsf_mask * fake_bg + (1 - tsf_mask) * tsf_color
        fake_masks = torch.cat([src_mask, tsf_mask], dim=1)
        # ---- G step (optimize_G :733-795) ----
        tsf_cond = b["tsf_inputs"][:, :, -3:].reshape(bs * nt, 3, h, w)
        f_tsf = fake_tsf.reshape(bs * nt, 3, h, w)
        r_tsf = b["real_tsf"].reshape(bs * nt, 3, h, w)
        d_fake = self.D(torch.cat([f_tsf, tsf_cond], dim=1))       # D runs in fp32/TF32: its ends have 6 and 1 channels (see _torch_conv)
        l_adv = lsgan(d_fake, 0.0) * self.lam["adv"]
        l_rec = (F.l1_loss(fake_src, b["real_src"]) + F.l1_loss(fake_bg.reshape(-1, 3, h, w), b["real_bg"])) / 2 * self.lam["rec"]
        l_tsf = vgg_loss(self.vgg, f_tsf, r_tsf) * self.lam["tsf"]
        fm = fake_masks.reshape(bs * (ns + nt), 1, h, w)
        l_mask = F.l1_loss(fm, b["body_mask"].reshape(bs * (ns + nt), 1, h, w)) * self.lam["mask"]
        l_smooth = tv_loss(fm) * self.lam["smooth"]
        loss_G = l_rec + l_tsf + l_adv + l_mask + l_smooth
        self._zero(self.opt_G, self.bk_G)
        loss_G.backward()
        if self.bk_G is not None:
            self.bk_G.finish()
        self.opt_G.step()
        # ---- D step (optimize_D :797-834) ----
        real_in = torch.cat([r_tsf, tsf_cond], dim=1)
        fake_in = torch.cat([f_tsf.detach(), tsf_cond], dim=1)
        loss_D = lsgan(self.D(real_in), 1.0) + lsgan(self.D(fake_in), -1.0)
        self._zero(self.opt_D, self.bk_D)
        loss_D.backward()
        if self.bk_D is not None:
            self.bk_D.finish()
        self.opt_D.step()
        return dict(G=loss_G.detach(), D=loss_D.detach(), rec=l_rec.detach(), tsf=l_tsf.detach(), adv=l_adv.detach(),
                    mask=l_mask.detach(), smooth=l_smooth.detach())
