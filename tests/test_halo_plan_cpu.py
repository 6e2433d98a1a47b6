"""CPU emulation of the halo kernel's tap program (host logic, no GPU): the plan exported by iper_conv_halo_plan is executed
with numpy on the packed weights exactly as conv_halo_pair_kernel walks it — A box per load, a view per entry, the weight
rows/K columns of the entry (per CTA rank for fused-N groups), accumulator blocks, phase of each block — and must
reproduce torch's Conv2d / ConvTranspose2d(4,2,1) / the 5x5 heads.  This pins schedule + weight packing + output mapping
together; the GPU parity tests then only have to show that the kernel executes the same program."""
import numpy as np
import torch
import torch.nn.functional as F


def _plan(mode, cin, rows, fuse_n=0):
    from ipercore_b200 import _lib
    buf = np.zeros(256, dtype=np.int32)
    n = _lib.lib.iper_conv_halo_plan(mode, cin, rows, fuse_n, buf.ctypes.data, buf.size)
    assert n > 0
    n_loads, acc_blocks, box_rows = (int(v) for v in buf[:3])
    blk_phase = [int(v) for v in buf[3:7]]
    loads = buf[7:7 + 4 * n_loads].reshape(n_loads, 4)
    ne = int(loads[-1, 2] + loads[-1, 3])
    ent = buf[7 + 4 * n_loads:7 + 4 * n_loads + 13 * ne].reshape(ne, 13)
    return acc_blocks, box_rows, loads, ent, blk_phase


def _run_plan(x, wm, mode, cin, bn, rows, tw, th, fuse_n=0):
    """x (H, W, Cin) float64 NHWC, wm (rows_total, K) packed weight matrix -> accumulators (H, W, acc_blocks, bn) walked
    tile by tile like the kernel (tiles of tw x th pixels, zero fill outside the image = TMA out-of-bounds fill)."""
    acc_blocks, box_rows, loads, ent, blk_phase = _plan(mode, cin, rows, fuse_n)
    H, W, _ = x.shape
    bh = box_rows // tw
    pad = 40      # tiles may overhang the image by up to a tile (+ halo): zero fill, like TMA
    xp = np.zeros((H + 2 * pad, W + 2 * pad, cin)); xp[pad:pad + H, pad:pad + W] = x
    out = np.zeros((H, W, acc_blocks, bn))
    for y0 in range(0, H, th):
        for x0 in range(0, W, tw):
            D = np.zeros((th, tw, acc_blocks * bn))
            for cc in range(cin // 64):
                for ox, oy, first, count in loads:
                    box = xp[pad + y0 + oy:pad + y0 + oy + bh, pad + x0 + ox:pad + x0 + ox + tw, cc * 64:(cc + 1) * 64]
                    for e in ent[first:first + count]:
                        a_off, b_row, b_k, blk, nblk = (int(v) for v in e[:5])
                        view = box[a_off // tw:a_off // tw + th]                    # (th, tw, 64): rows a_off .. a_off + 128
                        if nblk == 1:
                            B = wm[b_row:b_row + bn, b_k + cc * 64:b_k + (cc + 1) * 64]
                        else:                                                       # concatenation over the CTA pair's halves
                            fb_row, fb_k = e[5:9].reshape(2, 2), e[9:13].reshape(2, 2)
                            B = np.concatenate([wm[int(fb_row[r][i]):int(fb_row[r][i]) + bn,
                                                   int(fb_k[r][i]) + cc * 64:int(fb_k[r][i]) + (cc + 1) * 64]
                                                for r in range(2) for i in range(nblk // 2)], 0)
                        D[:, :, blk * bn:(blk + nblk) * bn] += view @ B.T
            yy, xx = min(th, H - y0), min(tw, W - x0)
            out[y0:y0 + yy, x0:x0 + xx] = D[:yy, :xx].reshape(yy, xx, acc_blocks, bn)
    return out, blk_phase


def _rand(shape, seed):
    return np.random.Generator(np.random.PCG64(seed)).standard_normal(shape)


def test_plan_reproduces_conv3x3():
    from ipercore_b200 import ops
    from ipercore_b200._lib import IPER_CONV_S1
    cin, cout, H, W = 128, 64, 11, 21
    x, w = _rand((H, W, cin), 1), _rand((cout, cin, 3, 3), 2)
    wm = torch.from_numpy(w).permute(0, 2, 3, 1).reshape(cout, -1).numpy()          # fp64 copy of pack_conv_weight's layout
    pw = ops.pack_conv_weight(torch.from_numpy(w).float(), 1)
    real = pw.w[0].float().numpy() / pw.scale                                     # the packer stores w * 2^k (exact)
    np.testing.assert_allclose(real, wm.astype(np.float32), rtol=2.0 ** -10, atol=0)   # same layout as the real packer (fp16 of w * 2^k)
    got, _ = _run_plan(x, wm, IPER_CONV_S1, cin, cout, cout, 16, 8)
    exp = F.conv2d(torch.from_numpy(x).permute(2, 0, 1)[None], torch.from_numpy(w), padding=1)[0].permute(1, 2, 0).numpy()
    np.testing.assert_allclose(got[:, :, 0], exp, atol=1e-9)


def _packT(w):
    """fp64 copy of ops.pack_convT_weight's layout (checked against the real packer below)."""
    kidx = {0: (1, 3), 1: (0, 2)}
    phases = []
    for py in range(2):
        for px in range(2):
            phases.append(np.concatenate([w[:, :, kidx[py][ta], kidx[px][tb]].T for ta in range(2) for tb in range(2)], 1))
    return np.concatenate(phases, 0)


def test_plan_reproduces_transposed_conv_both_forms():
    from ipercore_b200 import ops
    from ipercore_b200._lib import IPER_CONVT_4S2
    cin, cout, H, W = 128, 64, 9, 19
    x, w = _rand((H, W, cin), 3), _rand((cin, cout, 4, 4), 4)
    wm = _packT(w)
    pw = ops.pack_convT_weight(torch.from_numpy(w).float(), 1)
    real = pw.w[0].float().numpy() / pw.scale
    np.testing.assert_allclose(real, wm.astype(np.float32), rtol=2.0 ** -10, atol=0)   # same layout as the real packer (fp16 of w * 2^k)
    exp = F.conv_transpose2d(torch.from_numpy(x).permute(2, 0, 1)[None], torch.from_numpy(w), stride=2, padding=1)[0]
    exp = exp.permute(1, 2, 0).numpy()
    for fuse in (0, 1):
        acc, blk_phase = _run_plan(x, wm, IPER_CONVT_4S2, cin, cout, cout, 16, 8, fuse_n=fuse)
        got = np.zeros((2 * H, 2 * W, cout))
        for blk, phase in enumerate(blk_phase):
            got[(phase >> 1)::2, (phase & 1)::2] = acc[:, :, blk]                    # epilogue: oy = 2y + py, ox = 2x + px
        np.testing.assert_allclose(got, exp, atol=1e-9, err_msg="fuse_n=%d" % fuse)


def test_plan_reproduces_heads_row5():
    from ipercore_b200._lib import IPER_CONV_ROW5
    cin, H, W = 64, 10, 40
    x, wi, wk = _rand((H, W, cin), 5), _rand((3, cin, 5, 5), 6), _rand((1, cin, 5, 5), 7)
    w = np.concatenate([wi, wk], 0)                                               # (4, C, dy, dx)
    wm = np.concatenate([w.transpose(3, 0, 2, 1).reshape(20, 5 * cin), np.zeros((12, 5 * cin))], 0)   # pack_heads_weight
    # the kernel's tiles are 32 wide with a 4-pixel overlap; the emulation runs the vertical part on full-width "tiles"
    # (tw = W) and then the horizontal shift-add of the epilogue
    acc_blocks, box_rows, loads, ent, _ = _plan(IPER_CONV_ROW5, cin, 32)
    xp = np.zeros((H + 8, W, cin)); xp[4:4 + H] = x
    D = np.zeros((H, W, 32))
    for e in ent:
        dy = int(e[0]) // 32
        D += xp[4 + int(loads[0][1]) + dy:4 + int(loads[0][1]) + dy + H] @ wm[:, int(e[2]):int(e[2]) + cin].T
    Dp = np.zeros((H, W + 4, 32)); Dp[:, 2:2 + W] = D
    got = sum(Dp[:, dx:dx + W, dx * 4:dx * 4 + 4] for dx in range(5))            # out[x, o] = sum_dx D[x + dx - 2, (dx, o)]
    exp = F.conv2d(torch.from_numpy(x).permute(2, 0, 1)[None], torch.from_numpy(w), padding=2)[0].permute(1, 2, 0).numpy()
    np.testing.assert_allclose(got, exp, atol=1e-9)
