"""CPU oracle of the mask morphology used by source_setup — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Restates iPERCore/tools/utils/morphology/morph_ops.py: morph() :7-36 (erode: pad with 1, box sum == ks*ks; dilate: pad
with 0, box sum >= 1) and soft_dilate() :39-61 (pad with 0, box sum >= ks*ks/2) with an integral image in float64, which
is exact for the 0/1 masks the reference feeds it (flowcomposition.py:121,176-180,258).  Pinned against the reference's
own function by tests/golden/morph.npz (tests/golden/make_golden.py:make_morph)."""
import numpy as np

ERODE, DILATE, SOFT_DILATE = 0, 1, 2


def box_sum(mask, ks, pad_value):
    """mask (N,1,H,W) -> sum over the ks x ks window centred on every pixel, border padded with pad_value."""
    n, c, h, w = mask.shape
    p = ks // 2
    m = np.pad(mask.astype(np.float64), ((0, 0), (0, 0), (p, p), (p, p)), constant_values=pad_value)
    ii = np.zeros((n, c, h + 2 * p + 1, w + 2 * p + 1), dtype=np.float64)
    ii[:, :, 1:, 1:] = m.cumsum(2).cumsum(3)
    return ii[:, :, ks:ks + h, ks:ks + w] - ii[:, :, :h, ks:ks + w] - ii[:, :, ks:ks + h, :w] + ii[:, :, :h, :w]


def morph(mask, ks, mode):
    """mode ERODE / DILATE / SOFT_DILATE; returns float32 0/1 of the same shape (odd ks, as in the reference's configs)."""
    n_ks = ks * ks
    if mode == ERODE:
        return (box_sum(mask, ks, 1.0) == n_ks).astype(np.float32)
    s = box_sum(mask, ks, 0.0)
    return (s >= 1).astype(np.float32) if mode == DILATE else (s >= n_ks / 2).astype(np.float32)
