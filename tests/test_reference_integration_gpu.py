"""The zero-edit integration route on a GPU with the REAL reference classes (VERDICT r01 item 2 / SURVEY.md §8b seam B4):

    ipercore_b200.patch.install()  ->  iPERCore.models.imitator.Imitator(opt).source_setup(...) / .inference(...)

against the stock reference (same classes, torch/cuDNN true fp32, rasteriser = host C oracle) on the same inputs: same
pred_%08d.png files within one code value, same source-side uv image / background / face maps.  The reference tree is the
staged copy oracle/_ref (oracle/build_ref.py); SMPLH pkl and checkpoint are synthetic files in the reference's formats.
Each arm runs in its own process because install() patches classes process-wide.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(patched, work, size, frames, batch, temporal=0):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "ref_run_imitator.py"), "--patched", str(patched), "--work", work,
           "--size", str(size), "--frames", str(frames), "--batch", str(batch), "--temporal", str(temporal)]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert p.returncode == 0, p.stdout[-4000:]
    d = os.path.join(work, "patched" if patched else "stock")
    return d, json.load(open(os.path.join(d, "run.json")))


@pytest.mark.parametrize("temporal", [0, 1])
def test_real_imitator_patched_equals_stock(temporal, tmp_path):
    """temporal=0: the batched engine replaces Imitator.inference; temporal=1: the upstream bs=1 loop with its TemporalFIFO
    recurrence (imitator.py:18-127, 341-380) drives the B200 generator / renderer frame by frame."""
    from oracle import ref_runtime as rr
    if not rr.available():
        pytest.skip("reference tree not staged (oracle/build_ref.py needs /root/reference once)")
    import cv2
    S, T, B = (256, 5, 2) if not temporal else (128, 3, 2)      # 2 + 2 + 1 frames: CUDA-graph batches and a ragged tail
    work = str(tmp_path)
    d0, r0 = _run(0, work, S, T, B, temporal)
    d1, r1 = _run(1, work, S, T, B, temporal)
    assert r0["generator"].startswith("iPERCore.") and r0["nr_is_stub"]
    assert r1["generator"] == "ipercore_b200.generator" and r1["renderer_nr"] == "ipercore_b200.neural_renderer"
    assert r0["outputs"] == r1["outputs"] == ["pred_%08d.png" % i for i in range(T)]
    s0, s1 = np.load(os.path.join(d0, "source.npz")), np.load(os.path.join(d1, "source.npz"))
    np.testing.assert_array_equal(s0["fim"], s1["fim"])                       # source face-index maps bit-exact
    np.testing.assert_array_equal(s0["f2pts"], s1["f2pts"])
    np.testing.assert_allclose(s1["uv_img"], s0["uv_img"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(s1["bg"], s0["bg"], atol=1e-3, rtol=0)         # forward_bg on the tensor path
    worst, frac = 0, 0.0
    for name in r0["outputs"]:
        a = cv2.imread(os.path.join(d0, name), cv2.IMREAD_COLOR).astype(np.int32)
        b = cv2.imread(os.path.join(d1, name), cv2.IMREAD_COLOR).astype(np.int32)
        assert a.shape == b.shape == (S, S, 3)
        d = np.abs(a - b)
        worst, frac = max(worst, int(d.max())), max(frac, float((d > 0).mean()))
    print("real Imitator, patched vs stock: max code diff %d, worst fraction of differing pixels %.4f; stock %.2fs + %.2fs, "
          "patched %.2fs + %.2fs (source_setup + inference)" % (worst, frac, r0["source_setup_s"], r0["inference_s"],
                                                                 r1["source_setup_s"], r1["inference_s"]))
    assert worst <= 1 and frac < 0.02
