"""ctypes face of the stand-alone C generator handle (include/iper_b200.h: iper_gen_*, csrc/generator.cu).

Everything the per-frame network needs — weight repacking, layer graph, workspace planning — happens inside
libiper_b200.so; this class only owns the device buffers the C API asks the caller to provide (torch is the allocator) and
forwards raw pointers.  It exists (a) as the proof that a non-Python host can run the hot path from the header alone and
(b) as an alternative backend of FrameEngine (``generator="c"``): same kernels, no Python in the layer loop.
"""
import ctypes

import torch

from ._lib import check, lib
from .ops import _stream


class GeneratorHandle:
    def __init__(self, state_dict, n_res=6, num_filters=(64, 128, 256), precision="fp16x2", device="cuda"):
        fmt = {"fp16": (1, 0), "fp16x2": (2, 0), "mixed": (2, 1)}[precision]
        self.dev = torch.device(device)
        self.h = ctypes.c_void_p()
        nf = (ctypes.c_int * 3)(*num_filters)
        check(lib.iper_gen_create(nf, n_res, fmt[0], fmt[1], ctypes.byref(self.h)), "gen_create")
        self._keep = []          # the fp32 reference tensors must stay alive until pack() has run
        for name, t in state_dict.items():
            t = t.detach().to(self.dev, torch.float32).contiguous()
            self._keep.append(t)
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            check(lib.iper_gen_load_weight(self.h, name.encode(), t.data_ptr(), shape, t.dim()), "gen_load_weight")
        n = lib.iper_gen_packed_bytes(self.h)
        if n == 0:
            raise RuntimeError("iper_gen_packed_bytes: %s" % lib.iper_last_error().decode())
        self.packed = torch.empty((n,), dtype=torch.uint8, device=self.dev)
        check(lib.iper_gen_pack(self.h, self.packed.data_ptr(), n, _stream()), "gen_pack")
        torch.cuda.current_stream(self.dev).synchronize()
        self._keep = []
        self.src_cache, self.ns, self.S = None, 0, 0

    def __del__(self):
        if getattr(self, "h", None):
            lib.iper_gen_destroy(self.h)
            self.h = None

    def forward_src(self, src_inputs):
        """src_inputs (1, ns, 6, S, S) or (ns, 6, S, S) fp32 -> fills the per-source cache (opaque device buffer)."""
        x = src_inputs.reshape(-1, 6, src_inputs.shape[-2], src_inputs.shape[-1]).to(self.dev, torch.float32).contiguous()
        ns, S = x.shape[0], x.shape[-1]
        self.src_cache = torch.empty((lib.iper_gen_src_cache_bytes(self.h, ns, S),), dtype=torch.uint8, device=self.dev)
        ws = torch.empty((lib.iper_gen_src_workspace_bytes(self.h, ns, S),), dtype=torch.uint8, device=self.dev)
        check(lib.iper_gen_forward_src(self.h, x.data_ptr(), ns, S, self.src_cache.data_ptr(), self.src_cache.numel(),
                                       ws.data_ptr(), ws.numel(), _stream()), "gen_forward_src")
        self.ns, self.S = ns, S
        return self.src_cache

    def workspace(self, B):
        return torch.empty((lib.iper_gen_tsf_workspace_bytes(self.h, self.ns, B, self.S),), dtype=torch.uint8, device=self.dev)

    def forward_tsf(self, tsf_inputs, Tst, bg_img=None, return_pred=False, workspace=None):
        """tsf_inputs (B,6,S,S), Tst (B,ns,S,S,2) -> (img (B,3,S,S), mask (B,1,S,S)[, pred (B,3,S,S)])."""
        if self.src_cache is None:
            raise RuntimeError("forward_src must run first")
        B, S = tsf_inputs.shape[0], self.S
        t = tsf_inputs.to(self.dev, torch.float32).contiguous(); T = Tst.to(self.dev, torch.float32).contiguous()
        if tuple(T.shape) != (B, self.ns, S, S, 2) or tuple(t.shape) != (B, 6, S, S):
            raise ValueError("forward_tsf: shapes %s / %s do not match the cached source set (ns=%d, S=%d)" %
                             (tuple(t.shape), tuple(T.shape), self.ns, S))
        img = torch.empty((B, 3, S, S), dtype=torch.float32, device=self.dev)
        mask = torch.empty((B, 1, S, S), dtype=torch.float32, device=self.dev)
        pred = bg = None
        if return_pred:
            bg = bg_img.to(self.dev, torch.float32).reshape(-1, 3, S, S).contiguous()
            pred = torch.empty((B, 3, S, S), dtype=torch.float32, device=self.dev)
        ws = self.workspace(B) if workspace is None else workspace
        check(lib.iper_gen_forward_tsf(self.h, t.data_ptr(), T.data_ptr(), self.src_cache.data_ptr(), self.ns, B, S,
                                       0 if bg is None else bg.data_ptr(), int(bg is not None and bg.shape[0] > 1), img.data_ptr(),
                                       mask.data_ptr(), 0 if pred is None else pred.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _stream()), "gen_forward_tsf")
        return (img, mask, pred) if return_pred else (img, mask)
