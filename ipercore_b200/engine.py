"""Batched per-frame synthesis engine — the B200 replacement of ``Imitator.inference``'s bs=1 loop
(iPERCore/models/imitator.py:327-382) for ``temporal=false`` (deploy.toml:40), where every target frame depends only
on the per-source cache and its own SMPL vertices/camera.

Per batch of B target frames (one CUDA-graph replay):
    raster_frames (fused: projection, rasterise, cond, UV sample, ns flows)       1 launch
    forward_tsf   (stem, 2 stride-2 convs, 9 AttLWB blocks, 12 res convs, SkipDecoder, heads + composite)
    pred_to_u8    ((x+1)/2*255, RGB->BGR like cv_utils.save_cv2_img)              1 launch
Host <-> device traffic per frame: 82.7 KB of vertices + 12 B camera in, 0.75 MB of uint8 pixels out, on a side
stream, double buffered so copies overlap the next batch's compute.  Frames shard contiguously over ranks with no
collective (``shard_range``).
"""
import torch

from . import ops


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) chunk of n_items for `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_batch(n_frames, max_batch):
    """Batch size <= max_batch that splits n_frames into equally sized batches (the last one at most nb-1 frames short):
    a 38-frame shard runs as one batch of 38 instead of a 60-frame graph with 22 idle slots."""
    if n_frames <= 0:
        return max(1, max_batch)
    nb = -(-n_frames // max(1, max_batch))
    return -(-n_frames // nb)


class FrameEngine:
    def __init__(self, generator, renderer, batch=16, use_graph=True, device=None):
        self.gen, self.render = generator, renderer
        self.B, self.S = int(batch), int(renderer.image_size)
        self.dev = device or next(generator.parameters()).device
        self.use_graph = use_graph
        self.graph = None
        self.src = None
        nv = None
        self.launches_per_batch = 0
        self._nv = nv
        self.compute = torch.cuda.Stream(device=self.dev)
        self.copy = torch.cuda.Stream(device=self.dev)

    # ---- one-time per source (imitator.py:177-246 source_setup; the geometry/morphology part stays upstream) -------
    @torch.no_grad()
    def set_source(self, src_inputs, uv_img, bg_img, src_f2pts):
        """src_inputs (1,ns,6,S,S) = input_G_src, uv_img (1,3,S,S), bg_img (1,3,S,S), src_f2pts (ns,nf,3,2)."""
        dev = self.dev
        enc, res = self.gen.forward_src(src_inputs.to(dev), only_enc=True)
        self.src = dict(enc=enc, res=res, uv_img=uv_img.to(dev).float().contiguous(),
                        bg=bg_img.to(dev).float().reshape(1, 3, self.S, self.S).contiguous(),
                        src_f2pts=src_f2pts.to(dev).float().contiguous())
        self.graph = None
        return self.src

    # ---- device-side step on static buffers ------------------------------------------------------------------------
    def _alloc_static(self, nv):
        B, S, dev = self.B, self.S, self.dev
        self.cams_d = torch.zeros((B, 3), dtype=torch.float32, device=dev)
        self.verts_d = torch.zeros((B, nv, 3), dtype=torch.float32, device=dev)
        self.u8_d = torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev)
        self._nv = nv

    def _step(self, n=None):
        """One batch on the static buffers; n < B (ragged last batch, eager only) processes just the first n frames."""
        s = self.src
        n = self.B if n is None else n
        fi = self.render.frame_inputs(self.cams_d[:n], self.verts_d[:n], s["uv_img"], s["src_f2pts"])
        img, mask, pred = self.gen.forward_tsf(fi["tsf_inputs"], s["enc"], s["res"], fi["Tst"], bg_img=s["bg"],
                                               return_pred=True)
        ops.pred_to_u8(pred, out=self.u8_d[:n])
        self.last_pred = pred

    def _ensure_ready(self, nv):
        if self.src is None:
            raise RuntimeError("FrameEngine.set_source must be called first")
        if self._nv != nv:
            with torch.cuda.stream(self.compute):
                self._alloc_static(nv)
            self.graph = None
        if self.graph is None:
            from . import _lib
            # the warm-up / capture below reads the source cache produced on the caller's stream (set_source)
            self.compute.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(self.compute):
                n0 = _lib.launch_count()
                self._step()                      # warm-up (also triggers the one-time weight repack)
                self.launches_per_batch = _lib.launch_count() - n0
                self.compute.synchronize()
                if self.use_graph:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.compute):
                        self._step()
                    self.graph = g
                else:
                    self.graph = False

    def run_batch_device(self, cams_d, verts_d):
        """Synthesize B frames whose inputs are already on the device; returns the static uint8 (B,S,S,3) BGR buffer.
        Work is issued on `self.compute`: the inputs must be safe to read there (long-lived tensors, or produced on that
        stream) and the caller orders its own stream against `self.compute`."""
        self._ensure_ready(verts_d.shape[1])
        with torch.cuda.stream(self.compute):
            n = cams_d.shape[0]
            self.cams_d[:n].copy_(cams_d, non_blocking=True)
            self.verts_d[:n].copy_(verts_d, non_blocking=True)
            self._run(n)
        return self.u8_d

    def _run(self, n):
        """Replay the B-frame graph for a full batch; a ragged batch (n < B) runs eagerly on exactly n frames — no stale
        frame of the previous batch is re-rendered."""
        if n == self.B and self.graph:
            self.graph.replay()
        else:
            self._step(n)

    # ---- streaming API: batches flow feed -> synthesize -> pinned ring -> sink, all three overlapped ------------------
    @torch.no_grad()
    def synthesize_stream(self, T, feed, sink, ring=3):
        """Synthesize T frames batch by batch with the OUTPUT STAGE overlapped.

        feed(lo, hi) -> (cams (n,3), verts (n,nv,3)) device tensors; it is called inside ``torch.cuda.stream(self.compute)``
            so whatever it launches (device LBS, an H2D copy from pinned memory) is ordered before the batch's kernels.
        sink(lo, hi, frames) is called on the calling thread as soon as the batch's uint8 BGR frames (n,S,S,3) are in
            pinned host memory — while the GPU already runs the next batch.  It may return futures (e.g. PNG encodes on a
            thread pool) that still read `frames`; the ring slot is not reused before they are done.
        Host memory is a bounded ring of `ring` pinned (B,S,S,3) buffers, not T frames."""
        B, S = self.B, self.S
        ring = max(2, int(ring))
        if not hasattr(self, "_ring") or len(self._ring) != ring or self._ring[0].shape[0] != B:
            self._ring = [torch.empty((B, S, S, 3), dtype=torch.uint8).pin_memory() for _ in range(ring)]
        done = [torch.cuda.Event() for _ in range(ring)]
        busy = [None] * ring                    # futures of the sink still reading slot r
        cur = torch.cuda.current_stream(self.dev)
        self.compute.wait_stream(cur)
        nb = (T + B - 1) // B

        def drain(i):                           # batch i's frames are on the host: hand them to the sink
            lo, hi = i * B, min((i + 1) * B, T)
            r = i % ring
            done[r].synchronize()
            busy[r] = sink(lo, hi, self._ring[r][:hi - lo])

        for i in range(nb):
            lo, hi = i * B, min((i + 1) * B, T)
            r = i % ring
            if busy[r]:
                for f in busy[r]:
                    f.result()
                busy[r] = None
            with torch.cuda.stream(self.compute):
                cams_d, verts_d = feed(lo, hi)
                self._ensure_ready(verts_d.shape[1])
                self.cams_d[:hi - lo].copy_(cams_d, non_blocking=True)
                self.verts_d[:hi - lo].copy_(verts_d, non_blocking=True)
                self._run(hi - lo)
                self._ring[r][:hi - lo].copy_(self.u8_d[:hi - lo], non_blocking=True)
                done[r].record(self.compute)
            if i >= 1:
                drain(i - 1)                    # overlaps batch i on the GPU
        if nb:
            drain(nb - 1)
        for fs in busy:
            for f in fs or ():
                f.result()
        cur.wait_stream(self.compute)

    # ---- public API: host in, host out (the call a run_imitator user makes) -----------------------------------------
    @torch.no_grad()
    def synthesize(self, cams, verts, out=None):
        """cams (T,3), verts (T,nv,3) host tensors (pinned for async copies) -> uint8 (T,S,S,3) BGR frames on the host."""
        T, nv = verts.shape[0], verts.shape[1]
        B, S = self.B, self.S
        self.compute.wait_stream(torch.cuda.current_stream(self.dev))
        self._ensure_ready(nv)
        if out is None:
            out = torch.empty((T, S, S, 3), dtype=torch.uint8).pin_memory()
        if not hasattr(self, "_stage") or self._stage[0][0].shape[1] != nv:
            self._stage = [(torch.empty((B, nv, 3), dtype=torch.float32, device=self.dev),
                            torch.empty((B, 3), dtype=torch.float32, device=self.dev),
                            torch.empty((B, S, S, 3), dtype=torch.uint8, device=self.dev)) for _ in range(2)]
            self._in_ready = [torch.cuda.Event() for _ in range(2)]
            self._out_ready = [torch.cuda.Event() for _ in range(2)]
            self._out_done = [torch.cuda.Event() for _ in range(2)]
            self._in_consumed = [torch.cuda.Event() for _ in range(2)]
        nb = (T + B - 1) // B
        cur = torch.cuda.current_stream(self.dev)
        self.copy.wait_stream(cur)
        self.compute.wait_stream(cur)

        def upload(i):
            lo, hi = i * B, min((i + 1) * B, T)
            sv, sc, _ = self._stage[i % 2]
            with torch.cuda.stream(self.copy):
                if i >= 2:
                    self.copy.wait_event(self._in_consumed[i % 2])     # batch i-2 has copied this slot out
                sv[:hi - lo].copy_(verts[lo:hi], non_blocking=True)
                sc[:hi - lo].copy_(cams[lo:hi], non_blocking=True)
                self._in_ready[i % 2].record(self.copy)

        upload(0)
        for i in range(nb):
            lo, hi = i * B, min((i + 1) * B, T)
            sv, sc, so = self._stage[i % 2]
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(self._in_ready[i % 2])
                self.cams_d[:hi - lo].copy_(sc[:hi - lo], non_blocking=True)
                self.verts_d[:hi - lo].copy_(sv[:hi - lo], non_blocking=True)
                self._in_consumed[i % 2].record(self.compute)
                self._run(hi - lo)
                if i >= 2:
                    self.compute.wait_event(self._out_done[i % 2])     # D2H of batch i-2 has drained this slot
                so[:hi - lo].copy_(self.u8_d[:hi - lo], non_blocking=True)
                self._out_ready[i % 2].record(self.compute)
            if i + 1 < nb:
                upload(i + 1)       # overlaps batch i's compute (slot (i+1)%2 was consumed by batch i-1's d2d copy)
            with torch.cuda.stream(self.copy):
                self.copy.wait_event(self._out_ready[i % 2])
                out[lo:hi].copy_(so[:hi - lo], non_blocking=True)
                self._out_done[i % 2].record(self.copy)
        cur.wait_stream(self.copy)
        cur.wait_stream(self.compute)
        return out
