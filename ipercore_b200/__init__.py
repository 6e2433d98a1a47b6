"""ipercore_b200 — B200-native (sm_100a) kernels for iPERCore's motion-imitation hot path.

Layout (SURVEY.md §8b seams):
  _lib.py            ctypes binding of libiper_b200.so (C ABI in include/iper_b200.h); fails loudly if missing
  ops.py             tensor-level wrappers (torch is only the allocator / stream provider)
  neural_renderer.py seam B1: drop-in for the `neural_renderer` functions iPERCore calls
  renders.py         seam B2: SMPLRenderer-compatible class on the fused raster/flow kernels
  generator.py       seam B3: AttentionLWBGenerator-compatible nn.Module (loads the reference's 221-key state_dict)
  engine.py          batched per-frame engine replacing Imitator.inference's bs=1 loop (+ frame sharding over ranks)
  patch.py           install() — swaps the seams into an importable iPERCore so run_imitator runs unchanged
"""
__all__ = ["build"]
