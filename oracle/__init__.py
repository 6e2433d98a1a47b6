"""CPU oracle for the iPERCore motion-imitation hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``ipercore_b200/`` may import this package.  Allowed users: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and there only as
the checker or the timed CPU baseline, never as the product.

Pieces and how each is pinned (see DESIGN.md "Oracle"):

* ``raster_ref.c``   rasteriser restatement — **parity unpinned** (third-party CUDA dependency absent).
* ``flow_ref.py``    projection / flow / cond / syn-image restatement — pinned against the reference's own
                     ``SMPLRenderer`` + ``FlowComposition`` code (golden fixtures, tests/golden/make_golden.py).
* ``generator_ref.py`` AttLWB-SPADE generator restatement (plain torch fp32) — pinned against the reference's
                     own ``AttentionLWBGenerator`` run on CPU (golden fixtures).
"""
