"""CPU oracle for the TLXCV CNN-backbone forward path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker or as the timed
CPU baseline.  ``tlxcv_b200`` never imports this package.

Contents
--------
``tlx_compat``   a torch-CPU-fp32 stand-in for the ~25 ``tensorlayerx`` symbols
                 the reference hot-path files touch.
                 ``tensorlayerx`` itself (requirements/requirements.txt:1,
                 ``tensorlayerx>=0.5.8``, unpinned, not vendored) is not
                 installable here; with TL_BACKEND=torch each of its layers
                 lowers to the ``torch.nn.functional`` call restated here.
``ref_loader``   executes the reference's OWN model files, unmodified, by path
                 from ``/root/reference`` against ``tlx_compat``.  Only works
                 where ``/root/reference`` is mounted (the build container).
``restated``     a self-contained functional restatement of the six hot-path
                 topologies (each function cites the reference file:line it
                 follows).  This is what travels to the GPU box.

Pinning
-------
The reference holds no tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8(c)).  The pins are minted here: ``tests/golden/make_golden.py``
runs the reference's own model files through ``ref_loader`` on seeded weights
and inputs and commits the logits / feature-map digests under
``tests/golden/``; ``restated`` is checked bit-for-bit against those fixtures
(tests/test_oracle.py) and, in the build container, directly against
``ref_loader``.  Per-op numerics come from PyTorch, not from tensorlayerx —
at the tensorlayerx boundary parity is therefore "pinned to the reference's
model files + torch.nn.functional", which is the closest attainable oracle.
"""
