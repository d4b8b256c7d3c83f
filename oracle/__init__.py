"""CPU oracle for the UniMP / OpenFlamingo hot path.  TEST INFRASTRUCTURE ONLY.

Pinning status:

* **PINNED against the reference run here** — everything that IS in the reference tree:
  label masking + focal loss + its gradient (`loss_oracle.py` vs the unmodified
  ``UniMP/mmrec.py::train_one_epoch``), and the vision tower (HF ``CLIPVisionModel`` vs the
  unmodified ``UniMP/xformers_model/clip.py``).  Fixtures ``tests/golden/ref_*.pt`` are
  produced by ``tests/golden/make_reference_golden.py`` (committed), which imports the
  reference from ``/root/reference`` in the build container; ``tests/test_reference_golden.py``
  checks the oracle (CPU) and the CUDA path (GPU) against them.
* **PARITY UNPINNED** — the Flamingo / Perceiver / gated masked cross-attention half
  (`flamingo_oracle.py`): the reference imports it from the un-vendored pip dependency
  ``open-flamingo==2.0.1`` (reference ``requirements.txt:35``), absent from
  ``/root/reference`` and not importable in this image, and the reference ships no tests or
  fixtures for it.  That file is a plain-PyTorch *restatement* of the
  published algorithm anchored on the reference's call sites and on
  self-consistency checks (tests/test_oracle.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package, and only as the checker or the
CPU arm that is timed *beside* the product.  Nothing under ``unimp_b200/`` imports it.
"""
