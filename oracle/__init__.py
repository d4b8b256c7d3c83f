"""CPU oracle for the UniMP / OpenFlamingo hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (`/root/reference`, weitianxin/UniMP) ships no tests,
golden vectors or fixtures and the model half of the path lives in
the un-vendored pip dependency ``open-flamingo==2.0.1`` (reference
``requirements.txt:35``), which is not importable in this image.  This package is a
plain-PyTorch *restatement* of that published algorithm plus the
in-tree label masking and focal loss (reference ``UniMP/mmrec.py:143-213``).  Its only
anchors are self-consistency checks (tests/test_oracle.py) and the golden vectors
generated from it under tests/golden/.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package, and only as the checker or the
CPU arm that is timed *beside* the product.  Nothing under ``unimp_b200/`` imports it.
"""
