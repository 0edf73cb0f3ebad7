"""CPU oracle for the MEM pretraining hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline -- never on the CUDA product path
(``mem_b200`` raises if its CUDA library is missing; it has no CPU fallback).

Each module restates one piece of the reference algorithm (numpy / torch-CPU
fp32 / plain C) and cites the reference file:line it follows.

Pinning status: the reference (tum-vision/mem) ships no tests, golden vectors
or fixtures of its own, so the oracle is pinned against
outputs of the reference *itself*, executed in the build container through
``oracle/ref_shims.py`` by ``oracle/make_golden.py``; the resulting vectors are
committed under ``tests/golden/`` and re-checked by the ``-m "not gpu"`` tests.
"""
