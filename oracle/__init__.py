"""CPU oracle for the diffICP hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch, dtype-agnostic: fp32 or fp64) restatement of the
reference algorithm (AdrienWohrer/diff-icp) for the hot path named in
BASELINE.json.  It exists so that the CUDA product path in ``diff_icp_b200`` can
be checked against an independent implementation.

Rules (enforced by tests/test_layout.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this package;
  * nothing under ``diff_icp_b200/`` imports it, and the product path has no CPU
    fallback: it raises when the CUDA library is missing.

Pinning: the reference ships no tests and no golden vectors.
The oracle is therefore pinned against outputs of the *reference itself*, run in
the build container from /root/reference by ``tests/golden/make_golden.py``
(script committed, fixtures committed as ``tests/golden/*.npz``).  The KeOps
variant of the EM step cannot be executed anywhere (pykeops absent, unpinned):
for that single function the oracle is a restatement of the published formulas
at ``core/GMM.py:402-529`` and is pinned only through the ``skip_M=True`` case,
where it coincides with the executable torch twin ("parity unpinned" for the
M-step ordering difference of the KeOps variant).
"""

from . import kernels, lddmm, gmm, pointsets  # noqa: F401
