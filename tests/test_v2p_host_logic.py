"""CPU: v2p / KpinvSolve -- the dense branch and the matrix-free truncated pseudo-inverse -- on the CPU
emulation of the kernel sums, against the reference's own v2p outputs."""
import numpy as np
import torch

import emu_backend
from v2p_cases import check_v2p_against_reference

CPU = {"device": "cpu", "dtype": torch.float32}


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))


def test_v2p_dense_branch_matches_reference(golden, monkeypatch):
    emu_backend.install_all(monkeypatch)
    check_v2p_against_reference(golden("v2p"), _t, CPU, dense_max=4000)


def test_v2p_matrix_free_branch_matches_reference(golden, monkeypatch):
    """DENSE_SOLVE_MAX = 0 forces the large-M path (randomised subspace iteration on the kernel sum) at the fixture size."""
    emu_backend.install_all(monkeypatch)
    check_v2p_against_reference(golden("v2p"), _t, CPU, dense_max=0)


def test_top_eigenpairs_match_dense_spectrum(golden, monkeypatch):
    emu_backend.install_all(monkeypatch)
    from diff_icp_b200.tools.kernel import GaussKernel
    g = golden("v2p")
    tag = "3d_M280_classic"
    q = _t(g[f"{tag}_in_q"])
    sv = g[f"{tag}_gold_svals"]
    lam, U = GaussKernel(float(g[f"{tag}_meta"][2]), 3, spec=CPU).top_eigenpairs(q, 1e-3)
    n = int((sv > 1e-3 * sv[0]).sum())
    assert lam.shape[0] >= n
    assert np.abs(lam[:n].numpy() - sv[:n]).max() < 1e-5 * sv[0]
    assert float((U.t() @ U - torch.eye(U.shape[1], dtype=U.dtype)).abs().max()) < 1e-5
