"""GPU: v2p / KpinvSolve / ridge solve against the reference's own v2p outputs, both the dense branch
and the matrix-free truncated pseudo-inverse, plus the large-M path against a dense fp64 eigendecomposition."""
import numpy as np
import pytest
import torch

from conftest import relerr
from v2p_cases import check_v2p_against_reference

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def spec():
    return {"device": dev(), "dtype": torch.float32}


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(dev())


def test_v2p_dense_branch_matches_reference(golden):
    check_v2p_against_reference(golden("v2p"), cu, spec(), dense_max=4000)


def test_v2p_matrix_free_branch_matches_reference(golden):
    check_v2p_against_reference(golden("v2p"), cu, spec(), dense_max=0)


@pytest.mark.parametrize("D,M,sig", [(3, 6000, 0.2), (2, 5000, 0.15)])
def test_large_support_pinv_equals_dense_truncated_svd(D, M, sig):
    """Above DENSE_SOLVE_MAX the matrix-free path must give the reference's truncated pseudo-inverse: checked against a
    dense fp64 eigendecomposition of K(q,q) (what numpy lstsq computes, tools/kernel.py:227-232) with the cut-off moved
    into the nearest spectral gap; logdet model with zero target speeds = the a0 initialisation of the two-set API run."""
    from diff_icp_b200.core.LDDMM import LDDMMModel
    g = torch.Generator().manual_seed(5)
    q = torch.rand(M, D, generator=g).to(dev())
    LM = LDDMMModel(sigma=sig, D=D, lambd=500.0, version="logdet", spec=spec())
    qd = q.double()
    K = torch.exp(-torch.cdist(qd, qd) ** 2 / (2 * sig ** 2))
    lam, U = torch.linalg.eigh(K)
    lam, U = lam.flip(0), U.flip(1)
    rel = lam / lam[0]
    # choose rcond in the widest gap (in log scale) of the spectrum between 3e-4 and 3e-3
    cand = torch.nonzero((rel[:-1] < 3e-3) & (rel[1:] > 3e-4)).flatten()
    i = int(cand[torch.argmax(torch.log(rel[cand] / rel[cand + 1]))])
    rcond = float(torch.sqrt(rel[i] * rel[i + 1]))
    for v in (torch.zeros_like(q), LM.v(q, q, 0.1 * torch.randn(M, D, generator=g).to(dev()))):
        rhs = (v + LM.eta * LM.Kernel.GradKRed(q, q)).double()
        keep = rel > rcond
        gold = U[:, keep] @ ((U[:, keep].t() @ rhs) / lam[keep, None])
        p = LM.v2p(q, v, rcond=rcond)
        assert relerr(p.cpu().numpy(), gold.cpu().numpy()) < 2e-4
        # fitted speeds v(p) = K p - eta gradK: compared on the scale of the solve's right-hand side (for zero target
        # speeds v(p) itself is only the small residual of the truncated solve)
        vb = LM.v(q, q, p).double()
        want = K @ gold - LM.eta * LM.Kernel.GradKRed(q, q).double()
        assert float((vb - want).abs().max()) < 5e-5 * float(rhs.abs().max())


def test_v2p_round_trip_of_the_author(golden):
    """The reference's own self-check (core/LDDMM.py:809-813): v -> p = v2p(v, rcond=None) -> v(p) gives back v although
    p differs from the momenta that generated v."""
    from diff_icp_b200.core.LDDMM import LDDMMModel
    g = torch.Generator().manual_seed(1)
    M, D, sig, lam = 10, 2, 2.0, 100.0
    xt = torch.randn(M, D, generator=g).to(dev())
    bt = torch.randn(M, D, generator=g).to(dev())
    LM = LDDMMModel(sig, D, lambd=lam, version="classic", spec=spec())
    vt = LM.v(xt, xt, bt)
    pt = LM.v2p(xt, vt, rcond=None)
    assert relerr(LM.v(xt, xt, pt).cpu().numpy(), vt.cpu().numpy()) < 1e-4
    assert float(pt.norm()) <= float(bt.norm()) * (1 + 1e-4)          # minimum-norm solution
