"""GPU: the (q,q) adjoint at the bench size (20 000 points, 3-D, logdet: the symmetric ring engine, csrc/sym_engine.cuh)
and at the stress size (10^6 points, classic: blocked symmetric engine), against the ORACLE on a row subset.

Row-subset VJP.  The adjoint returns d/dq_m, d/dp_m of  L = sum_i a_i.vq_i + u_i.dp_i + g dcost  = sum over ordered pairs
(i,j) of phi(i,j).  For a subset R of points, the pairs that involve a point of R are {i in R, all j} and {i not in R,
j in R}: the oracle evaluates L restricted to those pairs with its own row reductions (v, GenDKRed, HessKRed, GradLapKRed,
mdivsum: oracle/lddmm.py, oracle/kernels.py -- restating core/LDDMM.py:176-227) in fp64 and differentiates it by autograd
with respect to (q_R, p_R); that equals the full gradient on R at O(|R| M) cost."""
import numpy as np
import pytest
import torch

from adjoint_subset import oracle_subset_grad, oracle_subset_grad_lowmem

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("version,M,sigma", [("logdet", 20000, 0.2), ("hybrid", 20000, 0.2), ("classic", 20000, 0.2),
                                             ("logdet", 20001, 0.1)])
def test_bench_size_adjoint_row_subset_vs_oracle(version, M, sigma):
    from diff_icp_b200 import ops
    from oracle.lddmm import LDDMMOracle
    D, lam = 3, 500.0
    gen = torch.Generator().manual_seed(17)
    q = torch.rand(M, D, generator=gen)
    p, a, u = (torch.randn(M, D, generator=gen) for _ in range(3))
    g = 0.7
    OR = LDDMMOracle(sigma=sigma, D=D, lambd=lam, version=version, chunk=4096)
    qd, pd, ad, ud = (t.to(dev()) for t in (q, p, a, u))
    gq, gp = torch.zeros_like(qd), torch.zeros_like(qd)
    ws = ops.alloc_workspace(M, M, dev())
    ops.rhs_adjoint(D, OR.withlogdet, sigma, OR.eta, qd, pd, None, ad, ud, None, torch.tensor([g], device=dev()), gq, gp, None, ws)
    R = torch.arange(5, M, M // 32)[:32]
    oq, op_ = oracle_subset_grad(OR, q.double(), p.double(), a.double(), u.double(), g, R)
    # scale = the largest entry of the full output (the subset's own maximum can sit on a quiet row)
    assert float((gq[R.to(dev())].cpu().double() - oq).abs().max()) < 2e-5 * float(gq.abs().max())
    assert float((gp[R.to(dev())].cpu().double() - op_).abs().max()) < 2e-5 * float(gp.abs().max())


def test_one_million_point_backward_row_subset_vs_oracle():
    """configs[4] size: ONE adjoint evaluation over 10^6 points (classic model; 10^12 ordered pairs through the blocked
    symmetric engine), 16 rows against the oracle."""
    from diff_icp_b200 import ops
    from oracle.lddmm import LDDMMOracle
    D, M, sigma = 3, 1_000_000, 0.05
    gen = torch.Generator().manual_seed(23)
    q = torch.rand(M, D, generator=gen)
    p = 1e-3 * torch.randn(M, D, generator=gen)
    a, u = torch.randn(M, D, generator=gen), 1e-3 * torch.randn(M, D, generator=gen)
    OR = LDDMMOracle(sigma=sigma, D=D, lambd=100.0, version="classic", chunk=4)
    qd, pd, ad, ud = (t.to(dev()) for t in (q, p, a, u))
    gq, gp = torch.zeros_like(qd), torch.zeros_like(qd)
    ws = ops.alloc_workspace(M, M, dev())
    ops.rhs_adjoint(D, False, sigma, 0.0, qd, pd, None, ad, ud, None, torch.zeros(1, device=dev()), gq, gp, None, ws)
    torch.cuda.synchronize()
    R = torch.arange(11, M, M // 16)[:16]
    OR.K.chunk = 32                                          # rows of the first part: one chunk of 32 x 10^6
    oq, op_ = oracle_subset_grad_lowmem(OR, q.double(), p.double(), a.double(), u.double(), R)
    assert float((gq[R.to(dev())].cpu().double() - oq).abs().max()) < 3e-5 * float(gq.abs().max())
    assert float((gp[R.to(dev())].cpu().double() - op_).abs().max()) < 3e-5 * float(gp.abs().max())
