"""CPU: the row-subset VJP checker used by the full-size GPU tests equals the oracle's full autograd gradient, and the
kernels' adjoint formulas (CPU emulation of the same Op structs) agree with it."""
import pytest
import torch

import emu_backend
from adjoint_subset import oracle_subset_grad, oracle_subset_grad_lowmem


@pytest.mark.parametrize("version", ["classic", "hybrid", "logdet"])
def test_subset_gradient_equals_full_gradient(version, monkeypatch):
    from oracle.lddmm import LDDMMOracle
    emu_backend.install_all(monkeypatch)
    from diff_icp_b200 import ops
    D, M, sigma, g = 3, 150, 0.3, 0.7
    gen = torch.Generator().manual_seed(3)
    q = torch.rand(M, D, generator=gen)
    p, a, u = (torch.randn(M, D, generator=gen) for _ in range(3))
    OR = LDDMMOracle(sigma=sigma, D=D, lambd=50.0, version=version)
    qd, pd = q.double().requires_grad_(True), p.double().requires_grad_(True)
    vq, dp, dcost = OR.ode(qd, pd, torch.zeros(1, dtype=torch.float64))
    L = (a.double() * vq).sum() + (u.double() * dp).sum() + g * dcost.sum()
    fq, fp = torch.autograd.grad(L, [qd, pd])
    R = torch.arange(3, M, 9)
    sq, sp = oracle_subset_grad(OR, q.double(), p.double(), a.double(), u.double(), g, R)
    assert float((sq - fq[R]).abs().max()) < 1e-10 * float(fq.abs().max())
    assert float((sp - fp[R]).abs().max()) < 1e-10 * float(fp.abs().max())
    if version == "classic":
        lq, lp = oracle_subset_grad_lowmem(OR, q.double(), p.double(), a.double(), u.double(), R, block=40)
        assert float((lq - fq[R]).abs().max()) < 1e-10 * float(fq.abs().max())
        assert float((lp - fp[R]).abs().max()) < 1e-10 * float(fp.abs().max())
    gq, gp = torch.zeros_like(q), torch.zeros_like(q)
    ops.rhs_adjoint(D, OR.withlogdet, sigma, OR.eta, q, p, None, a, u, None, torch.tensor([g]), gq, gp, None, None)
    assert float((gq.double() - fq).abs().max()) < 2e-5 * float(fq.abs().max())
    assert float((gp.double() - fp).abs().max()) < 2e-5 * float(fp.abs().max())
