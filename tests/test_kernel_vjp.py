"""CPU: standalone differentiable use of GaussKernel (first-order VJPs expressed as kernel sums) against autograd of the
oracle, with the kernels' arithmetic on the CPU emulation."""
import pytest
import torch

import emu_backend
from oracle.kernels import GaussOracle

CPU = {"device": "cpu", "dtype": torch.float32}


@pytest.mark.parametrize("D", [2, 3])
def test_vjps_match_oracle_autograd(monkeypatch, D):
    emu_backend.install_all(monkeypatch)
    from diff_icp_b200.tools.kernel import GaussKernel
    g = torch.Generator().manual_seed(D)
    M, N, sig = 40, 70, 0.6
    x, y = torch.rand(M, D, generator=g), torch.rand(N, D, generator=g)
    b, d = torch.randn(N, D, generator=g), torch.randn(N, generator=g)
    K, O = GaussKernel(sig, D, spec=CPU), GaussOracle(sig, D)
    for name, ins in [("KRed", [x, y, b]), ("KBase", [x, y]), ("KRedScal", [x, y, d]), ("GradKRed", [x, y])]:
        a32 = [t.clone().requires_grad_(True) for t in ins]
        out = getattr(K, name)(*a32)
        cot = torch.randn(out.shape, generator=g)
        out.backward(cot)
        a64 = [t.double().clone().requires_grad_(True) for t in ins]
        getattr(O, name)(*a64).backward(cot.double())
        for u, v in zip(a32, a64):
            assert float((u.grad.double() - v.grad).abs().max() / v.grad.abs().max()) < 2e-5, name
    with pytest.raises(NotImplementedError):
        K.HessKRed(x.clone().requires_grad_(True), y, b, torch.rand(M, D)).sum().backward()
