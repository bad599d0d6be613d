"""Eager fused closures in the small-support regime (25 support points, 10k data points, 2-D hybrid Euler) for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_icp_b200 import shooting
from diff_icp_b200.core.LDDMM import LDDMMModel
dev = torch.device("cuda:0")
M, Nx, D = int(sys.argv[1]) if len(sys.argv) > 1 else 25, int(sys.argv[2]) if len(sys.argv) > 2 else 10000, 2
g = torch.Generator().manual_seed(0)
q0 = torch.rand(M, D, generator=g).to(dev); p0 = (0.01 * torch.randn(M, D, generator=g)).to(dev)
x0 = torch.rand(Nx, D, generator=g).to(dev); y = torch.rand(Nx, D, generator=g).to(dev); inv = torch.full((Nx,), 50.0, device=dev)
LM = LDDMMModel(sigma=0.2, D=D, lambd=500.0, version="hybrid", scheme="Euler", nt=10, spec={"device": dev, "dtype": torch.float32})
sp = LM._spec_for(M, Nx, dev)
cp = shooting.ClosurePlan.get(sp, False, LM.lam)
cp.set_problem(q0, x0, y, inv)
for _ in range(3):
    L, gr = cp.evaluate(p0)
torch.cuda.synchronize()
print("loss", L)
