"""Run each hot kernel once (after one warm-up) at the bench size, for ncu: forward/adjoint RHS of the three LDDMM
models, the EM row pass 20k x 20k and the plain KRed reduction."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from diff_icp_b200 import ops, shooting, em_ops
from diff_icp_b200.core.LDDMM import LDDMMModel

M = int(sys.argv[1]) if len(sys.argv) > 1 else bench.M_POINTS
D = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
xA, y, p0 = bench.make_workload(1234, M=M, D=D)
q, p, yy = xA.to(dev), p0.to(dev), y.to(dev)
ws = ops.alloc_workspace(M, M, dev)
for rep in range(2):
    for variant in ("classic", "hybrid", "logdet"):
        LM = LDDMMModel(sigma=bench.SIGMA_LDDMM, D=D, lambd=bench.LAMBDA_LDDMM, spec={"device": dev, "dtype": torch.float32},
                        version=variant, scheme="Euler", nt=bench.NT)
        sp = LM._spec_for(M, 0, dev)
        state = torch.cat([q.reshape(-1), p.reshape(-1), torch.zeros(1, device=dev)])
        lam = torch.randn(sp.S, device=dev)
        F = torch.zeros(sp.S + 3, device=dev)
        G = torch.zeros(sp.S, device=dev)
        shooting._rhs(sp, state, F, ws)
        shooting._vjp(sp, state, lam, G, ws)
    ops.ksum(ops.K_RED, 0.2, q, q, b=p, ws=ws)
    mu = yy.contiguous()
    wl2 = torch.zeros(M, device=dev)
    lpi = torch.zeros(M, device=dev)
    em_ops.rowpass(0.1, q, mu, wl2, mu, lpi)
torch.cuda.synchronize()
print("done")
