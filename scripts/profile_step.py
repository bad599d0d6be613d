"""Minimal driver for ncu: N eager (no CUDA graph) closure evaluations of the bench workload, nothing else."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from diff_icp_b200.core.LDDMM import LDDMMModel

variant = sys.argv[1] if len(sys.argv) > 1 else "hybrid"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
M = int(sys.argv[3]) if len(sys.argv) > 3 else bench.M_POINTS
dev = torch.device("cuda:0")
xA, y, p0 = bench.make_workload(1234, M=M)
LM = LDDMMModel(sigma=bench.SIGMA_LDDMM, D=bench.DIM, lambd=bench.LAMBDA_LDDMM, spec={"device": dev, "dtype": torch.float32},
                version=variant, scheme="Euler", nt=bench.NT)
q, yy, pp = xA.to(dev), y.to(dev), p0.to(dev)
for _ in range(n):
    p = pp.clone().requires_grad_(True)
    sh = LM.Shoot(q, p)
    L = LM.trajloss(sh) + ((sh[-1][0] - yy) ** 2).sum() * 50.0
    L.backward()
torch.cuda.synchronize()
print("loss", float(L))
