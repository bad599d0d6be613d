"""Time the fused RHS forward/adjoint kernels for the three models (median of 10) -- used for tuning sweeps."""
import os, sys, torch, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from diff_icp_b200 import ops, shooting
from diff_icp_b200.core.LDDMM import LDDMMModel
M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
xA, y, p0 = bench.make_workload(1234, M=M, D=D)
q, p = xA.to(dev), p0.to(dev)
ws = ops.alloc_workspace(M, M, dev)
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
res = {}
for variant in ("classic", "hybrid", "logdet"):
    LM = LDDMMModel(sigma=0.2, D=D, lambd=500.0, spec={"device": dev, "dtype": torch.float32}, version=variant, scheme="Euler", nt=10)
    sp = LM._spec_for(M, 0, dev)
    state = torch.cat([q.reshape(-1), p.reshape(-1), torch.zeros(1, device=dev)])
    lam = torch.randn(sp.S, device=dev); F = torch.zeros(sp.S + 3, device=dev); G = torch.zeros(sp.S, device=dev)
    res[variant] = (round(timeit(lambda: shooting._rhs(sp, state, F, ws)), 4), round(timeit(lambda: shooting._vjp(sp, state, lam, G, ws)), 4))
print(os.environ.get("DICP_B200_LIB", "default").split("_")[-1], json.dumps(res))
