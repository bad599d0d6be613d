import math, os, sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scripts")
from groupwise_iteration import spiral_frames
from diff_icp_b200 import shooting
from diff_icp_b200.core.LDDMM import LDDMMModel
dev = torch.device("cuda:0"); spec = {"device": dev, "dtype": torch.float32}
LM = LDDMMModel(sigma=0.2, D=2, lambd=500.0, version="hybrid", scheme="Euler", nt=10, spec=spec)
LM.use_cuda_graph = True
q0 = torch.rand(25, 2, device=dev); p0 = 0.01 * torch.randn(25, 2, device=dev); x0 = torch.rand(10000, 2, device=dev)
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        sh = LM.Shoot(q0, p0, x0)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("shoot call ms", 1e3 * (t1 - t0), "sync ms", 1e3 * (t2 - t1))
sp = LM._spec_for(25, 10000, dev)
plan = shooting.ShootPlan.get(sp, True)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); plan.run_forward(q0, p0, x0); t1 = time.perf_counter()
    c = plan.traj.clone(); t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("run_forward ms", 1e3 * (t1 - t0), "clone ms", 1e3 * (t2 - t1), "sync", 1e3 * (t3 - t2))
