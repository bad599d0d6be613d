"""Host profile of one steady-state Reg_opt of the two-set API at 20k x 20k (dense support, logdet): where the time outside the
closure's CUDA graph goes."""
import cProfile, pstats, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda:0")
from diff_icp_b200.api.ICP_two_set import ICP_two_set
xA, y, _ = bench.make_workload(1234)
PSR, evol = ICP_two_set(xA.to(dev), y.to(dev), {"sigma": 0.1, "optimize_sigma": True, "outlier_weight": None},
                        {"type": "diffeomorphic", "lambda_LDDMM": 500.0, "sigma_LDDMM": 0.2},
                        numerical_options={"support_LDDMM": {"scheme": "dense"}},
                        optim_options={"max_iterations": 1}, plotstuff=False, printstuff=False)
PSR.LMi.use_cuda_graph = True
for it in range(2):
    PSR.GMM_opt(max_iterations=10, tol=1e-3)
    PSR.Reg_opt(tol=1e-3, nmax=1)
PSR.GMM_opt(max_iterations=10, tol=1e-3)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter(); pr.enable()
PSR.Reg_opt(tol=1e-3, nmax=1)
torch.cuda.synchronize()
pr.disable(); print("Reg_opt s", time.perf_counter() - t0)
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
