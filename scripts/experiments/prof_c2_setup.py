import cProfile, pstats, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda:0")
from diff_icp_b200.api.ICP_two_set import ICP_two_set
xA, y, _ = bench.make_workload(1234)
torch.zeros(1, device=dev); torch.cuda.synchronize()
def run():
    PSR, evol = ICP_two_set(xA.to(dev), y.to(dev), {"sigma": 0.1, "optimize_sigma": True, "outlier_weight": None},
                            {"type": "diffeomorphic", "lambda_LDDMM": 500.0, "sigma_LDDMM": 0.2},
                            numerical_options={"support_LDDMM": {"scheme": "dense"}},
                            optim_options={"max_iterations": 1}, plotstuff=False, printstuff=False)
    torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
print("first call", time.perf_counter() - t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
t0 = time.perf_counter(); run(); print("second call", time.perf_counter() - t0)
