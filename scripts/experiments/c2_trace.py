import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda:0")
from diff_icp_b200.api.ICP_two_set import ICP_two_set
xA, y, _ = bench.make_workload(1234)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
PSR, evol = ICP_two_set(xA[:n].to(dev), y[:n].to(dev), {"sigma": 0.1, "optimize_sigma": True, "outlier_weight": None},
                        {"type": "diffeomorphic", "lambda_LDDMM": 500.0, "sigma_LDDMM": 0.2},
                        numerical_options={"support_LDDMM": {"scheme": "dense"}},
                        optim_options={"max_iterations": 4}, plotstuff=False, printstuff=True)
print("regloss", PSR.regloss, "quadloss", PSR.quadloss, "Cfe", PSR.Cfe, "a0 absmax", float(PSR.a0[0].abs().max()))
