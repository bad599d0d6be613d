"""Host-side profile (cProfile) of one steady-state groupwise iteration: where do GMM_opt and Reg_opt spend their time?
    python scripts/profile_groupwise.py [--frames 64] [--points 10000] [--lockstep 1]"""
import argparse
import cProfile
import io
import math
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--points", type=int, default=10000)
    ap.add_argument("--lockstep", type=int, default=1)
    ap.add_argument("--top", type=int, default=35)
    args = ap.parse_args()
    from groupwise_iteration import spiral_frames
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.core.PSR import DiffPSR
    dev = torch.device("cuda:0")
    spec = {"device": dev, "dtype": torch.float32}
    frames = spiral_frames(args.frames, args.points)
    torch.manual_seed(1234)
    G = GaussianMixtureUnif(torch.zeros(50, 2), spec=spec)
    LM = LDDMMModel(sigma=0.2, D=2, lambd=500.0, version="hybrid", scheme="Euler", nt=10, spec=spec)
    LM.use_cuda_graph = True
    P = DiffPSR([f.to(dev) for f in frames], G, LM, dataspec=spec, compspec=spec)
    P.printstuff = False
    P.batched_lbfgs = bool(args.lockstep)
    P.set_support_scheme("grid", rho=math.sqrt(2))
    P.reinitialize_GMM()
    for _ in range(2):
        P.GMM_opt(max_iterations=10, tol=1e-3)
        P.Reg_opt(tol=1e-3, nmax=1)
    torch.cuda.synchronize()
    for name, fn in (("GMM_opt", lambda: P.GMM_opt(max_iterations=10, tol=1e-3)), ("Reg_opt", lambda: P.Reg_opt(tol=1e-3, nmax=1))):
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        fn()
        torch.cuda.synchronize()
        pr.disable()
        dt = time.perf_counter() - t0
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(args.top)
        print(f"==== {name}: {1e3 * dt:.2f} ms (under cProfile)")
        print(s.getvalue()[:9000])
    plans = getattr(P, "_bplan", None)
    plan = plans[0] if plans else None
    if plan is not None:
        print("closure evaluation rounds so far:", plan.evaluations)


if __name__ == "__main__":
    main()
