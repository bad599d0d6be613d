"""Device time of ONE L-BFGS closure evaluation of one frame (shoot + quadratic loss + adjoint, shooting.ClosurePlan) at a
given size, e.g. a frame of configs[3]:  python scripts/closure_timing.py --M 1584 --N 50000 --D 3 --scheme Ralston
(--N 0: dense support, data points = support points).  Under `ncu --metrics gpu__time_duration.sum` with --graph 0 it
gives the per-kernel launch list of one closure."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=1584)
    ap.add_argument("--N", type=int, default=50000)
    ap.add_argument("--D", type=int, default=3)
    ap.add_argument("--nt", type=int, default=10)
    ap.add_argument("--version", default="hybrid")
    ap.add_argument("--scheme", default="Ralston")
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    from diff_icp_b200 import shooting
    from diff_icp_b200.core.LDDMM import LDDMMModel
    dev = torch.device("cuda:0")
    LM = LDDMMModel(sigma=0.2, D=a.D, lambd=500.0, version=a.version, scheme=a.scheme, nt=a.nt,
                    spec={"device": dev, "dtype": torch.float32})
    g = torch.Generator().manual_seed(0)
    q0 = torch.rand(a.M, a.D, generator=g).to(dev)
    x0 = torch.rand(a.N, a.D, generator=g).to(dev) if a.N else None
    nd = a.N if a.N else a.M
    y = torch.rand(nd, a.D, generator=g).to(dev)
    inv = torch.full((nd,), 50.0, device=dev)
    p = (1e-3 * torch.randn(a.M, a.D, generator=g)).to(dev)
    cp = shooting.ClosurePlan(LM._spec_for(a.M, a.N, dev), bool(a.graph), 500.0)
    cp.set_problem(q0, x0, y, inv)
    for _ in range(2):
        cp.evaluate(p)
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cp.evaluate(p)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    stages = a.nt * (2 if a.scheme == "Ralston" else 1)
    pairs = 2.0 * stages * (float(nd if a.N else 0) * a.M + float(a.M) * a.M)
    print(json.dumps({"M": a.M, "N": a.N, "D": a.D, "version": a.version, "scheme": a.scheme, "graph": bool(a.graph),
                      "closure_ms_median": ts[len(ts) // 2], "pairs_per_closure": pairs,
                      "pairs_per_s": pairs / (ts[len(ts) // 2] * 1e-3)}))


if __name__ == "__main__":
    main()
