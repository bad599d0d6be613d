"""Device time of one batched closure evaluation (CUDA events around graph replays), by frame count / sizes.
    python scripts/bench_batched_closure.py [--K 64] [--N 10000] [--M 25] [--D 2] [--version hybrid] [--scheme Euler]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=64)
    ap.add_argument("--N", type=int, default=10000)
    ap.add_argument("--M", type=int, default=25)
    ap.add_argument("--D", type=int, default=2)
    ap.add_argument("--nt", type=int, default=10)
    ap.add_argument("--version", default="hybrid")
    ap.add_argument("--scheme", default="Euler")
    ap.add_argument("--reps", type=int, default=30)
    args = ap.parse_args()
    from diff_icp_b200 import shooting
    from diff_icp_b200.core.LDDMM import LDDMMModel
    dev = torch.device("cuda:0")
    D, K, M, N = args.D, args.K, args.M, args.N
    LM = LDDMMModel(sigma=0.2, D=D, lambd=500.0, version=args.version, scheme=args.scheme, nt=args.nt,
                    spec={"device": dev, "dtype": torch.float32})
    g = torch.Generator().manual_seed(0)
    plan = shooting.BatchedClosurePlan(D, args.nt, args.scheme, LM.withlogdet, 0.2, LM.eta, 500.0, dev, [M] * K, [N] * K)
    plan.set_geometry([torch.rand(M, D, generator=g).to(dev) for _ in range(K)],
                      [torch.rand(N, D, generator=g).to(dev) for _ in range(K)])
    plan.set_targets(torch.rand(K * N, D, generator=g).to(dev), torch.full((K * N,), 50.0, device=dev))
    plan.active[:] = 1
    for k in range(K):
        plan.X[k, :M * D] = (0.01 * torch.randn(M * D, generator=g)).numpy()
    for _ in range(3):
        plan.evaluate()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
    ev[0].record()
    for r in range(args.reps):
        plan.graph.replay()
        ev[r + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[r].elapsed_time(ev[r + 1]) for r in range(args.reps))
    stages = args.nt * (1 if args.scheme == "Euler" else 2)
    pairs = 2.0 * stages * K * (N * M + M * M)
    print(json.dumps({"K": K, "N": N, "M": M, "D": D, "version": args.version, "scheme": args.scheme,
                      "xpass_env": os.environ.get("DICP_SMALL_XPASS"),
                      "closure_ms_median": ts[len(ts) // 2], "closure_ms_min": ts[0],
                      "pairs_per_s": pairs / (ts[len(ts) // 2] * 1e-3)}))


if __name__ == "__main__":
    main()
