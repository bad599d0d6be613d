"""Groupwise PSR iteration time (the second half of BASELINE.json's metric): one GMM_opt(max 10 EM steps) + one
Reg_opt(nmax=1) over all frames of a diffICP_multi-like atlas (configs[2]: 64 frames x 10k pts, 2-D, C = 50 inferred,
hybrid model, Euler nt = 10, grid support rho = sqrt(2)), frames sharded over the ranks (strong scaling).

    python scripts/groupwise_iteration.py [--frames 64] [--points 10000] [--iters 3] [--graph 1]
    torchrun --nproc-per-node N scripts/groupwise_iteration.py ...
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def spiral_frames(K, N, seed=1234):
    g = torch.Generator().manual_seed(seed)
    C = 20
    t = torch.linspace(0, 2 * np.pi, C + 1)[:-1]
    mu0 = torch.stack((0.5 + 0.4 * (t / 7) * t.cos(), 0.5 + 0.3 * t.sin()), 1)
    frames = []
    for k in range(K):
        c = torch.randint(0, C, (N,), generator=g)
        x = mu0[c] + 0.025 * torch.randn(N, 2, generator=g)
        cen = torch.rand(4, 2, generator=g)
        amp = 0.04 * torch.randn(4, 2, generator=g)
        w = torch.exp(-((x[:, None, :] - cen[None]) ** 2).sum(-1) / (2 * 0.25 ** 2))
        frames.append((x + w @ amp).contiguous())
    return frames


def run_groupwise(rank, world, dev, comm, n_frames=64, n_points=10000, C=50, iters=3, graph=True, workers=1,
                  lockstep=True, weak=False, groups=None, scheme="Euler"):
    """Returns the dict reported as `groupwise_psr_iteration` (rank 0) -- times are the max over ranks.
    weak=True: `n_frames` frames PER RANK (atlas of n_frames * world frames) instead of n_frames in total."""
    if weak:
        n_frames = n_frames * world
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.core.PSR import DiffPSR
    from diff_icp_b200.dist import shard_frames
    spec = {"device": dev, "dtype": torch.float32}
    frames = spiral_frames(n_frames, n_points)
    mine = shard_frames(n_frames, rank, world)
    torch.manual_seed(1234)
    G = GaussianMixtureUnif(torch.zeros(C, 2), spec=spec)
    LM = LDDMMModel(sigma=0.2, D=2, lambd=500.0, version="hybrid", scheme=scheme, nt=10, spec=spec)
    LM.use_cuda_graph = bool(graph)
    P = DiffPSR([frames[k].to(dev) for k in mine], G, LM, dataspec=spec, compspec=spec, comm=comm)
    P.printstuff = False
    P.frame_workers = workers
    P.batched_lbfgs = bool(lockstep)
    if groups is not None:
        P.lockstep_groups = int(groups)
    P.set_support_scheme("grid", rho=math.sqrt(2))
    P.reinitialize_GMM()
    times = []
    for it in range(iters):
        torch.cuda.synchronize()
        if comm is not None:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        P.GMM_opt(max_iterations=10, tol=1e-3)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        P.Reg_opt(tol=1e-3, nmax=1)
        torch.cuda.synchronize()
        if comm is not None:
            torch.distributed.barrier()
        t2 = time.perf_counter()
        times.append((t1 - t0, t2 - t1))
    if comm is not None:
        tt = torch.tensor(times, device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        times = tt.tolist()
    return {"metric": "groupwise_psr_iteration_ms", "n_gpus": world, "frames": n_frames, "points_per_frame": n_points,
            "C": C, "support_points": int(P.q0[0].shape[0]), "model": f"hybrid, {scheme} nt=10, grid support rho=sqrt(2), 2-D",
            "scaling": "weak (frames per rank fixed)" if weak else "strong (frames sharded over ranks)", "cuda_graph": bool(graph), "frame_workers": workers, "lockstep_lbfgs": bool(lockstep), "lockstep_groups": len(getattr(P, "_bplan", None) or []),
            "FE": P.FE, "sigma": P.GMMi[0].sigma,
            "gmm_opt_ms": [1e3 * a for a, _ in times], "reg_opt_ms": [1e3 * b for _, b in times],
            "iteration_ms_steady": 1e3 * sum(times[-1])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--points", type=int, default=10000)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--C", type=int, default=50)
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--lockstep", type=int, default=1)
    ap.add_argument("--weak", type=int, default=0, help="1: --frames is the number of frames per rank")
    ap.add_argument("--scheme", default="Euler", help="Euler (examples/diffICP_multi.py) or Ralston (the model's default)")
    ap.add_argument("--groups", type=int, default=None, help="frame groups of the lock-step registration (default: DiffPSR's)")
    args = ap.parse_args()
    rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
        from diff_icp_b200.dist import StatsComm
        comm = StatsComm()
    res = run_groupwise(rank, world, dev, comm, args.frames, args.points, args.C, args.iters, args.graph, args.workers, args.lockstep, bool(args.weak), args.groups, args.scheme)
    if rank == 0:
        print(json.dumps(res))
    if comm is not None:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
