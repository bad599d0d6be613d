"""Groupwise PSR iteration at the size of BASELINE.json configs[3] ("diffICP_full": 256 frames x 3 structures, 3-D,
50k points per frame, frames sharded k mod G over the ranks): one GMM_opt(<= 10 EM steps per structure) + one
Reg_opt(nmax=1) over all frames.  3-D lift of the reference's examples/diffICP_full.py:36-56 (three curve-shaped GMMs with
C_true = 20 and sigma 0.025 / 0.04 / 0.2), inferred with C = 20 per structure, w optimised, hybrid LDDMM model
(sigma = 0.2, lambda = 500, the example's default Ralston scheme, nt = 10), 3-D grid support with spacing rho*sigma,
rho = sqrt(2).

    python scripts/groupwise_c4.py [--frames 256] [--points 50000] [--iters 3]
    torchrun --nproc-per-node N scripts/groupwise_c4.py ...
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


def full_frames_3d(frames, n_points, seed=1234):
    """frames: iterable of frame indices (each frame's data depends only on (seed, index): ranks generate only their own)."""
    C = 20
    t = torch.linspace(0, 2 * np.pi, C + 1)[:-1]
    mus = [torch.stack((0.5 + 0.4 * (t / 7) * t.cos(), 0.5 + 0.3 * t.sin(), 0.5 + 0.15 * (2 * t).sin()), 1),
           torch.stack((1 + 0.4 * t.cos(), 0.5 + 0.4 * t.sin(), 0.5 + 0.2 * t.cos()), 1),
           torch.stack((0.8 + 0.1 * (t - np.pi), -0.06 * (t - np.pi), 0.5 + 0.05 * (t - np.pi)), 1)]
    sig = [0.025, 0.04, 0.2]
    out = []
    for k in frames:
        g = torch.Generator().manual_seed(seed * 100003 + int(k))
        cen = torch.rand(4, 3, generator=g) * torch.tensor([2.0, 2.0, 1.0]) + torch.tensor([0.0, -0.5, 0.0])
        amp = 0.04 * torch.randn(4, 3, generator=g)
        sets = []
        for s in range(3):
            n = n_points // 3 + (1 if s < n_points % 3 else 0)
            c = torch.randint(0, C, (n,), generator=g)
            x = mus[s][c] + sig[s] * torch.randn(n, 3, generator=g)
            w = torch.exp(-((x[:, None, :] - cen[None]) ** 2).sum(-1) / (2 * 0.4 ** 2))
            sets.append((x + w @ amp).contiguous())
        out.append(sets)
    return out


def run_c4(rank, world, dev, comm, n_frames=256, n_points=50000, C=20, iters=3, graph=True, lockstep=True, scheme="Ralston",
           workers=None, rho=None, groups=None, group_min=None):
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.core.PSR import DiffPSR
    from diff_icp_b200.dist import shard_frames
    spec = {"device": dev, "dtype": torch.float32}
    mine = shard_frames(n_frames, rank, world)
    frames = full_frames_3d(mine, n_points)
    torch.manual_seed(1234)
    G = GaussianMixtureUnif(torch.zeros(C, 3), spec=spec)
    G.to_optimize = {"mu": True, "sigma": True, "w": True, "eta0": True}
    LM = LDDMMModel(sigma=0.2, D=3, lambd=500.0, version="hybrid", scheme=scheme, nt=10, spec=spec)
    LM.use_cuda_graph = bool(graph)
    P = DiffPSR([[x.to(dev) for x in fr] for fr in frames], G, LM, dataspec=spec, compspec=spec, comm=comm)
    P.printstuff = False
    P.batched_lbfgs = bool(lockstep)
    if workers is not None:
        P.frame_workers = int(workers)
    if groups is not None:
        P.lockstep_groups = int(groups)
    if group_min is not None:
        P.lockstep_group_min_frames = int(group_min)
    P.set_support_scheme("grid", rho=math.sqrt(2) if rho is None else float(rho))
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
    plans = getattr(P, "_bplan", None)
    plan = plans[0] if plans else None
    return {"metric": "groupwise_psr_iteration_ms", "config": "diffICP_full-like (configs[3])", "n_gpus": world,
            "frames": n_frames, "structures": 3, "points_per_frame": n_points, "C_per_structure": C,
            "support_points": int(P.q0[0].shape[0]),
            "model": f"hybrid, {scheme} nt=10, sigma=0.2, lambda=500, 3-D grid support rho=sqrt(2)",
            "scaling": "strong (frames sharded k mod G over ranks)", "cuda_graph": bool(graph),
            "lockstep_lbfgs": bool(lockstep) and plan is not None, "lockstep_groups": len(plans or []),
            "closure_rounds_total": None if plan is None else plan.evaluations,
            "FE": P.FE, "sigma": [g.sigma for g in P.GMMi],
            "gmm_opt_ms": [1e3 * a for a, _ in times], "reg_opt_ms": [1e3 * b for _, b in times],
            "iteration_ms_steady": 1e3 * sum(times[-1])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--points", type=int, default=50000)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--lockstep", type=int, default=1)
    ap.add_argument("--rho", type=float, default=None, help="grid spacing in units of sigma (default sqrt(2))")
    ap.add_argument("--workers", type=int, default=None, help="frames registered concurrently on the per-frame path (threads + streams)")
    ap.add_argument("--scheme", default="Ralston")
    ap.add_argument("--groups", type=int, default=None, help="lock-step frame groups (DiffPSR.lockstep_groups)")
    ap.add_argument("--group-min", type=int, default=None, help="DiffPSR.lockstep_group_min_frames")
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
    res = run_c4(rank, world, dev, comm, args.frames, args.points, 20, args.iters, args.graph, args.lockstep, args.scheme,
                 args.workers, args.rho, args.groups, args.group_min)
    if rank == 0:
        print(json.dumps(res))
    if comm is not None:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
