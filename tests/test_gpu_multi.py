"""GPU, world_size = 2 over NCCL (skipped with fewer than two devices; run with `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multi.py -m gpu`): the frame-sharded groupwise mode on real hardware gives the model a single GPU gives --
EM statistics through the one all-reduce per step, the pipelined EM loop against the single-GPU loop, and a whole atlas
iteration (GMM_opt + lock-step Reg_opt) against the single-process run."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _need_two():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")


def _frames(K=6, N=3000, seed=5, D=2):
    g = torch.Generator().manual_seed(seed)
    cent = torch.rand(7, D, generator=g)
    return [(cent[torch.randint(0, 7, (N + 37 * k,), generator=g)] + 0.03 * torch.randn(N + 37 * k, D, generator=g)
             + 0.02 * torch.randn(1, D, generator=g)).contiguous() for k in range(K)], cent


def _em_run(X, cent, comm, dev, steps=6):
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    spec = {"device": dev, "dtype": torch.float32}
    G = GaussianMixtureUnif((cent + 0.05).to(dev), sigma=0.1, spec=spec)
    G.to_optimize = {"mu": True, "sigma": True, "w": True, "eta0": True}
    G.comm = comm
    Y, Cfe, FE, n = G.EM_optimization(X.to(dev), max_iterations=steps, tol=1e-7)
    return {"mu": G.mu.cpu(), "w": G.w.cpu(), "sigma": float(G.sigma), "FE": float(FE), "Cfe": float(Cfe), "steps": n,
            "Y": Y.cpu()}


def _psr_run(frames, cent, comm, dev, mine):
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.core.PSR import DiffPSR
    import math
    spec = {"device": dev, "dtype": torch.float32}
    G = GaussianMixtureUnif((cent + 0.05).to(dev), sigma=0.1, spec=spec)
    LM = LDDMMModel(sigma=0.25, D=2, lambd=300.0, version="hybrid", scheme="Euler", nt=8, spec=spec)
    P = DiffPSR([frames[k].to(dev) for k in mine], G, LM, dataspec=spec, compspec=spec, comm=comm)
    P.printstuff = False
    P.set_support_scheme("grid", rho=math.sqrt(2))
    fes = []
    for _ in range(2):
        P.GMM_opt(max_iterations=5, tol=1e-4)
        P.Reg_opt(nmax=1, tol=1e-3)
        fes.append(P.FE)
    return {"mu": P.GMMi[0].mu.cpu(), "sigma": float(P.GMMi[0].sigma), "FE": fes, "q0": P.q0[0].cpu(),
            "x1": {k: P.x1[i, 0].cpu() for i, k in enumerate(mine)}}


def _worker(rank, world, port, out):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from diff_icp_b200.dist import StatsComm, shard_frames
    comm = StatsComm()
    frames, cent = _frames()
    mine = shard_frames(len(frames), rank, world)
    res = {"em": _em_run(torch.cat([frames[k] for k in mine]), cent, comm, dev),
           "psr": _psr_run(frames, cent, comm, dev, mine), "mine": mine}
    torch.save(res, os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def two_rank_results():
    _need_two()
    out = tempfile.mkdtemp()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    return [torch.load(os.path.join(out, f"r{r}.pt"), weights_only=False) for r in (0, 1)]


def test_two_gpu_em_equals_one_gpu(two_rank_results):
    r0, r1 = (r["em"] for r in two_rank_results)
    # both ranks hold the same model bit for bit (they read the same all-reduced numbers)
    assert torch.equal(r0["mu"], r1["mu"]) and torch.equal(r0["w"], r1["w"]) and r0["sigma"] == r1["sigma"]
    assert r0["FE"] == r1["FE"] and r0["steps"] == r1["steps"]
    frames, cent = _frames()
    one = _em_run(torch.cat(frames), cent, None, torch.device("cuda", 0))
    assert one["steps"] == r0["steps"]
    assert (one["mu"] - r0["mu"]).abs().max().item() <= 2e-5
    assert (one["w"] - r0["w"]).abs().max().item() <= 2e-5
    assert abs(one["sigma"] - r0["sigma"]) <= 2e-6 * one["sigma"] + 1e-8
    assert abs(one["FE"] - r0["FE"]) <= 2e-6 * abs(one["FE"])
    # targets of the rank's own points = the corresponding rows of the single-GPU targets
    sizes = [f.shape[0] for f in frames]
    offs = [sum(sizes[:k]) for k in range(len(frames))]
    for r in two_rank_results:
        rows = torch.cat([torch.arange(offs[k], offs[k] + sizes[k]) for k in r["mine"]])
        assert (one["Y"][rows] - r["em"]["Y"]).abs().max().item() <= 2e-5


def test_two_gpu_atlas_iteration_equals_one_gpu(two_rank_results):
    r0, r1 = (r["psr"] for r in two_rank_results)
    assert torch.equal(r0["mu"], r1["mu"]) and r0["sigma"] == r1["sigma"] and r0["FE"] == r1["FE"]
    assert torch.equal(r0["q0"], r1["q0"])                      # same grid support on every rank (global bounds)
    frames, cent = _frames()
    one = _psr_run(frames, cent, None, torch.device("cuda", 0), list(range(len(frames))))
    # the EM statistics are reduced in another order and the frames are batched differently (3 per rank instead of 6):
    # rounding-level differences that one unconverged L-BFGS step per frame amplifies
    assert abs(one["FE"][0] - r0["FE"][0]) <= 2e-4 * abs(one["FE"][0])
    assert abs(one["FE"][1] - r0["FE"][1]) <= 5e-4 * abs(one["FE"][1])
    assert (one["mu"] - r0["mu"]).abs().max().item() <= 2e-3 * 0.25
    assert abs(one["sigma"] - r0["sigma"]) <= 2e-3 * one["sigma"]
    for r in two_rank_results:
        for k, x1 in r["psr"]["x1"].items():
            assert (one["x1"][k] - x1).abs().max().item() <= 5e-3 * 0.25      # deformed points: well inside one kernel width
