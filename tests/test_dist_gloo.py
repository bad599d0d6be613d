"""CPU, world_size = 2 over gloo: the only collective of the algorithm (GMM statistics all-reduce) and the frame-sharded
groupwise mode give the same model as a single process holding all frames.  Kernels' arithmetic runs on the CPU
emulation (tests/hostemu); the NCCL path is exercised by bench.py --gpus N on the GPU box."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CPU = {"device": "cpu", "dtype": torch.float32}


class _Patch:
    def setattr(self, obj, name, val, raising=True):
        setattr(obj, name, val)


def _setup(rank, world, port):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emu_backend
    emu_backend.install_all(_Patch())
    torch.set_num_threads(1)


def _frames(K=4, N=120, seed=5):
    g = torch.Generator().manual_seed(seed)
    cent = torch.rand(5, 2, generator=g)
    return [(cent[torch.randint(0, 5, (N + 7 * k,), generator=g)] + 0.04 * torch.randn(N + 7 * k, 2, generator=g)).contiguous()
            for k in range(K)], cent


def _worker_em(rank, world, port, out):
    _setup(rank, world, port)
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.dist import StatsComm, shard_frames
    frames, cent = _frames()
    mine = shard_frames(len(frames), rank, world)
    X = torch.cat([frames[k] for k in mine])
    G = GaussianMixtureUnif(cent + 0.05, sigma=0.1, use_outliers=True, spec=CPU)
    G.outliers["vol0"] = 1.0
    G.comm = StatsComm()
    fes = []
    for _ in range(3):
        Y, Cfe, FE = G.EM_step(X)
        fes.append(float(FE))
    torch.save({"mu": G.mu, "w": G.w, "sigma": G.sigma, "FE": fes, "eta0": G.outliers["eta0"], "Y": Y, "mine": mine},
               os.path.join(out, f"em{rank}.pt"))
    dist.destroy_process_group()


def _worker_em_opt(rank, world, port, out):
    """EM_optimization over sharded points: the pipelined loop (ONE all-reduce per EM step, next step's first half issued
    speculatively) and the step-by-step loop, with a stop tolerance that triggers and with the step limit reached."""
    _setup(rank, world, port)
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.dist import StatsComm, shard_frames
    import torch.distributed as tdist
    frames, cent = _frames(K=5, N=150)
    mine = shard_frames(len(frames), rank, world)
    X = torch.cat([frames[k] for k in mine])
    res = {}
    for tag, pipelined, max_it, tol in (("pipe_tol", True, 40, 1e-4), ("seq_tol", False, 40, 1e-4),
                                        ("pipe_max", True, 4, 1e-12), ("seq_max", False, 4, 1e-12)):
        G = GaussianMixtureUnif(cent + 0.05, sigma=0.1, spec=CPU)
        G.comm = StatsComm()
        G.pipelined_allreduce = pipelined
        calls = {"n": 0}
        orig = tdist.all_reduce

        def counting(*a, **k):
            calls["n"] += 1
            return orig(*a, **k)
        tdist.all_reduce = counting
        try:
            Y, Cfe, FE, steps = G.EM_optimization(X, max_iterations=max_it, tol=tol)
        finally:
            tdist.all_reduce = orig
        res[tag] = {"mu": G.mu, "w": G.w, "sigma": G.sigma, "FE": float(FE), "Cfe": float(Cfe), "steps": steps, "Y": Y,
                    "allreduces": calls["n"]}
    torch.save(res, os.path.join(out, f"emopt{rank}.pt"))
    dist.destroy_process_group()


def _worker_vol0(rank, world, port, out):
    """Outlier reference volume left to be set automatically: must be the bounding box of ALL ranks' points."""
    _setup(rank, world, port)
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.dist import StatsComm, shard_frames
    frames, cent = _frames()
    mine = shard_frames(len(frames), rank, world)
    X = torch.cat([frames[k] for k in mine])
    G = GaussianMixtureUnif(cent + 0.05, sigma=0.1, use_outliers=True, spec=CPU)
    G.comm = StatsComm()
    Y, Cfe, FE = G.EM_step(X)
    torch.save({"vol0": G.outliers["vol0"], "eta0": G.outliers["eta0"], "FE": float(FE)}, os.path.join(out, f"vol{rank}.pt"))
    dist.destroy_process_group()


def _worker_init_from_set(rank, world, port, out):
    """init_components = {"set": 0, "C": 4}: every rank fits its initial GMM to ITS OWN first frame from random centroids;
    the constructor must make rank 0's model everyone's."""
    _setup(rank, world, port)
    from diff_icp_b200.api.ICP_atlas import ICP_atlas
    from diff_icp_b200.dist import StatsComm, shard_frames
    frames, cent = _frames(K=4, N=60)
    comm = StatsComm()
    mine = shard_frames(len(frames), rank, world)
    torch.manual_seed(100 + rank)                # different random draws per rank on purpose
    PSR, evol = ICP_atlas([frames[k] for k in mine], GMM_parameters={"init_components": {"set": 0, "C": 4}},
                          registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.3},
                          numerical_options={"compspec": CPU, "dataspec": CPU, "comm": comm,
                                             "support_LDDMM": {"scheme": "grid", "rho": 1.0}},
                          optim_options={"max_iterations": 1, "max_repeat_GMM": 2}, printstuff=False)
    torch.save({"mu0": evol["GMMi"][0].mu, "sigma0": evol["GMMi"][0].sigma, "mu": PSR.GMMi[0].mu, "FE": PSR.FE},
               os.path.join(out, f"init{rank}.pt"))
    dist.destroy_process_group()


def _worker_psr(rank, world, port, out):
    _setup(rank, world, port)
    from diff_icp_b200.api.ICP_atlas import ICP_atlas
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.dist import StatsComm, shard_frames
    frames, cent = _frames(K=3, N=60)
    comm = StatsComm()
    mine = shard_frames(len(frames), rank, world)
    G = GaussianMixtureUnif(cent + 0.05, sigma=0.1, spec=CPU)
    PSR, evol = ICP_atlas([frames[k] for k in mine], GMM_parameters={"init_components": [G]},
                          registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.3},
                          numerical_options={"compspec": CPU, "dataspec": CPU, "comm": comm,
                                             "support_LDDMM": {"scheme": "grid", "rho": 1.0}},
                          optim_options={"max_iterations": 2, "max_repeat_GMM": 3}, printstuff=False)
    torch.save({"mu": PSR.GMMi[0].mu, "sigma": PSR.GMMi[0].sigma, "FE": PSR.FE, "q0": PSR.q0[0]},
               os.path.join(out, f"psr{rank}.pt"))
    dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _spawn(fn, port=None):
    port = _free_port() if port is None else port
    out = tempfile.mkdtemp()
    mp.spawn(fn, args=(2, port, out), nprocs=2, join=True)
    return out


@pytest.fixture(scope="module", autouse=True)
def _build_emulation():
    sys.path.insert(0, HERE)
    import emu_backend
    emu_backend.lib()


def test_em_statistics_allreduce_equals_single_process(monkeypatch):
    out = _spawn(_worker_em)
    r0, r1 = (torch.load(os.path.join(out, f"em{r}.pt"), weights_only=False) for r in (0, 1))
    # identical replicated model on both ranks
    assert torch.equal(r0["mu"], r1["mu"]) and torch.equal(r0["w"], r1["w"]) and r0["sigma"] == r1["sigma"]
    assert r0["FE"] == r1["FE"] and r0["eta0"] == r1["eta0"]
    # and the same as one process holding all frames
    import emu_backend
    emu_backend.install_all(monkeypatch)
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    frames, cent = _frames()
    G = GaussianMixtureUnif(cent + 0.05, sigma=0.1, use_outliers=True, spec=CPU)
    G.outliers["vol0"] = 1.0
    fes = []
    for _ in range(3):
        Y, Cfe, FE = G.EM_step(torch.cat(frames))
        fes.append(float(FE))
    assert torch.allclose(G.mu, r0["mu"], rtol=0, atol=2e-6)
    assert torch.allclose(G.w, r0["w"], rtol=0, atol=2e-5)
    assert abs(G.sigma - r0["sigma"]) < 1e-6 * G.sigma
    assert np.allclose(fes, r0["FE"], rtol=2e-5)          # fp32 sums in a different order
    assert abs(G.outliers["eta0"] - r0["eta0"]) < 1e-5
    # targets of rank 0's frames are the corresponding rows of the global targets
    sizes = [f.shape[0] for f in frames]
    offs = np.concatenate(([0], np.cumsum(sizes)))
    Yg = torch.cat([Y[offs[k]:offs[k + 1]] for k in r0["mine"]])
    assert torch.allclose(Yg, r0["Y"], atol=2e-6)


def test_pipelined_em_optimization_one_allreduce_per_step(monkeypatch):
    out = _spawn(_worker_em_opt)
    r0, r1 = (torch.load(os.path.join(out, f"emopt{r}.pt"), weights_only=False) for r in (0, 1))
    for tag in r0:
        assert torch.equal(r0[tag]["mu"], r1[tag]["mu"]) and r0[tag]["FE"] == r1[tag]["FE"] and r0[tag]["steps"] == r1[tag]["steps"]
    for a, b in (("pipe_tol", "seq_tol"), ("pipe_max", "seq_max")):
        p, q = r0[a], r0[b]
        assert p["steps"] == q["steps"]
        assert torch.allclose(p["mu"], q["mu"], atol=2e-6) and torch.allclose(p["w"], q["w"], atol=2e-5)
        assert abs(p["sigma"] - q["sigma"]) < 1e-6 * q["sigma"]
        assert abs(p["FE"] - q["FE"]) < 2e-5 * abs(q["FE"]) and abs(p["Cfe"] - q["Cfe"]) < 2e-5 * abs(q["Cfe"])
        assert torch.allclose(p["Y"], q["Y"], atol=2e-6)
        # one collective per EM step (+ the MAX round of the very first step, + the closing reduction of the sums)
        assert p["allreduces"] <= p["steps"] + 3, (p["allreduces"], p["steps"])
        assert q["allreduces"] >= 2 * q["steps"]
    assert r0["pipe_tol"]["steps"] < 40 and r0["pipe_max"]["steps"] == 4
    # and the same model as one process holding all points
    import emu_backend
    emu_backend.install_all(monkeypatch)
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    frames, cent = _frames(K=5, N=150)
    G = GaussianMixtureUnif(cent + 0.05, sigma=0.1, spec=CPU)
    Y, Cfe, FE, steps = G.EM_optimization(torch.cat(frames), max_iterations=40, tol=1e-4)
    assert steps == r0["pipe_tol"]["steps"]
    assert torch.allclose(G.mu, r0["pipe_tol"]["mu"], atol=5e-6)
    assert abs(float(FE) - r0["pipe_tol"]["FE"]) < 5e-5 * abs(float(FE))


def test_outlier_volume_is_the_global_bounding_box(monkeypatch):
    out = _spawn(_worker_vol0)
    r0, r1 = (torch.load(os.path.join(out, f"vol{r}.pt"), weights_only=False) for r in (0, 1))
    assert r0 == r1
    import emu_backend
    emu_backend.install_all(monkeypatch)
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    frames, cent = _frames()
    G = GaussianMixtureUnif(cent + 0.05, sigma=0.1, use_outliers=True, spec=CPU)
    Y, Cfe, FE = G.EM_step(torch.cat(frames))
    assert abs(G.outliers["vol0"] - r0["vol0"]) < 1e-6 * G.outliers["vol0"]
    assert abs(G.outliers["eta0"] - r0["eta0"]) < 1e-5
    assert abs(float(FE) - r0["FE"]) < 2e-5 * abs(float(FE))


def test_initial_model_built_per_rank_is_broadcast():
    out = _spawn(_worker_init_from_set)
    r0, r1 = (torch.load(os.path.join(out, f"init{r}.pt"), weights_only=False) for r in (0, 1))
    assert torch.equal(r0["mu0"], r1["mu0"]) and r0["sigma0"] == r1["sigma0"]
    assert torch.equal(r0["mu"], r1["mu"]) and r0["FE"] == r1["FE"]


def test_sharded_atlas_equals_single_process(monkeypatch):
    out = _spawn(_worker_psr)
    r0, r1 = (torch.load(os.path.join(out, f"psr{r}.pt"), weights_only=False) for r in (0, 1))
    assert torch.equal(r0["mu"], r1["mu"]) and r0["sigma"] == r1["sigma"] and r0["FE"] == r1["FE"]
    assert torch.equal(r0["q0"], r1["q0"])                    # same grid support on every rank (global bounds)
    import emu_backend
    emu_backend.install_all(monkeypatch)
    from diff_icp_b200.api.ICP_atlas import ICP_atlas
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    frames, cent = _frames(K=3, N=60)
    G = GaussianMixtureUnif(cent + 0.05, sigma=0.1, spec=CPU)
    PSR, _ = ICP_atlas(frames, GMM_parameters={"init_components": [G]},
                       registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.3},
                       numerical_options={"compspec": CPU, "dataspec": CPU, "support_LDDMM": {"scheme": "grid", "rho": 1.0}},
                       optim_options={"max_iterations": 2, "max_repeat_GMM": 3}, printstuff=False)
    assert torch.allclose(PSR.q0[0], r0["q0"], atol=1e-6)
    assert torch.allclose(PSR.GMMi[0].mu, r0["mu"], atol=5e-4)       # frames are registered independently, in a
    assert abs(PSR.GMMi[0].sigma - r0["sigma"]) < 5e-3 * r0["sigma"]   # different order: L-BFGS paths agree to ~1e-4
    assert abs(PSR.FE - r0["FE"]) < 5e-3 * abs(r0["FE"])


def test_shard_frames_balances_work():
    from diff_icp_b200.dist import shard_frames
    K, W = 10, 4
    owned = [shard_frames(K, r, W) for r in range(W)]
    assert sorted(sum(owned, [])) == list(range(K))
    w = [100, 1, 1, 1, 50, 50, 1, 1, 1, 1]
    owned = [shard_frames(K, r, W, weights=w) for r in range(W)]
    assert sorted(sum(owned, [])) == list(range(K))
    loads = [sum(w[k] for k in o) for o in owned]
    assert max(loads) == 100
