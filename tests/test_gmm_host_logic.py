"""CPU: the GMM host class (M-step algebra on the column statistics, variant semantics, outliers, stop rule) driven by
the CPU emulation of the EM kernels' arithmetic, against the reference's own outputs (golden) and the oracle."""
import numpy as np
import pytest
import torch

import emu_backend
from conftest import relerr
from oracle.gmm import GMMOracle

CPU = {"device": "cpu", "dtype": torch.float32}


def T32(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float32))


def build(g, tag, version):
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    D, N, C, outl, skip, steps, sig0 = g[f"{tag}_meta"]
    G = GaussianMixtureUnif(T32(g[f"{tag}_in_mu"]), sigma=float(sig0), use_outliers=bool(outl), spec=CPU, computversion=version)
    G.w = T32(g[f"{tag}_in_w"])
    G.to_optimize = dict(zip(("mu", "sigma", "w", "eta0"), (bool(v) for v in g[f"{tag}_opt"])))
    if outl:
        G.outliers["eta0"] = -1.0
    return G, T32(g[f"{tag}_in_X"]), bool(skip), int(steps)


def test_em_step_matches_reference_torch_twin(golden, monkeypatch):
    emu_backend.install(monkeypatch)
    g = golden("gmm")
    for tag in g["cases"]:
        tag = str(tag)
        G, X, skip, steps = build(g, tag, "torch")
        fes = []
        for _ in range(steps):
            Y, Cfe, FE = G.EM_step(X, skip_M=skip)
            fes.append(float(FE))
        # tolerance: 1e-5 relative to gold, or the reference's own fp32 distance to gold if that is larger
        def ok(a, key, base=2e-5):
            gold, ref = g[f"{tag}_gold_{key}"], g[f"{tag}_ref32_{key}"]
            e, e_ref = relerr(a, gold), relerr(ref, gold)
            assert e < max(base, 2 * e_ref), (tag, key, e, e_ref)
        ok(Y.numpy(), "Y")
        ok(G.mu.numpy(), "mu")
        ok(G.w.numpy(), "w", 5e-5)
        ok(np.array(G.sigma), "sigma")
        ok(np.array(float(Cfe)), "Cfe", 5e-5)
        ok(np.array(fes), "FE", 5e-5)
        if G.outliers is not None:
            assert abs(G.outliers["eta0"] - float(g[f"{tag}_gold_eta0"])) < 1e-4
            assert isinstance(Cfe, float) and isinstance(FE, float)
        else:
            assert isinstance(Cfe, torch.Tensor) and Cfe.dim() == 0


def test_em_step_keops_semantics_match_oracle(golden, monkeypatch):
    emu_backend.install(monkeypatch)
    g = golden("gmm")
    for tag in ["2d_full", "3d_full", "2d_outl", "3d_offset", "2d_opt5"]:
        G, X, skip, steps = build(g, tag, "keops")
        O = GMMOracle(torch.from_numpy(g[f"{tag}_in_mu"]).double(), G.sigma, w=torch.from_numpy(g[f"{tag}_in_w"]).double(),
                      outliers={"vol0": None, "eta0": -1.0} if G.outliers is not None else None, to_optimize=G.to_optimize)
        for _ in range(steps):
            Y, Cfe, FE = G.EM_step(X, skip_M=skip)
            Yo, Cfeo, FEo = O.em_step(X.double(), skip_M=skip, variant="keops")
        assert relerr(Y.numpy(), Yo.numpy()) < 3e-5, tag
        assert relerr(G.mu.numpy(), O.mu.numpy()) < 3e-5, tag
        assert abs(G.sigma - O.sigma) < 3e-5 * O.sigma, tag
        assert abs(float(FE) - float(FEo)) < 1e-4 * abs(float(FEo)), tag


def test_em_optimization_stop_rule(golden, monkeypatch):
    emu_backend.install(monkeypatch)
    g = golden("gmm")
    G, X, _, _ = build(g, "2d_full", "keops")
    Y, Cfe, FE, n = G.EM_optimization(X, max_iterations=50, tol=1e-4)
    assert 2 <= n < 50 and Y.shape == X.shape


def test_far_component_gets_finite_log_domain_update(monkeypatch):
    """A component farther than 13 sigma from every point (sum_n gamma_nc underflows in the linear domain) must still
    get finite mu / w, as with the reference's softmax-over-n update."""
    emu_backend.install(monkeypatch)
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    g = torch.Generator().manual_seed(0)
    X = torch.rand(300, 2, generator=g)
    mu = torch.tensor([[0.3, 0.3], [0.7, 0.6], [9.0, 9.0]])
    G = GaussianMixtureUnif(mu, sigma=0.1, spec=CPU)
    O = GMMOracle(mu.double(), 0.1)
    G.EM_step(X)
    O.em_step(X.double(), variant="keops")
    assert torch.isfinite(G.mu).all() and torch.isfinite(G.w).all()
    assert relerr(G.mu.numpy(), O.mu.numpy()) < 1e-4
    assert relerr(G.w.numpy(), O.w.numpy()) < 1e-4
