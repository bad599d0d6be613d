"""GPU parity tests of the GMM EM step, the outer PSR alternation and the API entry points against the reference's
own outputs (golden fixtures; fp64 "gold" and the reference's fp32 run "ref32") and the oracle."""
import numpy as np
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


SPEC = None


def spec():
    return {"device": dev(), "dtype": torch.float32}


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(dev())


def build(g, tag, version):
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    D, N, C, outl, skip, steps, sig0 = g[f"{tag}_meta"]
    G = GaussianMixtureUnif(cu(g[f"{tag}_in_mu"]), sigma=float(sig0), use_outliers=bool(outl), spec=spec(), computversion=version)
    G.w = cu(g[f"{tag}_in_w"])
    G.to_optimize = dict(zip(("mu", "sigma", "w", "eta0"), (bool(v) for v in g[f"{tag}_opt"])))
    if outl:
        G.outliers["eta0"] = -1.0
    return G, cu(g[f"{tag}_in_X"]), bool(skip), int(steps)


def test_em_step_matches_reference_torch_twin(golden):
    g = golden("gmm")
    for tag in g["cases"]:
        tag = str(tag)
        G, X, skip, steps = build(g, tag, "torch")
        lg = G.log_responsibilities(X)
        assert relerr(lg.cpu().numpy(), g[f"{tag}_gold_lgam"]) < 2e-5, tag
        fes = []
        for _ in range(steps):
            Y, Cfe, FE = G.EM_step(X, skip_M=skip)
            fes.append(float(FE))

        def ok(a, key, base=2e-5):
            gold, ref = g[f"{tag}_gold_{key}"], g[f"{tag}_ref32_{key}"]
            e, e_ref = relerr(a, gold), relerr(ref, gold)
            assert e < max(base, 2 * e_ref), (tag, key, e, e_ref)
        ok(Y.cpu().numpy(), "Y")
        ok(G.mu.cpu().numpy(), "mu")
        ok(G.w.cpu().numpy(), "w", 5e-5)
        ok(np.array(G.sigma), "sigma")
        ok(np.array(float(Cfe)), "Cfe", 5e-5)
        ok(np.array(fes), "FE", 5e-5)
        if G.outliers is not None:
            assert abs(G.outliers["eta0"] - float(g[f"{tag}_gold_eta0"])) < 1e-4


def test_argmax_assignments_bit_exact_on_tie_free_rows(golden):
    g = golden("gmm")
    for tag in ["2d_full", "3d_frozen", "3d_offset"]:
        G, X, _, _ = build(g, tag, "torch")
        lg = torch.from_numpy(g[f"{tag}_gold_lgam"])
        top2 = lg.topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > 1e-4            # rows whose decision does not hinge on fp32 rounding
        got = G.hard_assignments(X).cpu()
        assert got.dtype == torch.int64
        assert torch.equal(got[safe], lg.argmax(1)[safe])
        assert int((~safe).sum()) <= max(2, int(0.01 * len(safe))), (tag, int((~safe).sum()))
        assert torch.equal(G.log_responsibilities(X).argmax(1).cpu()[safe], lg.argmax(1)[safe])


def test_em_large_vs_oracle_both_variants():
    """640k x 50 is the atlas size of BASELINE configs[2]; checked here at 40k x 50 (oracle finishes in seconds) plus a
    size-independent property at full size: sum_c gamma_nc = 1 => Y inside the bounding box of mu, FE decreasing."""
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from oracle.gmm import GMMOracle
    g = torch.Generator().manual_seed(11)
    C, D, N = 50, 2, 40000
    cent = torch.rand(C, D, generator=g)
    X = cent[torch.randint(0, C, (N,), generator=g)] + 0.03 * torch.randn(N, D, generator=g)
    mu0 = cent + 0.02 * torch.randn(C, D, generator=g)
    for variant in ("keops", "torch"):
        G = GaussianMixtureUnif(mu0.to(dev()), sigma=0.05, spec=spec(), computversion=variant)
        O = GMMOracle(mu0.double(), 0.05)
        for _ in range(3):
            Y, Cfe, FE = G.EM_step(X.to(dev()))
            Yo, Cfeo, FEo = O.em_step(X.double(), variant=variant)
        assert relerr(Y.cpu().numpy(), Yo.numpy()) < 2e-5
        assert relerr(G.mu.cpu().numpy(), O.mu.numpy()) < 2e-5
        assert abs(G.sigma - O.sigma) < 2e-5 * O.sigma
        assert abs(float(FE) - float(FEo)) < 2e-5 * abs(float(FEo))
    # full size property run
    N = 640000
    X = (cent[torch.randint(0, C, (N,), generator=g)] + 0.03 * torch.randn(N, D, generator=g)).to(dev())
    G = GaussianMixtureUnif(mu0.to(dev()), sigma=0.05, spec=spec())
    last = None
    for _ in range(4):
        Y, Cfe, FE = G.EM_step(X)
        assert last is None or float(FE) <= last + 1e-6 * abs(last)
        last = float(FE)
    lo, hi = G.mu.min(0).values, G.mu.max(0).values
    assert bool(((Y >= lo - 1e-4) & (Y <= hi + 1e-4)).all())


@pytest.fixture(params=["small_support_stage_kernels", "general_pair_engine"])
def shoot_path(request):
    from diff_icp_b200 import ops, shooting
    old = ops.small_enabled
    ops.small_enabled = request.param == "small_support_stage_kernels"
    shooting.ShootPlan._cache.clear()
    shooting.ClosurePlan._cache.clear()
    yield request.param
    ops.small_enabled = old
    shooting.ShootPlan._cache.clear()
    shooting.ClosurePlan._cache.clear()


def test_c1_single_set_registration_matches_reference(golden, shoot_path):
    """BASELINE configs[0]-like: one 2-D set registered to a known GMM (sigma optimised), classic LDDMM, grid support."""
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.core.PSR import DiffPSR
    g = golden("psr")
    x0, mu0 = cu(g["c1_in_x0"]), cu(g["c1_in_mu"])
    # golden = the reference's torch twin, whose EM free energy uses the OLD sigma in the Gaussian normalisation: the
    # number of EM steps before the relative-FE stop depends on that, so the same semantic variant is selected here
    G = GaussianMixtureUnif(mu0, sigma=0.1, spec=spec(), computversion="torch")
    G.to_optimize = {"mu": False, "sigma": True, "w": False, "eta0": False}
    LM = LDDMMModel(sigma=0.2, D=2, lambd=5e2, version="classic", scheme="Euler", spec=spec())
    P = DiffPSR(x0, G, LM, dataspec=spec(), compspec=spec())
    P.printstuff = False
    P.set_support_scheme("grid", rho=np.sqrt(2))
    assert relerr(P.q0[0].cpu().numpy(), g["c1_gold_q0"]) < 1e-6
    fes, sigs = [], []
    for it in range(3):
        P.GMM_opt()
        fes.append(P.FE)
        P.Reg_opt(tol=1e-5)
        fes.append(P.FE)
        sigs.append(P.GMMi[0].sigma)
    # the reference's own fp32 run sits at ~1e-5 of its fp64 run on this problem; allow 10x that
    assert np.allclose(fes, g["c1_gold_FE"], rtol=2e-4), (fes, g["c1_gold_FE"])
    assert np.allclose(sigs, g["c1_gold_sigma"], rtol=2e-4)
    assert np.abs(P.x1[0, 0].cpu().numpy() - g["c1_gold_x1"]).max() < 1e-3 * 0.2
    # targets are softmax-weighted centroids: a point half-way between two centroids amplifies dx by |dmu|^2/sigma_GMM^2 ~ 20
    assert np.abs(P.y[0, 0].cpu().numpy() - g["c1_gold_y"]).max() < 5e-3 * 0.2


def test_atlas_api_matches_reference(golden, shoot_path):
    """Small groupwise atlas through ICP_atlas (3 frames, C = 6, hybrid, Euler, grid): L-BFGS paths of the reference's
    own fp32 and fp64 runs already differ by 0.17 % in FE here, so the bar is 'as close to gold as ref32 is'."""
    from diff_icp_b200.api.ICP_atlas import ICP_atlas
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    g = golden("psr")
    sets = [cu(g[f"atlas_in_x{k}"]) for k in range(3)]
    allx = torch.cat(sets)
    G = GaussianMixtureUnif(cu(g["atlas_in_mu"]), sigma=0.25 * float(allx.std()), spec=spec(), computversion="torch")
    PSR, evol = ICP_atlas(sets, GMM_parameters={"init_components": [G], "optimize_weights": True},
                          registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.2},
                          numerical_options={"computversion": "torch", "compspec": spec(), "dataspec": spec(),
                                             "support_LDDMM": {"scheme": "grid", "rho": 1.0}},
                          optim_options={"max_iterations": 3, "max_repeat_GMM": 10, "convergence_tolerance": 1e-3},
                          printstuff=False)
    gold, ref = float(g["atlas_gold_FE"]), float(g["atlas_ref32_FE"])
    assert abs(PSR.FE - gold) < 3 * abs(ref - gold) + 1e-3 * abs(gold), (PSR.FE, gold, ref)
    assert abs(PSR.GMMi[0].sigma - float(g["atlas_gold_sigma"])) < 3 * abs(float(g["atlas_ref32_sigma"]) - float(g["atlas_gold_sigma"])) + 1e-4
    assert len(evol["a0"]) == 3 and len(evol["GMMi"]) == 3


def test_two_set_api_runs_logdet_default_small():
    """api.ICP_two_set builds the full logdet model (reference quirk, api/ICP_two_set.py:203-207)."""
    from diff_icp_b200.api.ICP_two_set import ICP_two_set
    g = torch.Generator().manual_seed(3)
    xA = torch.rand(300, 3, generator=g)
    xB = xA + 0.02 * torch.randn(300, 3, generator=g) + 0.03
    PSR, evol = ICP_two_set(xA.to(dev()), xB.to(dev()), {"sigma": 0.1, "optimize_sigma": True, "outlier_weight": None},
                            {"type": "diffeomorphic", "lambda_LDDMM": 500.0, "sigma_LDDMM": 0.2},
                            numerical_options={"support_LDDMM": {"scheme": "dense"}},
                            optim_options={"max_iterations": 2}, plotstuff=False, printstuff=False)
    assert PSR.LMi.gradcomponent and PSR.LMi.eta == 1 / 500.0
    assert np.isfinite(PSR.FE)


def test_concurrent_frame_registration_is_bit_identical():
    """DiffPSR.Reg_opt with several frames in flight (threads + streams + per-slot CUDA graphs) must give exactly the
    results of the sequential loop: frames are independent and every frame's computation is deterministic."""
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.core.PSR import DiffPSR
    g = torch.Generator().manual_seed(21)
    cent = torch.rand(6, 2, generator=g)
    frames = [(cent[torch.randint(0, 6, (400 + 13 * k,), generator=g)] + 0.03 * torch.randn(400 + 13 * k, 2, generator=g)).to(dev())
              for k in range(6)]
    outs = []
    for workers, graph in ((1, False), (4, True), (6, True)):
        G = GaussianMixtureUnif(cent.to(dev()) + 0.02, sigma=0.08, spec=spec())
        LM = LDDMMModel(sigma=0.25, D=2, lambd=200.0, version="hybrid", scheme="Euler", nt=8, spec=spec())
        LM.use_cuda_graph = graph
        P = DiffPSR(frames, G, LM, dataspec=spec(), compspec=spec())
        P.printstuff = False
        P.frame_workers = workers
        P.set_support_scheme("grid", rho=1.0)
        for _ in range(2):
            P.GMM_opt(max_iterations=5, tol=1e-4)
            P.Reg_opt(nmax=1, tol=1e-3)
        outs.append((P.FE, [a.cpu() for a in P.a0], [P.x1[k, 0].cpu() for k in range(6)]))
    for o in outs[1:]:
        assert o[0] == outs[0][0]
        assert all(torch.equal(a, b) for a, b in zip(o[1], outs[0][1]))
        assert all(torch.equal(a, b) for a, b in zip(o[2], outs[0][2]))


@pytest.mark.parametrize("D,N,C", [(2, 1, 1), (2, 511, 7), (3, 512, 8), (2, 513, 50), (3, 70001, 33), (2, 200000, 64),
                                   (3, 3000, 65), (2, 40000, 50)])
def test_fused_first_sweep_equals_row_lse_then_column_statistics(D, N, C):
    """dicp_em_lse_colstats (row log-sum-exp + column statistics from ONE read of X when C <= 64, csrc/em_col_small.cuh)
    against the two separate sweeps and against a float64 evaluation of the same sums (core/GMM.py:410-415, :443-455).
    The reference exponents m_c of the two forms may differ (other merge orders), the statistics they stand for may not:
    log2 mass m + log2 S0, first moment B / S0, second moment A / S0.  One component sits far from every point (mass
    ~2^-400: representable only in the log domain)."""
    from diff_icp_b200 import em_ops
    g = torch.Generator().manual_seed(N + C)
    sig = 0.07
    X = torch.rand(N, D, generator=g)
    mu = torch.rand(C, D, generator=g)
    if C > 2:
        mu[C // 2] = 1.0 + 1.7           # dead component
    w = torch.randn(C, generator=g)
    lgn = D * (np.log(sig) + 0.5 * np.log(2 * np.pi))
    wl2 = ((w - torch.logsumexp(w, 0) - lgn) * 1.4426950408889634).contiguous()
    Xd, mud, wd = X.to(dev()), mu.to(dev()), wl2.to(dev())
    fused = em_ops.lse_colstats(sig, Xd, mud, wd).cpu().double()
    two = em_ops.colstats(sig, Xd, em_ops.rowpass(sig, Xd, mud, wd), mud, wd).cpu().double()
    # float64 evaluation
    X64, mu64 = X.double(), mu.double()
    t = wl2.double()[None, :] * np.log(2.0) - ((X64[:, None, :] - mu64[None]) ** 2).sum(-1) / (2 * sig * sig)
    lg = t - torch.logsumexp(t, 1, keepdim=True)                     # ln gamma_nc
    lmass = torch.logsumexp(lg, 0) / np.log(2.0)
    gam = torch.exp(lg - lg.max(0, keepdim=True).values)             # column-normalised, like the kernels
    S = gam.sum(0)
    d = X64[:, None, :] - mu64[None]
    B = (gam[:, :, None] * d).sum(0) / S[:, None]
    A = (gam * (d ** 2).sum(-1)).sum(0) / S

    def norm(st):
        return st[:, 0] + torch.log2(st[:, 1]), st[:, 2:2 + D] / st[:, 1:2], st[:, 2 + D] / st[:, 1]
    for st in (fused, two):
        assert torch.isfinite(st).all()
        lm, b, a = norm(st)
        assert (lm - lmass).abs().max() <= 2e-4 * max(1.0, lmass.abs().max().item())
        assert (b - B).abs().max() <= 2e-5
        assert (a - A).abs().max() <= 2e-5 * max(1.0, A.abs().max().item())
    lf, bf, af = norm(fused)
    lt, bt, at = norm(two)
    assert (lf - lt).abs().max() <= 1e-4 and (bf - bt).abs().max() <= 1e-5 and (af - at).abs().max() <= 1e-5


@pytest.mark.parametrize("D,C", [(2, 50), (3, 7), (2, 700)])
def test_allreduce_buffer_and_merged_mstep(D, C):
    """dicp_em_reduce_pack / dicp_em_mstep_merged (the two launches around the multi-GPU EM step's one all-reduce) against
    the element-wise formulas they replace (dist.StatsComm.merge_colstats_ref + dicp_em_mstep, core/GMM.py:286-297)."""
    from diff_icp_b200 import em_ops
    g = torch.Generator().manual_seed(C)
    stats = torch.rand(C, D + 3, generator=g)
    stats[:, 0] = torch.randint(-40, 5, (C,), generator=g).float() + torch.rand(C, generator=g)
    stats[:, 2:2 + D] -= 0.5
    m_ref = torch.round(stats[:, 0] + torch.randint(-3, 4, (C,), generator=g).float())
    extra = torch.rand(5, generator=g)
    mu, w = torch.rand(C, D, generator=g), torch.randn(C, generator=g)
    sd, md, ed, mud, wd = (t.to(dev()) for t in (stats, m_ref, extra, mu, w))
    for overflow in (False, True):
        if overflow:
            sd = sd.clone()
            sd[C // 2, 0] = md[C // 2] + 101.0
        buf = em_ops.reduce_pack(sd, md, ed)
        d = sd[:, 0] - md
        scaled = sd[:, 1:] * torch.exp2(torch.clamp(d, max=120.0))[:, None]
        assert buf.numel() == C * (D + 2) + 6
        assert torch.allclose(buf[:C * (D + 2)].view(C, D + 2), scaled, rtol=1e-6, atol=0)
        assert float(buf[C * (D + 2)]) == (1.0 if overflow else 0.0)
        assert torch.equal(buf[C * (D + 2) + 1:], ed)
    buf = em_ops.reduce_pack(stats.to(dev()), md, ed) * 2.0             # "two ranks with the same statistics"
    for sig_mode in (0, 1, 2):
        mu_new, w_new, lpi_new, m_next, host = em_ops.mstep_merged(buf, md, mud, wd, True, True, sig_mode, 5)
        merged = torch.cat((md[:, None], buf[:C * (D + 2)].view(C, D + 2)), dim=1).contiguous()
        mu_r, w_r, lpi_r, ms = em_ops.mstep(merged, mud, wd, True, True, sig_mode)
        assert torch.equal(mu_new, mu_r) and torch.equal(w_new, w_r) and torch.equal(lpi_new, lpi_r)
        assert torch.equal(m_next, torch.round(md + torch.log2(torch.clamp(merged[:, 1], min=1e-30))))
        h = host.tolist()
        assert h[0] == float(ms[0]) and h[1:6] == (2.0 * ed).tolist() and h[6] == 0.0 and h[7] == 0.0
    buf[3 * (D + 2)] = 0.0                                               # a vanished column mass is reported
    assert em_ops.mstep_merged(buf, md, mud, wd, True, True, 0, 5)[4].tolist()[7] == 1.0


@pytest.mark.parametrize("version,opt,tol,D", [
    ("keops", dict(mu=True, sigma=True, w=False, eta0=True), 1e-3, 2),       # atlas settings
    ("keops", dict(mu=True, sigma=True, w=True, eta0=True), 1e-2, 3),        # stops early
    ("torch", dict(mu=True, sigma=True, w=True, eta0=True), 1e-5, 2),        # sigma from the old centroids' distances
    ("keops", dict(mu=True, sigma=False, w=True, eta0=True), None, 3),       # fixed sigma, never stops
    ("torch", dict(mu=False, sigma=True, w=True, eta0=True), 1e-4, 2),
])
def test_em_loop_as_one_cuda_graph_equals_the_step_by_step_loop(version, opt, tol, D):
    """GaussianMixtureUnif.EM_optimization with the loop state on the device (em_loop.EMLoopGraph: max_iterations steps
    enqueued as one CUDA graph, stop flag tested by the kernels) against the host loop around EM_step_b200 -- same kernels,
    same arithmetic: identical bits for mu, w, sigma, targets, Cfe, FE and the step count.  Second and third calls replay
    the captured graph from the model the first call left."""
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    g = torch.Generator().manual_seed(11)
    C, N = 9, 20011
    cen = torch.rand(C, D, generator=g)
    X = (cen[torch.randint(0, C, (N,), generator=g)] + 0.04 * torch.randn(N, D, generator=g)).to(dev())
    mu0 = X[torch.randint(0, N, (C,), generator=g)].clone()

    def run(graph):
        G = GaussianMixtureUnif(mu0.clone(), sigma=0.2, spec=spec(), computversion=version)
        G.to_optimize = dict(opt)
        G.graph_em_loop = graph
        out = []
        for it, nmax in enumerate((7, 5, 9)):
            Xi = X + 0.001 * it
            Y, Cfe, FE, steps = G.EM_optimization(Xi, max_iterations=nmax, tol=tol)
            out.append((Y.clone(), float(Cfe), float(FE), steps, G.mu.clone(), G.w.clone(), float(G.sigma)))
        assert (getattr(G, "_em_loop", None) is not None) == graph
        return out
    a, b = run(True), run(False)
    for ra, rb in zip(a, b):
        assert ra[3] == rb[3], (ra[3], rb[3])
        assert ra[1] == rb[1] and ra[2] == rb[2] and ra[6] == rb[6], (ra[1:4], rb[1:4], ra[6], rb[6])
        assert torch.equal(ra[0], rb[0]) and torch.equal(ra[4], rb[4]) and torch.equal(ra[5], rb[5])
    if tol == 1e-2:
        assert any(r[3] < n for r, n in zip(a, (7, 5, 9)))         # the early stop was exercised
