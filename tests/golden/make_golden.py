#!/usr/bin/env python
"""Generate the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (the reference lives at /root/reference and does not
travel to the GPU box):

    python tests/golden/make_golden.py

The reference package is imported as-is through a loader that
  * stubs matplotlib / mpl_toolkits (not installed; only used for plotting),
  * provides a pykeops-free stand-in for diffICP.tools.point_sets (its line 8 hard-imports
    pykeops, which is absent), restating intrinsic_scale / decimate with torch,
  * leaves pykeops absent so every "keops" request falls back to the reference's own
    torch twin (tools/kernel.py:93-96, core/GMM.py:130-133).
Inputs are fp32-representable; every case is evaluated by the reference in fp32
("ref32") and in fp64 ("gold").  All seeds are fixed.
"""

import os
import sys
import types
import warnings

import numpy as np
import torch

REF_ROOT = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    """The shared loader (oracle/ref_loader.py): stubs for matplotlib, a pykeops-free point_sets module, torch twin."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle.ref_loader import load_reference as _load
    ref = _load(REF_ROOT)
    return ref.kernel, ref.LDDMM, ref.GMM


def spec_of(dt):
    return {"device": "cpu", "dtype": dt}


def np32(t):
    return t.detach().to(torch.float32).numpy() if t.dtype == torch.float32 else t.detach().numpy()


def gen_kernels(rk):
    """All ten reductions, at the author's own self-check size (tools/kernel.py:352) and a 3-D case."""
    out = {}
    for tag, (M, N, D, sig, seed) in {"a": (100, 1000, 2, 2.0, 11), "b": (64, 200, 3, 0.7, 12),
                                      "c": (33, 77, 3, 0.25, 13)}.items():
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(M, D, generator=g)
        y = torch.randn(N, D, generator=g)
        b = torch.randn(N, D, generator=g)
        c = torch.randn(M, D, generator=g)
        d = torch.randn(N, generator=g)
        out[f"{tag}_meta"] = np.array([M, N, D, sig], dtype=np.float64)
        for nm, t in [("x", x), ("y", y), ("b", b), ("c", c), ("d", d)]:
            out[f"{tag}_in_{nm}"] = t.numpy()
        for prec, dt in [("ref32", torch.float32), ("gold", torch.float64)]:
            GK = rk.GaussKernel(sig, D, computversion="torch", spec=spec_of(dt))
            X, Y, B, Cc, Dd = (t.to(dt) for t in (x, y, b, c, d))
            res = {
                "KBase": GK.KBase(X, Y), "KRedScal": GK.KRedScal(X, Y, Dd), "KRed": GK.KRed(X, Y, B),
                "GradKRed": GK.GradKRed(X, Y), "DDKRed": GK.DDKRed(X, Y, B),
                "GenDKRed": GK.GenDKRed(X, Y, B, Cc), "HessKRed": GK.HessKRed(X, Y, B, Cc),
                "LapKRed": GK.LapKRed(X, Y), "GradLapKRed": GK.GradLapKRed(X, Y),
                "GradKRed_rev": GK.GradKRed_rev(X, Y, Cc),
            }
            for k, v in res.items():
                out[f"{tag}_{prec}_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), **out)
    print("kernels.npz", len(out))


def gen_lddmm(rl):
    """Shoot + trajloss + quadratic data loss + gradients for every model variant / scheme / x mode."""
    out = {}
    cases = []
    for D in (2, 3):
        for version in ("classic", "hybrid", "logdet"):
            for scheme in ("Euler", "Ralston"):
                for with_x in (False, True):
                    cases.append((D, version, scheme, with_x))
    names = []
    for ci, (D, version, scheme, with_x) in enumerate(cases):
        g = torch.Generator().manual_seed(100 + ci)
        Nq, Nx, nt = 37, 91, 5
        sig, lam = 0.35, 7.0
        q0 = torch.rand(Nq, D, generator=g)
        p0 = 0.3 * torch.randn(Nq, D, generator=g)
        x0 = torch.rand(Nx, D, generator=g) if with_x else None
        ny = Nx if with_x else Nq
        y = torch.rand(ny, D, generator=g)
        sig2 = 0.05 + 0.1 * torch.rand(ny, generator=g)          # per-point GMM sigma^2 (core/PSR.py:511)
        tag = f"{D}d_{version}_{scheme}_{'x' if with_x else 'nox'}"
        names.append(tag)
        out[f"{tag}_meta"] = np.array([D, Nq, Nx if with_x else 0, nt, sig, lam], dtype=np.float64)
        out[f"{tag}_in_q0"], out[f"{tag}_in_p0"] = q0.numpy(), p0.numpy()
        out[f"{tag}_in_y"], out[f"{tag}_in_sig2"] = y.numpy(), sig2.numpy()
        if with_x:
            out[f"{tag}_in_x0"] = x0.numpy()
        for prec, dt in [("ref32", torch.float32), ("gold", torch.float64)]:
            LM = rl.LDDMMModel(sigma=sig, D=D, lambd=lam, spec=spec_of(dt), version=version,
                               computversion="torch", scheme=scheme, nt=nt)
            q = q0.to(dt).requires_grad_(True)
            p = p0.to(dt).requires_grad_(True)
            xx = x0.to(dt).requires_grad_(True) if with_x else None
            sh = LM.Shoot(q, p, xx)
            tl = LM.trajloss(sh)
            moved = sh[-1][3] if with_x else sh[-1][0]
            dl = ((moved - y.to(dt)) ** 2 / (2 * sig2.to(dt)[:, None])).sum()
            L = tl + dl
            grads = torch.autograd.grad(L, [q, p] + ([xx] if with_x else []))
            out[f"{tag}_{prec}_q1"] = sh[-1][0].detach().numpy()
            out[f"{tag}_{prec}_p1"] = sh[-1][1].detach().numpy()
            out[f"{tag}_{prec}_cost1"] = sh[-1][2].detach().numpy()
            if with_x:
                out[f"{tag}_{prec}_x1"] = sh[-1][3].detach().numpy()
                out[f"{tag}_{prec}_gx0"] = grads[2].numpy()
            out[f"{tag}_{prec}_qmid"] = sh[nt // 2][0].detach().numpy()
            out[f"{tag}_{prec}_trajloss"] = np.array(float(tl))
            out[f"{tag}_{prec}_loss"] = np.array(float(L))
            out[f"{tag}_{prec}_H0"] = np.array(float(LM.Hamiltonian(q, p)))
            out[f"{tag}_{prec}_gq0"] = grads[0].numpy()
            out[f"{tag}_{prec}_gp0"] = grads[1].numpy()
            # single right-hand-side evaluation (core/LDDMM.py:176-227)
            ode = LM.ODE(q.detach(), p.detach(), torch.zeros(1, dtype=dt), None if not with_x else xx.detach())
            out[f"{tag}_{prec}_ode_vq"] = ode[0].numpy()
            out[f"{tag}_{prec}_ode_dp"] = ode[1].numpy()
            out[f"{tag}_{prec}_ode_dcost"] = np.array(float(ode[2].sum()))
            if with_x:
                out[f"{tag}_{prec}_ode_vx"] = ode[3].numpy()
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "lddmm.npz"), **out)
    print("lddmm.npz", len(out))


def gen_gmm(rg):
    """EM_step_torch (the executable twin): skip_M, full M step, outliers, frozen mu/w, EM_optimization."""
    out = {}
    names = []
    cfgs = [
        # tag, D, N, C, outliers, to_optimize, skip_M, steps
        ("2d_full", 2, 500, 20, False, dict(mu=True, sigma=True, w=True, eta0=True), False, 1),
        ("2d_skipM", 2, 500, 20, False, dict(mu=True, sigma=True, w=True, eta0=True), True, 1),
        ("3d_full", 3, 400, 13, False, dict(mu=True, sigma=True, w=True, eta0=True), False, 1),
        ("3d_frozen", 3, 300, 150, False, dict(mu=False, sigma=True, w=False, eta0=False), False, 1),
        ("2d_outl", 2, 500, 20, True, dict(mu=True, sigma=True, w=True, eta0=True), False, 1),
        ("2d_outl_skipM", 2, 500, 20, True, dict(mu=True, sigma=True, w=True, eta0=True), True, 1),
        ("2d_now", 2, 500, 20, False, dict(mu=True, sigma=True, w=False, eta0=True), False, 1),
        ("2d_opt5", 2, 500, 20, False, dict(mu=True, sigma=True, w=True, eta0=True), False, 5),
        ("3d_offset", 3, 600, 10, False, dict(mu=True, sigma=True, w=True, eta0=True), False, 1),
    ]
    for ci, (tag, D, N, C, outl, topt, skip, steps) in enumerate(cfgs):
        g = torch.Generator().manual_seed(300 + ci)
        cent = torch.rand(C, D, generator=g)
        if tag == "3d_offset":
            cent = cent + 5.0
        X = cent[torch.randint(0, C, (N,), generator=g)] + 0.05 * torch.randn(N, D, generator=g)
        mu0 = cent + 0.03 * torch.randn(C, D, generator=g)
        w0 = 0.3 * torch.randn(C, generator=g)
        sig0 = 0.08
        names.append(tag)
        out[f"{tag}_meta"] = np.array([D, N, C, int(outl), int(skip), steps, sig0], dtype=np.float64)
        out[f"{tag}_opt"] = np.array([int(topt[k]) for k in ("mu", "sigma", "w", "eta0")])
        out[f"{tag}_in_X"], out[f"{tag}_in_mu"], out[f"{tag}_in_w"] = X.numpy(), mu0.numpy(), w0.numpy()
        for prec, dt in [("ref32", torch.float32), ("gold", torch.float64)]:
            G = rg.GaussianMixtureUnif(mu0.to(dt), sigma=sig0, use_outliers=outl, spec=spec_of(dt),
                                       computversion="torch")
            G.w = w0.to(dt)
            G.to_optimize = dict(topt)
            if outl:
                G.outliers["eta0"] = -1.0
            Xd = X.to(dt)
            if prec == "gold":
                out[f"{tag}_gold_lgam"] = G.log_responsibilities(Xd).numpy()
            fes = []
            for _ in range(steps):
                Y, Cfe, FE = G.EM_step(Xd, skip_M=skip)
                fes.append(float(FE))
            out[f"{tag}_{prec}_Y"] = Y.numpy()
            out[f"{tag}_{prec}_Cfe"] = np.array(float(Cfe))
            out[f"{tag}_{prec}_FE"] = np.array(fes)
            out[f"{tag}_{prec}_mu"] = G.mu.numpy()
            out[f"{tag}_{prec}_w"] = G.w.numpy()
            out[f"{tag}_{prec}_sigma"] = np.array(float(G.sigma))
            if outl:
                out[f"{tag}_{prec}_eta0"] = np.array(float(G.outliers["eta0"]))
                out[f"{tag}_{prec}_vol0"] = np.array(float(G.outliers["vol0"]))
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "gmm.npz"), **out)
    print("gmm.npz", len(out))




def spiral_sets(K, N, seed, D=2):
    """Deterministic spiral-GMM samples, smoothly warped per frame (no reference RNG involved)."""
    g = torch.Generator().manual_seed(seed)
    C = 20
    t = torch.linspace(0, 2 * np.pi, C + 1)[:-1]
    mu0 = torch.stack((0.5 + 0.4 * (t / 7) * t.cos(), 0.5 + 0.3 * t.sin()), 1)
    sets = []
    for k in range(K):
        c = torch.randint(0, C, (N,), generator=g)
        x = mu0[c] + 0.025 * torch.randn(N, 2, generator=g)
        cen = torch.rand(3, 2, generator=g)
        amp = 0.04 * torch.randn(3, 2, generator=g)
        w = torch.exp(-((x[:, None, :] - cen[None]) ** 2).sum(-1) / (2 * 0.25 ** 2))
        sets.append((x + w @ amp).contiguous())
    return sets, mu0


def gen_psr():
    """End-to-end DiffPSR runs of the unmodified reference (torch path, CPU): a C1-like single set registered to a known
    GMM, and a small groupwise atlas through api.ICP_atlas."""
    import diffICP.core.PSR as rp
    import diffICP.core.GMM as rg
    import diffICP.core.LDDMM as rl
    from diffICP.tools.kernel import GaussKernel
    # the torch branch of check_coverage is broken in the reference (tools/kernel.py:328); one-line fix of SURVEY App. C
    GaussKernel.check_coverage = lambda self, X, Y, R: ((X[:, None, :] - Y[None, :, :]) ** 2).sum(-1).min(dim=1).values > (R * self.sigma) ** 2
    out = {}
    # ---- C1-like: one set, known GMM (mu, w frozen; sigma optimised), classic LDDMM on a grid support ----------
    sets, mu0 = spiral_sets(1, 300, 77)
    x0 = sets[0]
    out["c1_in_x0"], out["c1_in_mu"] = x0.numpy(), mu0.numpy()
    for prec, dt in [("ref32", torch.float32), ("gold", torch.float64)]:
        sp = spec_of(dt)
        G = rg.GaussianMixtureUnif(mu0.to(dt), sigma=0.1, spec=sp, computversion="torch")
        G.to_optimize = {"mu": False, "sigma": True, "w": False, "eta0": False}
        LM = rl.LDDMMModel(sigma=0.2, D=2, lambd=5e2, version="classic", computversion="torch", scheme="Euler", spec=sp)
        P = rp.DiffPSR([[x0.to(dt)]], G, LM, dataspec=sp, compspec=sp)   # list-of-lists: read_point_sets only type-checks bare tensors
        P.printstuff = False
        P.set_support_scheme("grid", rho=np.sqrt(2))
        fes, sigs = [], []
        for it in range(3):
            P.GMM_opt()
            fes.append(float(P.FE))
            P.Reg_opt(tol=1e-5)
            fes.append(float(P.FE))
            sigs.append(float(P.GMMi[0].sigma))
        out[f"c1_{prec}_FE"], out[f"c1_{prec}_sigma"] = np.array(fes), np.array(sigs)
        out[f"c1_{prec}_x1"] = P.x1[0, 0].numpy()
        out[f"c1_{prec}_q0"] = P.q0[0].numpy()
        out[f"c1_{prec}_a0"] = P.a0[0].numpy()
        out[f"c1_{prec}_y"] = P.y[0, 0].numpy()
    # ---- small atlas through the API: 3 frames, C = 6 given initial centroids, hybrid, Euler, grid --------------
    import diffICP.api.ICP_atlas as ra
    sets, mu0 = spiral_sets(3, 150, 78)
    g = torch.Generator().manual_seed(5)
    mu_init = torch.cat(sets).mean(0) + 0.05 * torch.cat(sets).std() * torch.randn(6, 2, generator=g)
    for k, s in enumerate(sets):
        out[f"atlas_in_x{k}"] = s.numpy()
    out["atlas_in_mu"] = mu_init.numpy()
    for prec, dt in [("ref32", torch.float32), ("gold", torch.float64)]:
        sp = spec_of(dt)
        G = rg.GaussianMixtureUnif(mu_init.to(dt), sigma=0.25 * float(torch.cat(sets).std()), spec=sp, computversion="torch")
        PSR, evol = ra.ICP_atlas([[s.to(dt)] for s in sets],
                                 GMM_parameters={"init_components": [G], "optimize_weights": True},
                                 registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.2},
                                 numerical_options={"computversion": "torch", "compspec": sp, "dataspec": sp,
                                                    "support_LDDMM": {"scheme": "grid", "rho": 1.0}},
                                 optim_options={"max_iterations": 3, "max_repeat_GMM": 10, "convergence_tolerance": 1e-3},
                                 printstuff=False)
        out[f"atlas_{prec}_FE"] = np.array(float(PSR.FE))
        out[f"atlas_{prec}_sigma"] = np.array(float(PSR.GMMi[0].sigma))
        out[f"atlas_{prec}_mu"] = PSR.GMMi[0].mu.numpy()
        out[f"atlas_{prec}_w"] = PSR.GMMi[0].w.numpy()
        for k in range(3):
            out[f"atlas_{prec}_x1_{k}"] = PSR.x1[k, 0].numpy()
    np.savez_compressed(os.path.join(OUT, "psr.npz"), **out)
    print("psr.npz", len(out))


def gen_v2p(rl):
    """LDDMMModel.v2p (core/LDDMM.py:235-253) -> KpinvSolve (tools/kernel.py:227-232, numpy lstsq with a relative
    singular-value cut-off) and KridgeSolve_torch (:234-237): eta = 0 and eta != 0, zero and non-zero target speeds,
    rcond = 1e-3 / 1e-1 / None, plus the author's own round trip v -> p -> v (core/LDDMM.py:809-813)."""
    out, names = {}, []
    for D, M, sig, lam, seed in ((2, 300, 0.25, 7.0, 501), (3, 280, 0.3, 20.0, 502), (2, 10, 2.0, 100.0, 503)):
        g = torch.Generator().manual_seed(seed)
        q = torch.rand(M, D, generator=g) if M > 10 else torch.randn(M, D, generator=g)
        b = 0.3 * torch.randn(M, D, generator=g)
        for version in ("classic", "logdet"):
            tag = f"{D}d_M{M}_{version}"
            names.append(tag)
            out[f"{tag}_meta"] = np.array([D, M, sig, lam], dtype=np.float64)
            out[f"{tag}_in_q"], out[f"{tag}_in_b"] = q.numpy(), b.numpy()
            for prec, dt in [("ref32", torch.float32), ("gold", torch.float64)]:
                LM = rl.LDDMMModel(sigma=sig, D=D, lambd=lam, spec=spec_of(dt), version=version, computversion="torch")
                Q, B = q.to(dt), b.to(dt)
                v = LM.v(Q, Q, B)
                out[f"{tag}_{prec}_v"] = v.numpy()
                for rc_tag, rc in (("rc3", 1e-3), ("rc1", 1e-1), ("rcNone", None)):
                    p = LM.v2p(Q, v, rcond=rc)
                    out[f"{tag}_{prec}_p_{rc_tag}"] = p.numpy()
                    out[f"{tag}_{prec}_vback_{rc_tag}"] = LM.v(Q, Q, p).numpy()
                    p0 = LM.v2p(Q, torch.zeros_like(Q), rcond=rc)          # what DiffPSR.initialize_a0 asks for
                    out[f"{tag}_{prec}_pzero_{rc_tag}"] = p0.numpy()
                    out[f"{tag}_{prec}_vzero_{rc_tag}"] = LM.v(Q, Q, p0).numpy()
                # ridge solve (the reference's own dense torch form; its 'ridge_pytorch' dispatch name is broken, LDDMM.py:251)
                rhs = v + LM.eta * LM.Kernel.GradKRed(Q, Q)
                for a_tag, alpha in (("a2", 1e-2), ("a4", 1e-4)):
                    K_xx = LM.Kernel.K_torch(Q, Q)
                    pr = torch.linalg.solve(K_xx + alpha * torch.eye(M, dtype=dt), rhs)   # tools/kernel.py:234-237
                    out[f"{tag}_{prec}_pridge_{a_tag}"] = pr.numpy()
                # singular values of K(q,q): tells the tests how far the cut-offs sit from the nearest singular value
                if prec == "gold":
                    out[f"{tag}_gold_svals"] = torch.linalg.svdvals(LM.Kernel.K_torch(Q, Q)).numpy()
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "v2p.npz"), **out)
    print("v2p.npz", len(out))


class _fp64_defaults:
    """Run reference code that hard-wires `defspec` (api/ICP_two_set.py builds its models without a spec argument) in
    fp64: every module holds a reference to the SAME dict object, so its dtype entry is switched in place."""

    def __enter__(self):
        import diffICP.tools.spec as rs
        import diffICP.core.PSR as rp
        self.rs, self.rp = rs, rp
        self.old = rs.defspec["dtype"]
        rs.defspec["dtype"] = torch.float64
        # read_point_sets type-checks bare tensors against torch.FloatTensor only (tools/in_out.py:21): wrap a bare fp64
        # tensor the way it would wrap an fp32 one
        self.old_read = rp.read_point_sets
        rp.read_point_sets = lambda x: self.old_read([[x]] if isinstance(x, torch.Tensor) else x)

    def __exit__(self, *a):
        self.rs.defspec["dtype"] = self.old
        self.rp.read_point_sets = self.old_read


FE_TRACE = []


def _patch_coverage():
    from diffICP.tools.kernel import GaussKernel
    import diffICP.core.PSR as rp
    if not hasattr(rp.MultiPSR, "_orig_update_FE"):      # harness-side recording of every free-energy update
        rp.MultiPSR._orig_update_FE = rp.MultiPSR.update_FE

        def update_FE(self, message=None):
            rp.MultiPSR._orig_update_FE(self, message=message)
            FE_TRACE.append(float(self.FE))
        rp.MultiPSR.update_FE = update_FE
    # the torch branch of check_coverage is broken in the reference (tools/kernel.py:328); one-line fix of SURVEY App. C
    GaussKernel.check_coverage = lambda self, X, Y, R: ((X[:, None, :] - Y[None, :, :]) ** 2).sum(-1).min(dim=1).values > (R * self.sigma) ** 2


def gen_two_set(cases=None, fname="two_set.npz"):
    """api.ICP_two_set of the unmodified reference (api/ICP_two_set.py:73-288): 3-D clouds, dense support, API-default
    full logdet model, sigma optimised; 3 outer iterations.  Two cases: no outliers / optimised
    outlier weight is NOT run (device-less zeros in log_ratio_to_proba make it CPU-only in the reference anyway, fine here)."""
    import contextlib
    import diffICP.api.ICP_two_set as rt
    _patch_coverage()
    out = {}
    g = torch.Generator().manual_seed(611)
    NA, NB = 400, 300
    xA = torch.rand(NA, 3, generator=g)
    cen = torch.rand(4, 3, generator=g)
    amp = 0.05 * torch.randn(4, 3, generator=g)
    xall = torch.cat((xA, torch.rand(NB, 3, generator=g)))
    warp = torch.exp(-((xall[:, None, :] - cen[None]) ** 2).sum(-1) / (2 * 0.3 ** 2)) @ amp
    xB = (xall + warp)[torch.randperm(NA + NB, generator=g)[:NB]] + 0.01 * torch.randn(NB, 3, generator=g)
    out["in_xA"], out["in_xB"] = xA.numpy(), xB.contiguous().numpy()
    for case, gmm_par, num_opt in cases or (
            ("dense", {"sigma": 0.1, "optimize_sigma": True, "outlier_weight": None}, {"support_LDDMM": {"scheme": "dense"}}),
            ("decim", {"sigma": 0.1, "optimize_sigma": True, "outlier_weight": None}, {"support_LDDMM": {"scheme": "decim", "rho": 1.0}}),
    ):
        for prec, dt in [("ref32", torch.float32), ("gold", torch.float64)]:
            ctx = _fp64_defaults() if dt == torch.float64 else contextlib.nullcontext()
            with ctx:
                FE_TRACE.clear()
                PSR, evol = rt.ICP_two_set(xA.to(dt), xB.to(dt), dict(gmm_par),
                                           {"type": "diffeomorphic", "lambda_LDDMM": 500.0, "sigma_LDDMM": 0.2},
                                           numerical_options=dict(num_opt, computversion="torch"),
                                           optim_options={"max_iterations": 3, "convergence_tolerance": 1e-3},
                                           plotstuff=False, printstuff=False)
            assert PSR.LMi.gradcomponent and PSR.LMi.eta == 1 / 500.0
            out[f"{case}_{prec}_FE"] = np.array(float(PSR.FE))
            out[f"{case}_{prec}_sigma"] = np.array(float(PSR.GMMi[0].sigma))
            out[f"{case}_{prec}_x1"] = PSR.x1[0, 0].numpy()
            out[f"{case}_{prec}_y"] = PSR.y[0, 0].numpy()
            out[f"{case}_{prec}_q0"] = PSR.q0[0].numpy()
            out[f"{case}_{prec}_a0_init"] = evol["a0"][0][0].numpy()       # what initialize_a0 / update_a0 (v2p) produced
            out[f"{case}_{prec}_a0"] = PSR.a0[0].numpy()
            out[f"{case}_{prec}_regloss"] = np.array(float(PSR.regloss[0]))
            out[f"{case}_{prec}_quadloss"] = np.array(float(np.asarray(PSR.quadloss, dtype=np.float64).sum()))
            out[f"{case}_{prec}_sigma_evol"] = np.array([float(G.sigma) for G in evol["GMMi"]])
            out[f"{case}_{prec}_FE_trace"] = np.array(FE_TRACE)           # after every GMM_opt / Reg_opt (and set-up) update
            for it in range(len(evol["a0"])):
                out[f"{case}_{prec}_a0_it{it}"] = evol["a0"][it][0].numpy()
    np.savez_compressed(os.path.join(OUT, fname), **out)
    print(fname, len(out), {k: out[k].shape for k in out if k.endswith("gold_q0")})


def gen_two_set_fine():
    """The same two clouds with a finer decimated support (spacing 0.7 sigma_LDDMM: between 64 and 512 support points, the
    mid-size regime of the stage kernels; written to its own file so that two_set.npz stays byte-identical)."""
    gen_two_set(cases=(("decimfine", {"sigma": 0.1, "optimize_sigma": True, "outlier_weight": None},
                        {"support_LDDMM": {"scheme": "decim", "rho": 0.7}}),), fname="two_set_fine.npz")


def structure_sets(K, Ns, seed):
    """K frames x 3 structures in 3-D: the three curves of examples/diffICP_full.py:44-56 lifted to 3-D (z = a slow wave
    along the curve), sigma 0.025 / 0.04 / 0.2, each frame smoothly warped; ragged sizes.  No reference RNG involved."""
    g = torch.Generator().manual_seed(seed)
    C = 20
    t = torch.linspace(0, 2 * np.pi, C + 1)[:-1]
    mus = [torch.stack((0.5 + 0.4 * (t / 7) * t.cos(), 0.5 + 0.3 * t.sin(), 0.3 + 0.1 * t.cos()), 1),
           torch.stack((1 + 0.4 * t.cos(), 0.5 + 0.4 * t.sin(), 0.5 + 0.2 * t.sin()), 1),
           torch.stack((0.8 + 0.1 * (t - np.pi), -0.06 * (t - np.pi), 0.4 + 0.05 * (t - np.pi)), 1)]
    sigs = (0.025, 0.04, 0.2)
    frames = []
    for k in range(K):
        cen = torch.rand(3, 3, generator=g) * torch.tensor([1.5, 1.0, 0.8])
        amp = 0.05 * torch.randn(3, 3, generator=g)
        fr = []
        for s in range(3):
            n = Ns + 7 * k + 3 * s
            x = mus[s][torch.randint(0, C, (n,), generator=g)] + sigs[s] * torch.randn(n, 3, generator=g)
            w = torch.exp(-((x[:, None, :] - cen[None]) ** 2).sum(-1) / (2 * 0.3 ** 2))
            fr.append((x + w @ amp).contiguous())
        frames.append(fr)
    return frames, mus, sigs


def gen_atlas_s3(K=3, Ns=90, seed=712, rho=1.0, fname="atlas_s3.npz"):
    """api.ICP_atlas with S = 3 structures (core/PSR.py:242-271 per-structure GMM loop, :498-516 per-structure sigma in
    the data loss), 3 frames, 3-D, decimated support (the reference's grid is 2-D only), hybrid model, Euler."""
    import diffICP.core.GMM as rg
    import diffICP.api.ICP_atlas as ra
    _patch_coverage()
    out = {}
    frames, mus, sigs = structure_sets(K, Ns, seed)
    g = torch.Generator().manual_seed(9)
    Cs = (6, 5, 4)
    mu_init = []
    for s in range(3):
        alls = torch.cat([fr[s] for fr in frames])
        mu_init.append(alls[torch.randperm(len(alls), generator=g)[:Cs[s]]].clone())
        out[f"in_mu{s}"] = mu_init[s].numpy()
        out[f"in_sigma{s}"] = np.array(2.0 * sigs[s] + 0.05)
        for k in range(K):
            out[f"in_x{k}_{s}"] = frames[k][s].numpy()
    for prec, dt in [("ref32", torch.float32), ("gold", torch.float64)]:
        sp = spec_of(dt)
        GM = [rg.GaussianMixtureUnif(mu_init[s].to(dt), sigma=2.0 * sigs[s] + 0.05, spec=sp, computversion="torch") for s in range(3)]
        FE_TRACE.clear()
        PSR, evol = ra.ICP_atlas([[x.to(dt) for x in fr] for fr in frames],
                                 GMM_parameters={"init_components": GM, "optimize_weights": True},
                                 registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.3},
                                 numerical_options={"computversion": "torch", "compspec": sp, "dataspec": sp,
                                                    "support_LDDMM": {"scheme": "decim", "rho": rho}},
                                 optim_options={"max_iterations": 3, "max_repeat_GMM": 10, "convergence_tolerance": 1e-3},
                                 printstuff=False)
        out[f"{prec}_FE"] = np.array(float(PSR.FE))
        out[f"{prec}_Cfe"] = np.array([float(c) for c in PSR.Cfe])
        out[f"{prec}_FE_trace"] = np.array(FE_TRACE)
        for it in range(len(evol["a0"])):
            for k in range(K):
                out[f"{prec}_a0_it{it}_{k}"] = evol["a0"][it][k].numpy()
        out[f"{prec}_quadloss"] = np.asarray(PSR.quadloss, dtype=np.float64)
        out[f"{prec}_regloss"] = np.array([float(r) for r in PSR.regloss])
        for s in range(3):
            out[f"{prec}_sigma{s}"] = np.array(float(PSR.GMMi[s].sigma))
            out[f"{prec}_mu{s}"] = PSR.GMMi[s].mu.numpy()
            out[f"{prec}_w{s}"] = PSR.GMMi[s].w.numpy()
            for k in range(K):
                out[f"{prec}_x1_{k}_{s}"] = PSR.x1[k, s].numpy()
        for k in range(K):
            out[f"{prec}_q0_{k}"] = PSR.q0[k].numpy()
    np.savez_compressed(os.path.join(OUT, fname), **out)
    print(fname, len(out), [out[f"gold_q0_{k}"].shape[0] for k in range(K)])


def gen_atlas_mid():
    """The same three-structure atlas with 4 larger frames and a finer decimated support (spacing 0.5 sigma_LDDMM): more than 64
    support points per frame, i.e. the mid-size stage kernels and the frame groups of the lock-step registration end to end."""
    gen_atlas_s3(K=4, Ns=150, seed=713, rho=0.5, fname="atlas_mid.npz")


class _keops_ordering:
    """Run the reference's own outer loops with the KeOps ORDERING of the M step (core/GMM.py:432-458, 462-496: sigma'
    from distances to the NEW mu, Gaussian normalisation from the NEW sigma).  pykeops is absent, so EM_step_keops cannot
    execute; the class attribute EM_step_torch -- the function every GMM object binds here (core/GMM.py:126-144) -- is
    swapped for the oracle's restatement of that ordering (oracle/gmm.py, variant="keops"), acting on the reference
    object's own state.  Everything else (EM_optimization, PSR loops, L-BFGS, LDDMM) stays the reference's code.
    What this pins: the product's DEFAULT path (computversion="keops") through the reference's loops."""

    def __enter__(self):
        import diffICP.core.GMM as rg
        sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
        from oracle.gmm import GMMOracle
        self.rg, self.old = rg, rg.GaussianMixtureUnif.EM_step_torch

        def em_step(gmm, X, skip_M=False):
            O = GMMOracle(gmm.mu, gmm.sigma, w=gmm.w, outliers=gmm.outliers, to_optimize=gmm.to_optimize,
                          ensure_continuum=gmm.ensure_continuum)
            Y, Cfe, FE = O.em_step(X.detach(), skip_M=skip_M, variant="keops")
            gmm.mu, gmm.w, gmm.sigma = O.mu, O.w, O.sigma
            if gmm.outliers is not None:
                gmm.outliers.update(O.outliers)
            return Y, Cfe, FE
        rg.GaussianMixtureUnif.EM_step_torch = em_step

    def __exit__(self, *a):
        self.rg.GaussianMixtureUnif.EM_step_torch = self.old


def gen_keops_ordering():
    """The end-to-end fixtures again (two-set dense, 2-D atlas of gen_psr, S = 3 atlas) under the KeOps M-step ordering."""
    import contextlib
    import diffICP.core.GMM as rg
    import diffICP.api.ICP_atlas as ra
    import diffICP.api.ICP_two_set as rt
    _patch_coverage()
    out = {}
    two = np.load(os.path.join(OUT, "two_set.npz"))
    s3 = np.load(os.path.join(OUT, "atlas_s3.npz"))
    psr = np.load(os.path.join(OUT, "psr.npz"))
    with _keops_ordering():
        for prec, dt in [("ref32", torch.float32), ("gold", torch.float64)]:
            sp = spec_of(dt)
            # ---- two-set, dense support, API-default logdet model --------------------------------------------------
            ctx = _fp64_defaults() if dt == torch.float64 else contextlib.nullcontext()
            with ctx:
                FE_TRACE.clear()
                PSR, evol = rt.ICP_two_set(torch.from_numpy(two["in_xA"]).to(dt), torch.from_numpy(two["in_xB"]).to(dt),
                                           {"sigma": 0.1, "optimize_sigma": True, "outlier_weight": None},
                                           {"type": "diffeomorphic", "lambda_LDDMM": 500.0, "sigma_LDDMM": 0.2},
                                           numerical_options={"support_LDDMM": {"scheme": "dense"}, "computversion": "torch"},
                                           optim_options={"max_iterations": 3, "convergence_tolerance": 1e-3},
                                           plotstuff=False, printstuff=False)
            out[f"two_{prec}_FE_trace"] = np.array(FE_TRACE)
            out[f"two_{prec}_sigma"] = np.array(float(PSR.GMMi[0].sigma))
            out[f"two_{prec}_x1"] = PSR.x1[0, 0].numpy()
            out[f"two_{prec}_a0"] = PSR.a0[0].numpy()
            # ---- 2-D atlas of gen_psr (3 frames, C = 6, hybrid, Euler, grid) ------------------------------------------
            sets = [torch.from_numpy(psr[f"atlas_in_x{k}"]) for k in range(3)]
            G = rg.GaussianMixtureUnif(torch.from_numpy(psr["atlas_in_mu"]).to(dt), sigma=0.25 * float(torch.cat(sets).std()),
                                       spec=sp, computversion="torch")
            FE_TRACE.clear()
            PSR, evol = ra.ICP_atlas([[x.to(dt)] for x in sets],
                                     GMM_parameters={"init_components": [G], "optimize_weights": True},
                                     registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.2},
                                     numerical_options={"computversion": "torch", "compspec": sp, "dataspec": sp,
                                                        "support_LDDMM": {"scheme": "grid", "rho": 1.0}},
                                     optim_options={"max_iterations": 3, "max_repeat_GMM": 10, "convergence_tolerance": 1e-3},
                                     printstuff=False)
            out[f"atlas_{prec}_FE_trace"] = np.array(FE_TRACE)
            out[f"atlas_{prec}_sigma"] = np.array(float(PSR.GMMi[0].sigma))
            out[f"atlas_{prec}_mu"], out[f"atlas_{prec}_w"] = PSR.GMMi[0].mu.numpy(), PSR.GMMi[0].w.numpy()
            for k in range(3):
                out[f"atlas_{prec}_x1_{k}"] = PSR.x1[k, 0].numpy()
            # ---- S = 3 atlas (3-D, decimated support) ----------------------------------------------------------------------
            frames = [[torch.from_numpy(s3[f"in_x{k}_{s}"]).to(dt) for s in range(3)] for k in range(3)]
            GM = [rg.GaussianMixtureUnif(torch.from_numpy(s3[f"in_mu{s}"]).to(dt), sigma=float(s3[f"in_sigma{s}"]), spec=sp,
                                         computversion="torch") for s in range(3)]
            FE_TRACE.clear()
            PSR, evol = ra.ICP_atlas(frames, GMM_parameters={"init_components": GM, "optimize_weights": True},
                                     registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.3},
                                     numerical_options={"computversion": "torch", "compspec": sp, "dataspec": sp,
                                                        "support_LDDMM": {"scheme": "decim", "rho": 1.0}},
                                     optim_options={"max_iterations": 3, "max_repeat_GMM": 10, "convergence_tolerance": 1e-3},
                                     printstuff=False)
            out[f"s3_{prec}_FE_trace"] = np.array(FE_TRACE)
            for s in range(3):
                out[f"s3_{prec}_sigma{s}"] = np.array(float(PSR.GMMi[s].sigma))
                out[f"s3_{prec}_mu{s}"] = PSR.GMMi[s].mu.numpy()
                for k in range(3):
                    out[f"s3_{prec}_x1_{k}_{s}"] = PSR.x1[k, s].numpy()
    np.savez_compressed(os.path.join(OUT, "keops_order.npz"), **out)
    print("keops_order.npz", len(out))


def gen_keops_ordering_spread():
    """How far the reference's OWN fp32 run of the 2-D atlas moves when its inputs change by one ulp: the registration is
    one unconverged L-BFGS step per outer iteration, and line-search branches amplify rounding.  Three runs with every
    coordinate multiplied by 1 + {-1, 0, 1} * 6e-8; the largest deviations from the fp64 run are appended to
    keops_order.npz (atlas_ulp_spread_*) and bound the tolerances of the corresponding tests."""
    import diffICP.core.GMM as rg
    import diffICP.api.ICP_atlas as ra
    _patch_coverage()
    path = os.path.join(OUT, "keops_order.npz")
    out = dict(np.load(path))
    psr = np.load(os.path.join(OUT, "psr.npz"))
    sp = spec_of(torch.float32)
    dx, dfe, dmu, dsig = 0.0, 0.0, 0.0, 0.0
    with _keops_ordering():
        for trial in range(1, 4):
            g = torch.Generator().manual_seed(trial)
            sets = [torch.from_numpy(psr[f"atlas_in_x{k}"]) for k in range(3)]
            sets = [s * (1 + torch.randint(-1, 2, s.shape, generator=g).float() * 6e-8) for s in sets]
            G = rg.GaussianMixtureUnif(torch.from_numpy(psr["atlas_in_mu"]), sigma=0.25 * float(torch.cat(sets).std()),
                                       spec=sp, computversion="torch")
            PSR, _ = ra.ICP_atlas([[x] for x in sets], GMM_parameters={"init_components": [G], "optimize_weights": True},
                                  registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.2},
                                  numerical_options={"computversion": "torch", "compspec": sp, "dataspec": sp,
                                                     "support_LDDMM": {"scheme": "grid", "rho": 1.0}},
                                  optim_options={"max_iterations": 3, "max_repeat_GMM": 10, "convergence_tolerance": 1e-3},
                                  printstuff=False)
            dx = max(dx, max(np.abs(PSR.x1[k, 0].numpy() - out[f"atlas_gold_x1_{k}"]).max() for k in range(3)))
            dfe = max(dfe, abs(float(PSR.FE) - out["atlas_gold_FE_trace"][-1]))
            dmu = max(dmu, np.abs(PSR.GMMi[0].mu.numpy() - out["atlas_gold_mu"]).max())
            dsig = max(dsig, abs(float(PSR.GMMi[0].sigma) - float(out["atlas_gold_sigma"])))
    out["atlas_ulp_spread_x1"], out["atlas_ulp_spread_FE"] = np.array(dx), np.array(dfe)
    out["atlas_ulp_spread_mu"], out["atlas_ulp_spread_sigma"] = np.array(dmu), np.array(dsig)
    np.savez_compressed(path, **out)
    print("keops_order.npz + ulp spread:", dx, dfe, dmu, dsig)


def reference_function(relpath, name, namespace):
    """Compile ONE function of a reference module from its source text (for modules that cannot be imported here because
    they import pykeops at the top) and return it; nothing of the source is written anywhere."""
    import ast
    path = os.path.join(REF_ROOT, relpath)
    tree = ast.parse(open(path).read(), filename=path)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = dict(namespace)
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def gen_pointsets(rk):
    """decimate and point_set_distance of the reference itself (tools/point_sets.py:102-133, :46-95)."""
    import math
    ref_decimate = reference_function("diffICP/tools/point_sets.py", "decimate", {"np": np, "torch": torch})

    def kmin2_scale(x):        # the one KeOps call (Kmin(2)) that cannot run here: dense equivalent
        d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
        return float(d2.topk(2, dim=1, largest=False).values[:, 1].mean().sqrt())
    ref_psd = reference_function("diffICP/tools/point_sets.py", "point_set_distance",
                                 {"math": math, "warnings": warnings, "torch": torch, "intrinsic_scale": kmin2_scale,
                                  "GaussKernel": lambda s, D: rk.GaussKernel(s, D, computversion="torch")})
    out, cases = {}, []
    g = torch.Generator().manual_seed(99)
    sets = {
        "u2": (torch.rand(700, 2, generator=g), 0.08),
        "u3": (torch.rand(900, 3, generator=g), 0.17),
        "clustered2": (torch.cat([0.03 * torch.randn(250, 2, generator=g) + c for c in torch.rand(4, 2, generator=g)]), 0.05),
        # exact ties in the neighbour counts and distances exactly equal to R^2 (0.25^2 = 0.0625 is exact in fp32)
        "lattice2": (torch.stack(torch.meshgrid(torch.arange(12.) * 0.25, torch.arange(9.) * 0.25, indexing="ij"), -1).reshape(-1, 2), 0.25),
        "lattice3": (torch.stack(torch.meshgrid(*[torch.arange(6.) * 0.5] * 3, indexing="ij"), -1).reshape(-1, 3), 0.5),
    }
    for tag, (x, R) in sets.items():
        kept, rej = ref_decimate(x, R)
        out[f"{tag}_x"], out[f"{tag}_R"] = x.numpy(), np.array(R)
        out[f"{tag}_kept"], out[f"{tag}_rejected"] = np.array(kept, dtype=np.int64), np.array(rej, dtype=np.int64)
        cases.append(tag)
    X, Y = sets["u3"][0][:500], sets["u3"][0][400:] + 0.05
    for tag, kw in (("auto", {}), ("fixed", {"sigma_X": 0.2, "sigma_Y": 0.15})):
        out[f"psd_{tag}_ref32"] = np.array(float(ref_psd(X, Y, **kw)))
        out[f"psd_{tag}_gold"] = np.array(float(ref_psd(X.double(), Y.double(), **kw)))
    out["psd_X"], out["psd_Y"] = X.numpy(), Y.numpy()
    # data_distance of the comparator algorithm (core/PSR_standard.py:37-58), with and without template weights
    ref_dd = reference_function("diffICP/core/PSR_standard.py", "data_distance", {"GenKernel": rk.GenKernel})
    wts = torch.rand(Y.shape[0], generator=g)
    wts = wts / wts.sum()
    out["dd_w"], out["dd_sigma"] = wts.numpy(), np.array(0.15)
    for prec, dt in (("ref32", torch.float32), ("gold", torch.float64)):
        Kd = rk.GaussKernel(0.15, 3, computversion="torch")
        out[f"dd_plain_{prec}"] = np.array(float(ref_dd(Kd, X.to(dt), Y.to(dt))))
        out[f"dd_weighted_{prec}"] = np.array(float(ref_dd(Kd, X.to(dt), Y.to(dt), wts.to(dt))))
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "pointsets.npz"), **out)
    print("pointsets.npz", len(out))


if __name__ == "__main__":
    torch.set_num_threads(8)
    rk, rl, rg = load_reference()
    if len(sys.argv) > 1:          # regenerate selected fixtures only
        for what in sys.argv[1:]:
            {"pointsets": lambda: gen_pointsets(rk), "v2p": lambda: gen_v2p(rl), "two_set": gen_two_set,
             "two_set_fine": gen_two_set_fine, "atlas_s3": gen_atlas_s3, "atlas_mid": gen_atlas_mid, "kernels": lambda: gen_kernels(rk), "lddmm": lambda: gen_lddmm(rl),
             "gmm": lambda: gen_gmm(rg), "psr": gen_psr, "keops_order": gen_keops_ordering, "keops_order_spread": gen_keops_ordering_spread}[what]()
        sys.exit(0)
    gen_kernels(rk)
    gen_lddmm(rl)
    gen_gmm(rg)
    gen_psr()
    gen_pointsets(rk)
    gen_v2p(rl)
    gen_two_set()
    gen_two_set_fine()
    gen_atlas_s3()
    gen_atlas_mid()
    gen_keops_ordering()
    gen_keops_ordering_spread()
