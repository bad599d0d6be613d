"""Shared bodies of the end-to-end API parity checks (CPU emulation and GPU) against runs of the UNMODIFIED reference
(tests/golden/two_set.npz, atlas_s3.npz: torch-twin ordering of the M step) and of the reference's loops with the
restated KeOps ordering (tests/golden/keops_order.npz; see make_golden._keops_ordering).

Tolerances.  One L-BFGS step of an unconverged registration amplifies rounding differences (the reference's own fp32 and
fp64 runs differ by 1e-4 in the free energy and 1.6e-4 in the warped points on the two-set case, 2e-3 on one frame of the
S = 3 atlas), so the bar is "as close to the fp64 run as the reference's own fp32 run, times 3, plus a floor":
free energy at every update: 3 |ref32 - gold| + 5e-4 max|trace|; sigma: 3 |ref32 - gold| + 2e-3 relative;
warped points: 3 max|ref32 - gold| + 1e-2 sigma_LDDMM."""
import numpy as np
import torch


class FETrace:
    """Records the free energy after every MultiPSR.update_FE (the same hook the fixture generator puts on the reference)."""

    def __init__(self, monkeypatch):
        from diff_icp_b200.core import PSR as psr_mod
        self.values, self.all_frames = [], []
        orig = psr_mod.MultiPSR.update_FE
        rec, allf = self.values, self.all_frames

        def update_FE(this, message=None):
            orig(this, message=message)
            rec.append(float(this.FE))
            # the lock-step registration updates the free energy ONCE for all K frames where the reference (and the
            # frame-by-frame path) update it after every frame (core/PSR.py:569)
            allf.append(this.K if (message or "").startswith("Registration of all frames") else 0)
        monkeypatch.setattr(psr_mod.MultiPSR, "update_FE", update_FE)


def _close_trace(tr, gold, ref, floor=5e-4):
    """Every free-energy value of our run against the reference's trace; an all-frames update is compared with the
    reference's value after its LAST frame of that Reg_opt."""
    gold, ref = np.asarray(gold), np.asarray(ref)
    j, g, r = 0, [], []
    for k_all in tr.all_frames:
        j += k_all if k_all else 1
        g.append(gold[j - 1])
        r.append(ref[j - 1])
    assert j == len(gold), (tr.values, gold)
    ours, g, r = np.asarray(tr.values), np.asarray(g), np.asarray(r)
    tol = 3 * np.abs(r - g) + floor * np.abs(gold).max()
    assert (np.abs(ours - g) <= tol).all(), (ours, g, tol)


def _close_points(a, gold, ref, sig_lddmm, floor=1e-2):
    assert np.abs(a - gold).max() <= 3 * np.abs(ref - gold).max() + floor * sig_lddmm, (np.abs(a - gold).max(), np.abs(ref - gold).max())


def _close_sigma(a, gold, ref):
    assert abs(a - gold) <= 3 * abs(ref - gold) + 2e-3 * gold, (a, gold, ref)


def run_two_set(golden, to_dev, monkeypatch, case, ordering):
    """api.ICP_two_set (api/ICP_two_set.py:73-288), API-default full logdet model.  ordering="torch": fixture of the
    unmodified reference; the API gives no handle on the GMM's computversion, so the torch-twin ordering
    is selected through the reference's own 'xB = GMM' hack (:131-136).  ordering="keops": the product's default path."""
    from diff_icp_b200.api.ICP_two_set import ICP_two_set
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    g = golden("two_set_fine" if case == "decimfine" else "two_set")
    xA, xB = to_dev(g["in_xA"]), to_dev(g["in_xB"])
    support = {"dense": {"scheme": "dense"}, "decim": {"scheme": "decim", "rho": 1.0},
               "decimfine": {"scheme": "decim", "rho": 0.7}}[case]
    tr = FETrace(monkeypatch)
    if ordering == "torch":
        G = GaussianMixtureUnif(xB, sigma=0.1, computversion="torch")          # default spec, like the API itself
        G.to_optimize = {"mu": False, "sigma": True, "w": False, "eta0": False}
        B, gmm_par = G, None
        pre, key = g, case
    else:
        B, gmm_par = xB, {"sigma": 0.1, "optimize_sigma": True, "outlier_weight": None}
        pre, key = golden("keops_order"), "two"
    PSR, evol = ICP_two_set(xA, B, gmm_par, {"type": "diffeomorphic", "lambda_LDDMM": 500.0, "sigma_LDDMM": 0.2},
                            numerical_options={"support_LDDMM": support},
                            optim_options={"max_iterations": 3, "convergence_tolerance": 1e-3}, plotstuff=False, printstuff=False)
    assert PSR.LMi.gradcomponent and PSR.LMi.eta == 1 / 500.0            # quirk preserved: full logdet model
    _close_trace(tr, pre[f"{key}_gold_FE_trace"], pre[f"{key}_ref32_FE_trace"])
    _close_sigma(PSR.GMMi[0].sigma, float(pre[f"{key}_gold_sigma"]), float(pre[f"{key}_ref32_sigma"]))
    _close_points(PSR.x1[0, 0].cpu().numpy(), pre[f"{key}_gold_x1"], pre[f"{key}_ref32_x1"], 0.2)
    if ordering == "torch":
        # a0 before the first iteration = what initialize_a0 / update_a0 (v2p with eta != 0) produced
        a0i = g[f"{case}_gold_a0_init"]
        assert np.abs(evol["a0"][0][0].numpy() - a0i).max() < 2e-4 * np.abs(a0i).max()
        assert np.abs(PSR.q0[0].cpu().numpy() - g[f"{case}_gold_q0"]).max() < 1e-6
        assert len(evol["a0"]) == 3 and len(evol["GMMi"]) == 3
    return PSR


def run_atlas_s3(golden, to_dev, monkeypatch, spec, ordering, name="atlas_s3", K=3, rho=1.0, fe_floor=5e-4, pt_floor=1e-2):
    """api.ICP_atlas with S = 3 structures: per-structure GMM loop (core/PSR.py:242-271), per-structure sigma in the data
    loss (:498-516); 3 ragged frames, 3-D, decimated support, hybrid model."""
    from diff_icp_b200.api.ICP_atlas import ICP_atlas
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    g = golden(name)
    frames = [[to_dev(g[f"in_x{k}_{s}"]) for s in range(3)] for k in range(K)]
    GM = [GaussianMixtureUnif(to_dev(g[f"in_mu{s}"]), sigma=float(g[f"in_sigma{s}"]), spec=spec, computversion=ordering) for s in range(3)]
    tr = FETrace(monkeypatch)
    PSR, evol = ICP_atlas(frames, GMM_parameters={"init_components": GM, "optimize_weights": True},
                          registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.3},
                          numerical_options={"computversion": ordering, "compspec": spec, "dataspec": spec,
                                             "support_LDDMM": {"scheme": "decim", "rho": rho}},
                          optim_options={"max_iterations": 3, "max_repeat_GMM": 10, "convergence_tolerance": 1e-3},
                          printstuff=False)
    assert PSR.S == 3 and PSR.K == K
    if ordering == "torch":
        pre, k_ = g, ""
        for k in range(K):
            assert np.abs(PSR.q0[k].cpu().numpy() - g[f"gold_q0_{k}"]).max() < 1e-6          # decimated support: same points
    else:
        pre, k_ = golden("keops_order"), "s3_"
    _close_trace(tr, pre[f"{k_}gold_FE_trace"], pre[f"{k_}ref32_FE_trace"], fe_floor)
    for s in range(3):
        _close_sigma(PSR.GMMi[s].sigma, float(pre[f"{k_}gold_sigma{s}"]), float(pre[f"{k_}ref32_sigma{s}"]))
        _close_points(PSR.GMMi[s].mu.cpu().numpy(), pre[f"{k_}gold_mu{s}"], pre[f"{k_}ref32_mu{s}"], 0.3, pt_floor)
        for k in range(K):
            _close_points(PSR.x1[k, s].cpu().numpy(), pre[f"{k_}gold_x1_{k}_{s}"], pre[f"{k_}ref32_x1_{k}_{s}"], 0.3, pt_floor)
    return PSR


def run_atlas_2d(golden, to_dev, monkeypatch, spec, ordering):
    """The 2-D atlas of tests/golden/psr.npz (3 frames, C = 6, hybrid, Euler, grid support) under either M-step ordering."""
    from diff_icp_b200.api.ICP_atlas import ICP_atlas
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    g = golden("psr")
    sets = [to_dev(g[f"atlas_in_x{k}"]) for k in range(3)]
    G = GaussianMixtureUnif(to_dev(g["atlas_in_mu"]), sigma=0.25 * float(torch.cat(sets).std()), spec=spec, computversion=ordering)
    tr = FETrace(monkeypatch)
    PSR, evol = ICP_atlas(sets, GMM_parameters={"init_components": [G], "optimize_weights": True},
                          registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.2},
                          numerical_options={"computversion": ordering, "compspec": spec, "dataspec": spec,
                                             "support_LDDMM": {"scheme": "grid", "rho": 1.0}},
                          optim_options={"max_iterations": 3, "max_repeat_GMM": 10, "convergence_tolerance": 1e-3},
                          printstuff=False)
    if ordering == "keops":
        pre = golden("keops_order")
        _close_trace(tr, pre["atlas_gold_FE_trace"], pre["atlas_ref32_FE_trace"])
        _close_sigma(PSR.GMMi[0].sigma, float(pre["atlas_gold_sigma"]), float(pre["atlas_ref32_sigma"]))
        _close_points(PSR.GMMi[0].mu.cpu().numpy(), pre["atlas_gold_mu"], pre["atlas_ref32_mu"], 0.2)
        # warped points: this run is three unconverged L-BFGS steps whose line-search branches amplify rounding -- the
        # reference's OWN fp32 run moves by atlas_ulp_spread_x1 (1.5e-2, measured by make_golden.gen_keops_ordering_spread)
        # when its inputs change by one ulp; the bar is that spread, not the luck of one fp32 run
        spread = float(pre["atlas_ulp_spread_x1"])
        for k in range(3):
            err = np.abs(PSR.x1[k, 0].cpu().numpy() - pre[f"atlas_gold_x1_{k}"]).max()
            assert err <= 1.5 * spread, (k, err, spread)
    return PSR
