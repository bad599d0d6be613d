"""Two-point-set diffeomorphic ICP matching: drop-in for the reference's ``diffICP/api/ICP_two_set.py:73-288``.

``ICP_two_set(xA, xB, GMM_parameters, registration_parameters, numerical_options, optim_options, plotstuff, printstuff)``
registers the data set xA onto a GMM whose centroids are the fixed template xB (mu, w frozen; sigma optional), alternating
``GMM_opt`` and ``Reg_opt(nmax=1)`` until the free energy stalls.  Option names, defaults and the returned ``(PSR, evol)``
are the reference's.  Faithfully preserved quirk: ``gradcomponent_LDDMM`` gets a default but is
NOT forwarded to ``LDDMMModel`` (api/ICP_two_set.py:151 vs :203-207), so the two-set path runs the full logdet model.

Outside the B200 hot path, hence not provided: affine registration types, ``lambda_LDDMM="auto"`` calibration, plotting.
"""

import copy

import torch

from ..core.GMM import GaussianMixtureUnif
from ..core.LDDMM import LDDMMModel
from ..core.PSR import DiffPSR

_REG_TYPES = ["rigid", "similarity", "general_affine", "diffeomorphic"]


def _defaults(opts, **kw):
    out = dict(opts)
    for key, value in kw.items():
        if out.get(key) is None:
            out[key] = value
    return out


def ICP_two_set(xA, xB, GMM_parameters: dict, registration_parameters: dict,
                numerical_options={}, optim_options={}, plotstuff=True, printstuff=True):
    assert registration_parameters["type"] in _REG_TYPES, f"registration_parameters['type'] should be one of: {_REG_TYPES}"
    if registration_parameters["type"] != "diffeomorphic":
        raise NotImplementedError("diff_icp_b200 implements the diffeomorphic (LDDMM) path; affine registrations are "
                                  "closed-form D x D algebra outside the B200 hot path (SURVEY.md §2)")
    assert {"lambda_LDDMM", "sigma_LDDMM"}.issubset(registration_parameters.keys()), \
        "if type=diffeomorphic, registration_parameters should define values of lambda_LDDMM and sigma_LDDMM"

    xB_is_gmm = isinstance(xB, GaussianMixtureUnif)
    if xB_is_gmm:
        assert GMM_parameters is None, \
            "when using the 'xB=GMM' hack, set GMM_parameters=None (you can directly modify xB's GMM parameters if required)"
    else:
        assert {"optimize_sigma", "sigma"}.issubset(GMM_parameters.keys()), \
            "GMM_parameters should at least define values of sigma (float>0) and optimize_sigma (True/False)"
        ow = GMM_parameters.get("outlier_weight")
        assert ow is None or ow == "optimize" or isinstance(ow, (int, float)), "incorrect value for GMM_parameters['outlier_weight']"

    numerical_options = _defaults(numerical_options,
                                  support_LDDMM={"scheme": "grid", "rho": 1.0},
                                  computversion="keops",
                                  gradcomponent_LDDMM=False,           # defaulted but never forwarded, as in the reference
                                  integration_scheme_LDDMM="Euler",
                                  integration_nt_LDDMM=10)
    optim_options = _defaults(optim_options, max_iterations=25, convergence_tolerance=1e-3, max_repeat_GMM=10)

    if xB_is_gmm:
        GMMi = copy.deepcopy(xB)
        xB = GMMi.mu
    assert xA.shape[1] == xB.shape[1], "point sets xA and xB should have same vector dimension (dim 1)"
    D = xA.shape[1]

    if not xB_is_gmm:
        ow = GMM_parameters.get("outlier_weight")
        GMMi = GaussianMixtureUnif(xB, use_outliers=ow is not None, sigma=GMM_parameters["sigma"])
        if isinstance(ow, (int, float)):
            GMMi.outliers["eta0"] = ow
        GMMi.to_optimize = {"mu": False, "sigma": GMM_parameters["optimize_sigma"], "w": False, "eta0": ow == "optimize"}

    lam = registration_parameters["lambda_LDDMM"]
    if lam == "auto":
        raise NotImplementedError("lambda_LDDMM='auto' (core/calibration.py, self-described as unstable) is not part of the B200 hot path")
    LMi = LDDMMModel(sigma=registration_parameters["sigma_LDDMM"], D=D, lambd=lam, withlogdet=True,
                     computversion=numerical_options["computversion"],
                     scheme=numerical_options["integration_scheme_LDDMM"],
                     nt=numerical_options["integration_nt_LDDMM"])
    PSR = DiffPSR(xA, GMMi, LMi)
    if numerical_options["support_LDDMM"]["scheme"] != "dense":
        PSR.set_support_scheme(**numerical_options["support_LDDMM"])
    PSR.printstuff = printstuff
    evol = {"a0": [], "GMMi": []}

    tol = optim_options["convergence_tolerance"]
    last_FE = None
    for it in range(optim_options["max_iterations"]):
        if printstuff:
            print("ITERATION NUMBER ", it)
        evol["GMMi"].append(copy.deepcopy(PSR.GMMi[0]))
        evol["a0"].append([a.clone().detach().cpu() for a in PSR.a0])
        PSR.GMM_opt(max_iterations=optim_options["max_repeat_GMM"], tol=tol)
        PSR.Reg_opt(tol=tol, nmax=1)
        if it > 1 and abs(PSR.FE - last_FE) < tol * abs(last_FE):
            if printstuff:
                print("Difference in Free Energy is below tolerance threshold : optimization is over.")
            break
        last_FE = PSR.FE
    if printstuff and it + 1 == optim_options["max_iterations"]:
        print("Reached maximum number of iterations (before reaching convergence threshold).")
    return PSR, evol
