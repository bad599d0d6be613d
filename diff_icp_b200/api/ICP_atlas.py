"""Groupwise atlas building by diffeomorphic ICP: drop-in for the reference's ``diffICP/api/ICP_atlas.py:51-305``.

``ICP_atlas(x0, GMM_parameters, registration_parameters, numerical_options, optim_options, callback_function, printstuff)``
fits one GMM per structure to the warped points of all frames while registering every frame to it.  Option names,
defaults, the initialisation variants of ``GMM_parameters["init_components"]`` and the returned ``(PSR, evol)`` follow the
reference.

Multi-GPU: pass ``numerical_options["comm"]`` (a ``diff_icp_b200.dist.StatsComm``) and give each rank ITS frames
(see ``diff_icp_b200.dist.shard_frames``); every rank returns its local PSR, the GMMs are identical on all ranks.

Outside the B200 hot path, hence not provided: affine registration types, ``lambda_LDDMM="auto"``, plotting.
"""

import copy

import torch

from ..core.GMM import GaussianMixtureUnif
from ..core.LDDMM import LDDMMModel
from ..core.PSR import DiffPSR
from ..tools.in_out import read_point_sets
from ..tools.spec import defspec

_REG_TYPES = ["rigid", "similarity", "general_affine", "diffeomorphic"]


def _defaults(opts, **kw):
    out = dict(opts)
    for key, value in kw.items():
        if out.get(key) is None:
            out[key] = value
    return out


def ICP_atlas(x0, GMM_parameters={}, registration_parameters={}, numerical_options={}, optim_options={},
              callback_function=None, printstuff=True):
    init = GMM_parameters.get("init_components")
    assert type(init) is int or \
        type(init) is tuple and init[0] == "set" or \
        type(init) is dict and set(init.keys()) == {"set", "C"} or \
        type(init) is list and all(isinstance(g, GaussianMixtureUnif) for g in init), \
        "Wrong format for parameter GMM_parameters['init_components']. See docstring for ICP_atlas."
    ow = GMM_parameters.get("outlier_weight")
    assert ow is None or ow == "optimize" or isinstance(ow, (int, float)), \
        "incorrect value for GMM_parameters['outlier_weight'].  See docstring for ICP_atlas."
    assert GMM_parameters.get("fixed_sigma") is None or GMM_parameters["fixed_sigma"] > 0, \
        "GMM_parameters['fixed_sigma'] should be absent (normal setting), or a strictly positive number"
    assert registration_parameters.get("type") in _REG_TYPES, f"registration_parameters['type'] should be one of: {_REG_TYPES}"
    if registration_parameters["type"] != "diffeomorphic":
        raise NotImplementedError("diff_icp_b200 implements the diffeomorphic (LDDMM) path; affine registrations are "
                                  "closed-form D x D algebra outside the B200 hot path (SURVEY.md §2)")
    assert {"lambda_LDDMM", "sigma_LDDMM"}.issubset(registration_parameters.keys()), \
        "if type=diffeomorphic, registration_parameters should define values of lambda_LDDMM and sigma_LDDMM"

    numerical_options = _defaults(numerical_options,
                                  support_LDDMM={"scheme": "grid", "rho": 1.0},
                                  computversion="keops",
                                  gradcomponent_LDDMM=False,
                                  integration_scheme_LDDMM="Euler",
                                  integration_nt_LDDMM=10,
                                  compspec=defspec,
                                  dataspec=defspec)
    optim_options = _defaults(optim_options, max_iterations=25, convergence_tolerance=1e-3, max_repeat_GMM=10)
    compspec, dataspec = numerical_options["compspec"], numerical_options["dataspec"]
    comm = numerical_options.get("comm")

    x0, K, S, D = read_point_sets(x0)

    use_outliers = ow is not None
    opt_sigma = GMM_parameters.get("fixed_sigma") is None
    opt_weights = GMM_parameters.get("optimize_weights")
    opt_weights = True if opt_weights is None else opt_weights
    ensure_continuum = bool(GMM_parameters.get("ensure_continuum"))
    reinit_mu, reinit_sigma = False, False

    if type(init) is int:
        GMMi = [GaussianMixtureUnif(torch.zeros(init, D), use_outliers=use_outliers, spec=compspec) for _ in range(S)]
        reinit_mu, reinit_sigma = True, opt_sigma
    elif type(init) is tuple:
        GMMi = [GaussianMixtureUnif(x0[init[1]][s], use_outliers=use_outliers, spec=compspec) for s in range(S)]
        reinit_sigma = opt_sigma
    elif type(init) is dict:
        GMMi = [GaussianMixtureUnif.get_GMM_model(x0[init["set"]][s].to(**compspec), init["C"], fixed_sigma=None,
                                                  optimize_w=False, use_outliers=use_outliers, spec=compspec)
                for s in range(S)]
    else:
        GMMi = [copy.deepcopy(g) for g in init]

    for GMM in GMMi:
        if isinstance(ow, (int, float)):
            GMM.outliers["eta0"] = ow
        GMM.to_optimize = {"mu": True, "sigma": opt_sigma, "w": opt_weights, "eta0": ow == "optimize"}
        GMM.ensure_continuum = ensure_continuum
        if not opt_sigma:
            GMM.sigma = GMM_parameters["fixed_sigma"]

    lam = registration_parameters["lambda_LDDMM"]
    if lam == "auto":
        raise NotImplementedError("lambda_LDDMM='auto' (core/calibration.py, self-described as unstable) is not part of the B200 hot path")
    LMi = LDDMMModel(sigma=registration_parameters["sigma_LDDMM"], D=D, lambd=lam, withlogdet=True,
                     gradcomponent=numerical_options["gradcomponent_LDDMM"],
                     computversion=numerical_options["computversion"],
                     scheme=numerical_options["integration_scheme_LDDMM"],
                     spec=compspec, nt=numerical_options["integration_nt_LDDMM"])
    PSR = DiffPSR(x0, GMMi, LMi, compspec=compspec, dataspec=dataspec, comm=comm)
    if numerical_options["support_LDDMM"]["scheme"] != "dense":
        PSR.set_support_scheme(**numerical_options["support_LDDMM"])
    evol = {"a0": [], "GMMi": []}

    PSR.reinitialize_GMM(do_mu=reinit_mu, do_sigma=reinit_sigma)
    PSR.printstuff = printstuff

    tol = optim_options["convergence_tolerance"]
    last_FE = None
    for it in range(optim_options["max_iterations"]):
        if printstuff:
            print("ITERATION NUMBER ", it)
        evol["GMMi"].append(copy.deepcopy(PSR.GMMi[0]))
        evol["a0"].append([a.clone().detach().cpu() for a in PSR.a0])
        if it != 0 or reinit_mu:               # otherwise start by optimising the registrations (api/ICP_atlas.py:281)
            PSR.GMM_opt(max_iterations=optim_options["max_repeat_GMM"], tol=tol)
        if callback_function is not None:
            callback_function(PSR, True)
        PSR.Reg_opt(tol=tol, nmax=1)
        if callback_function is not None:
            callback_function(PSR, False)
        if it > 1 and abs(PSR.FE - last_FE) < tol * abs(last_FE):
            if printstuff:
                print("Difference in Free Energy is below tolerance threshold : optimization is over.")
            break
        last_FE = PSR.FE
    if it + 1 == optim_options["max_iterations"] and printstuff:
        print("Reached maximum number of iterations (before reaching convergence threshold).")
    return PSR, evol
