"""GPU: the API entry points end to end against runs of the reference itself -- api.ICP_two_set (default full logdet
model, v2p initialisation, dense and decimated supports), api.ICP_atlas with three structures, and the same loops under
the KeOps ordering of the M step (the product's default path)."""
import numpy as np
import pytest
import torch

from api_cases import run_atlas_2d, run_atlas_s3, run_two_set

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def spec():
    return {"device": dev(), "dtype": torch.float32}


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(dev())


@pytest.fixture
def default_spec_on_gpu(monkeypatch):
    """api.ICP_two_set builds its models with the default spec, like the reference (api/ICP_two_set.py:179-209): on a GPU
    box that is cuda (tools/spec.py:24-32).  Pin the device index so tensors created with cu() compare equal."""
    from diff_icp_b200.tools import spec as sp
    assert sp.defspec["device"] in ("cuda", torch.device("cuda"), torch.device("cuda:0"), "cuda:0")
    yield


@pytest.mark.parametrize("case,ordering", [("dense", "torch"), ("dense", "keops"), ("decim", "torch")])
def test_two_set_api_matches_reference(golden, monkeypatch, default_spec_on_gpu, case, ordering):
    run_two_set(golden, cu, monkeypatch, case, ordering)


@pytest.mark.parametrize("mid", ["0", "1"])
def test_two_set_api_midsize_support_matches_reference(golden, monkeypatch, default_spec_on_gpu, mid):
    """113 support points (decimation at 0.7 sigma_LDDMM) for 400 data points: the stage kernels in their one-launch form (mid = 0)
    and in the mid-size form (mid = 1: 128-register forward stage, ring adjoint over 64-column groups + finish launch), which the
    library only picks by itself when the data points fill the SMs; both against the run of the unmodified reference."""
    monkeypatch.setenv("DICP_SMALL_MID", mid)
    PSR = run_two_set(golden, cu, monkeypatch, "decimfine", "torch")
    assert PSR.q0[0].shape[0] == 113


@pytest.mark.parametrize("lockstep", [True, False])
@pytest.mark.parametrize("ordering", ["torch", "keops"])
def test_atlas_three_structures_matches_reference(golden, monkeypatch, ordering, lockstep):
    from diff_icp_b200.core.PSR import DiffPSR
    monkeypatch.setattr(DiffPSR, "batched_lbfgs", lockstep)
    run_atlas_s3(golden, cu, monkeypatch, spec(), ordering)


@pytest.mark.parametrize("mid", ["0", "1"])
def test_atlas_midsize_supports_matches_reference(golden, monkeypatch, mid):
    """4 frames x 3 structures with 78-94 support points per frame (decimation at 0.5 sigma_LDDMM), lock-step registration in two
    frame groups (DiffPSR.lockstep_groups = None: automatic above 64 support points); stage kernels in the one-launch form
    (mid = 0) and in the mid-size form (mid = 1), against the run of the unmodified reference.  Free-energy floor 1e-3 of the
    trace's scale instead of 5e-4: after the first registration all runs agree to 2e-5 (-1966.52 / -1966.53 here, -1966.56 / -1966.58
    reference fp32 / fp64); the second and third unconverged L-BFGS steps amplify the summation-order differences of the two
    kernel forms to 1 and 2.6 units of 3700 (the reference's own fp32 / fp64 runs drift apart by 0.7), one form above and one
    below the fp64 trace.  Momenta against the fp64 run: after the first registration 2e-4 / 7e-5 / 0 / 2.5e-4 per frame (reference
    fp32: 1.7e-4 / 5e-5 / 0 / 3e-5); after the second the reference's own fp32 run is off by 1.2e-2 and 1.6e-2 on two frames and
    the two kernel forms by 3e-2 on one resp. three frames -- hence the floor of 3 % of sigma_LDDMM on points and centroids."""
    monkeypatch.setenv("DICP_SMALL_MID", mid)
    PSR = run_atlas_s3(golden, cu, monkeypatch, spec(), "torch", name="atlas_mid", K=4, rho=0.5, fe_floor=1e-3, pt_floor=3e-2)
    assert min(int(q.shape[0]) for q in PSR.q0) > 64 and len(PSR._bplan) == 2


@pytest.mark.parametrize("lockstep", [True, False])
def test_atlas_2d_keops_ordering_matches_reference_loop(golden, monkeypatch, lockstep):
    from diff_icp_b200.core.PSR import DiffPSR
    monkeypatch.setattr(DiffPSR, "batched_lbfgs", lockstep)
    run_atlas_2d(golden, cu, monkeypatch, spec(), "keops")
