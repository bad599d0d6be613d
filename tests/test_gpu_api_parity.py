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


@pytest.mark.parametrize("lockstep", [True, False])
def test_atlas_2d_keops_ordering_matches_reference_loop(golden, monkeypatch, lockstep):
    from diff_icp_b200.core.PSR import DiffPSR
    monkeypatch.setattr(DiffPSR, "batched_lbfgs", lockstep)
    run_atlas_2d(golden, cu, monkeypatch, spec(), "keops")
