"""CPU: the API entry points end to end on the CPU emulation of the kernels' arithmetic, against runs of the reference
(api.ICP_two_set with its default logdet model and the v2p initialisation; api.ICP_atlas with three structures)."""
import numpy as np
import pytest
import torch

import emu_backend
from api_cases import run_atlas_2d, run_atlas_s3, run_two_set

CPU = {"device": "cpu", "dtype": torch.float32}


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))


@pytest.mark.parametrize("case,ordering", [("dense", "keops"), ("decim", "torch")])
def test_two_set_api_matches_reference(golden, monkeypatch, case, ordering):
    emu_backend.install_all(monkeypatch)
    run_two_set(golden, _t, monkeypatch, case, ordering)


def test_atlas_three_structures_matches_reference(golden, monkeypatch):
    emu_backend.install_all(monkeypatch)
    run_atlas_s3(golden, _t, monkeypatch, CPU, "torch")


def test_atlas_2d_keops_ordering_matches_reference_loop(golden, monkeypatch):
    emu_backend.install_all(monkeypatch)
    run_atlas_2d(golden, _t, monkeypatch, CPU, "keops")
