"""CPU: the whole host-side stack (DiffPSR alternation, L-BFGS driver, shooting adjoint sweeps, support schemes, API
option handling) driven by the CPU emulation of the kernels' arithmetic, against the reference's own end-to-end run."""
import numpy as np
import pytest
import torch

import emu_backend

CPU = {"device": "cpu", "dtype": torch.float32}


def T32(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float32))


def test_c1_registration_matches_reference_end_to_end(golden, monkeypatch):
    emu_backend.install_all(monkeypatch)
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.core.PSR import DiffPSR
    g = golden("psr")
    x0, mu0 = T32(g["c1_in_x0"]), T32(g["c1_in_mu"])
    G = GaussianMixtureUnif(mu0, sigma=0.1, spec=CPU, computversion="torch")
    G.to_optimize = {"mu": False, "sigma": True, "w": False, "eta0": False}
    LM = LDDMMModel(sigma=0.2, D=2, lambd=5e2, version="classic", scheme="Euler", spec=CPU)
    P = DiffPSR(x0, G, LM, dataspec=CPU, compspec=CPU)
    P.printstuff = False
    P.set_support_scheme("grid", rho=np.sqrt(2))
    assert np.abs(P.q0[0].numpy() - g["c1_gold_q0"]).max() < 1e-6
    fes, sigs = [], []
    for it in range(2):
        P.GMM_opt()
        fes.append(P.FE)
        P.Reg_opt(tol=1e-5)
        fes.append(P.FE)
        sigs.append(P.GMMi[0].sigma)
    assert np.allclose(fes, g["c1_gold_FE"][:4], rtol=2e-4), (fes, g["c1_gold_FE"][:4])
    assert np.allclose(sigs, g["c1_gold_sigma"][:2], rtol=2e-4)
    reg = P.Registration(0)
    assert reg.q0 is P.q0[0] and reg.a0 is P.a0[0]


def test_dense_scheme_and_structures(monkeypatch):
    """Two structures, dense support (q = all points of the frame), hybrid model, Ralston: shapes and bookkeeping."""
    emu_backend.install_all(monkeypatch)
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.core.PSR import DiffPSR
    g = torch.Generator().manual_seed(0)
    frames = [[torch.rand(40, 2, generator=g), 2 + torch.rand(30, 2, generator=g)] for _ in range(2)]
    gm = [GaussianMixtureUnif(torch.rand(4, 2, generator=g), sigma=0.3, spec=CPU),
          GaussianMixtureUnif(2 + torch.rand(3, 2, generator=g), sigma=0.3, spec=CPU)]
    LM = LDDMMModel(sigma=0.5, D=2, lambd=50.0, version="hybrid", scheme="Ralston", nt=3, spec=CPU)
    P = DiffPSR(frames, gm, LM, dataspec=CPU, compspec=CPU)
    P.printstuff = False
    assert P.K == 2 and P.S == 2 and P.q0[0].shape == (70, 2) and torch.count_nonzero(P.a0[0]) == 0
    fe0 = P.FE
    P.GMM_opt(max_iterations=3)
    P.Reg_opt(nmax=1)
    assert P.FE < fe0
    assert P.x1[1, 1].shape == (30, 2) and len(P.shoot[0]) == 4 and len(P.shoot[0][-1]) == 3


def test_api_option_handling(monkeypatch):
    emu_backend.install_all(monkeypatch)
    from diff_icp_b200.api.ICP_atlas import ICP_atlas
    from diff_icp_b200.api.ICP_two_set import ICP_two_set
    x = [torch.rand(30, 2) for _ in range(2)]
    with pytest.raises(AssertionError):
        ICP_atlas(x, GMM_parameters={"init_components": "five"}, registration_parameters={"type": "diffeomorphic"})
    with pytest.raises(NotImplementedError):
        ICP_atlas(x, GMM_parameters={"init_components": 3}, registration_parameters={"type": "rigid"})
    with pytest.raises(AssertionError):
        ICP_two_set(x[0], x[1], {"sigma": 0.1}, {"type": "diffeomorphic", "lambda_LDDMM": 1.0, "sigma_LDDMM": 0.2})
    torch.manual_seed(0)
    PSR, evol = ICP_atlas(x, GMM_parameters={"init_components": 3},
                          registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.3},
                          numerical_options={"compspec": CPU, "dataspec": CPU, "support_LDDMM": {"scheme": "grid", "rho": 1.0}},
                          optim_options={"max_iterations": 2, "max_repeat_GMM": 3}, printstuff=False)
    assert PSR.LMi.withlogdet and not PSR.LMi.gradcomponent and PSR.LMi.scheme == "Euler" and PSR.LMi.nt == 10
    assert PSR.support_scheme == "grid" and len(evol["a0"]) == 2 and np.isfinite(PSR.FE)


def test_lockstep_frame_groups_rule():
    """DiffPSR._frame_groups: one group for small supports, min(4, K // 2) contiguous groups above 64 support points (the other
    groups' launches fill the partly filled last waves and the late lock-step rounds of a group), a forced count subject to
    lockstep_group_min_frames; every frame in exactly one group, in order."""
    from diff_icp_b200.core.PSR import DiffPSR

    class Stub:
        lockstep_groups = None
        lockstep_group_min_frames = 8
        _frame_groups = DiffPSR._frame_groups

    def groups(K, M, forced=None, gmin=8):
        s = Stub()
        s.K, s.q0 = K, [torch.zeros(M if k else max(M - 3, 1), 2) for k in range(K)]
        s.lockstep_groups, s.lockstep_group_min_frames = forced, gmin
        g = s._frame_groups()
        assert [k for grp in g for k in grp] == list(range(K)) and all(len(grp) > 0 for grp in g)
        return [len(grp) for grp in g]

    assert groups(64, 25) == [64]                       # small supports: one group (device L-BFGS, one-launch closure)
    assert groups(64, 64) == [64]
    assert groups(64, 65) == [16, 16, 16, 16]
    assert groups(8, 1210) == [2, 2, 2, 2]
    assert groups(5, 300) == [2, 3] or groups(5, 300) == [3, 2]
    assert groups(3, 300) == [3]
    assert groups(1, 300) == [1]
    assert groups(64, 25, forced=2) == [32, 32]
    assert groups(10, 25, forced=2) == [10]             # forced counts need 2 x lockstep_group_min_frames frames
    assert groups(10, 25, forced=2, gmin=4) == [5, 5]
