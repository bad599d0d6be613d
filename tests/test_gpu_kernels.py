"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the reference-facing
Python API and the C ABI, against the golden fixtures (reference outputs) and the fp64 oracle.

Tolerances: kernel sums, momenta, deformed points: max|out-gold| <= 1e-5 * max|gold| for a single
reduction / right-hand side, 5e-5 after an integrated trajectory; scalars (H, loss): rel 1e-5; gradients 2e-4.
"""
import numpy as np
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(dev())


@pytest.fixture(scope="module")
def GK():
    from diff_icp_b200.tools.kernel import GaussKernel
    return GaussKernel


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_ten_reductions_match_reference_gold(golden, GK, tag):
    g = golden("kernels")
    M, N, D, sig = g[f"{tag}_meta"]
    x, y, b, c, d = (cu(g[f"{tag}_in_{k}"]) for k in "xybcd")
    K = GK(float(sig), int(D), spec={"device": dev(), "dtype": torch.float32})
    got = {
        "KBase": K.KBase(x, y), "KRedScal": K.KRedScal(x, y, d), "KRed": K.KRed(x, y, b),
        "GradKRed": K.GradKRed(x, y), "DDKRed": K.DDKRed(x, y, b), "GenDKRed": K.GenDKRed(x, y, b, c),
        "HessKRed": K.HessKRed(x, y, b, c), "LapKRed": K.LapKRed(x, y), "GradLapKRed": K.GradLapKRed(x, y),
        "GradKRed_rev": K.GradKRed_rev(x, y, c),
    }
    for k, v in got.items():
        gold = g[f"{tag}_gold_{k}"]
        assert tuple(v.shape) == tuple(gold.shape), k
        e = relerr(v.cpu().numpy(), gold)
        e_ref = relerr(g[f"{tag}_ref32_{k}"], gold)
        assert e < max(1e-5, 2 * e_ref), (k, e, e_ref)


@pytest.mark.parametrize("M,N,D", [(1, 1, 2), (5, 3, 3), (300, 129, 2), (257, 1000, 3), (4000, 9000, 3)])
def test_kred_shapes_and_splits_vs_oracle(GK, M, N, D):
    from oracle.kernels import GaussOracle
    g = torch.Generator().manual_seed(M * 7 + N)
    x, y, b = torch.rand(M, D, generator=g), torch.rand(N, D, generator=g), torch.randn(N, D, generator=g)
    sig = 0.2
    K = GK(sig, D, spec={"device": dev(), "dtype": torch.float32})
    out = K.KRed(x.to(dev()), y.to(dev()), b.to(dev())).cpu()
    ref = GaussOracle(sig, D).KRed(x.double(), y.double(), b.double())
    assert relerr(out.numpy(), ref.numpy()) < 1e-5
    # determinism: bit-identical on a second call
    out2 = K.KRed(x.to(dev()), y.to(dev()), b.to(dev())).cpu()
    assert torch.equal(out, out2)


def test_many_rows_single_split_path(GK):
    """More row blocks than CTA slots: the in-kernel finish path (no column split)."""
    from oracle.kernels import GaussOracle
    M, N, D, sig = 400_000, 150, 3, 0.3
    g = torch.Generator().manual_seed(3)
    x, y, b = torch.rand(M, D, generator=g), torch.rand(N, D, generator=g), torch.randn(N, D, generator=g)
    K = GK(sig, D, spec={"device": dev(), "dtype": torch.float32})
    out = K.KRed(x.to(dev()), y.to(dev()), b.to(dev())).cpu()
    lap = K.LapKRed(x.to(dev()), y.to(dev())).cpu()
    O = GaussOracle(sig, D, chunk=20000)
    assert relerr(out.numpy(), O.KRed(x.double(), y.double(), b.double()).numpy()) < 1e-5
    assert relerr(lap.numpy(), O.LapKRed(x.double(), y.double()).numpy()) < 1e-5


def test_offset_cloud_keeps_accuracy(GK):
    """Data far from the origin (the conditioning case of SURVEY Appendix B): origin-relative prescaling."""
    from oracle.kernels import GaussOracle
    M, N, D, sig = 500, 800, 3, 0.05
    g = torch.Generator().manual_seed(9)
    x, y, b = 100 + torch.rand(M, D, generator=g), 100 + torch.rand(N, D, generator=g), torch.randn(N, D, generator=g)
    K = GK(sig, D, spec={"device": dev(), "dtype": torch.float32})
    out = K.KRed(x.to(dev()), y.to(dev()), b.to(dev())).cpu()
    ref = GaussOracle(sig, D).KRed(x.double(), y.double(), b.double())
    assert relerr(out.numpy(), ref.numpy()) < 2e-5


def test_check_coverage_bit_exact(GK):
    from oracle.kernels import GaussOracle
    M, N, D, sig = 3000, 77, 2, 0.1
    g = torch.Generator().manual_seed(4)
    X, Y = torch.rand(M, D, generator=g), torch.rand(N, D, generator=g)
    K = GK(sig, D, spec={"device": dev(), "dtype": torch.float32})
    got = K.check_coverage(X.to(dev()), Y.to(dev()), 1.0).cpu()
    O = GaussOracle(sig, D)
    d2 = O.min_sqdist(X.double(), Y.double())
    thr = (1.0 * sig) ** 2
    safe = (d2 - thr).abs() > 1e-5 * thr          # rows whose decision does not hinge on fp32 rounding
    ref = d2 > thr
    assert got.dtype == torch.bool and got.shape == (M,)
    assert torch.equal(got[safe], ref[safe])
    assert int((~safe).sum()) < 5


@pytest.fixture(params=["small_support_stage_kernels", "general_pair_engine"])
def shoot_path(request):
    """The golden cases have 37 support points, which the product routes to the one-launch-per-stage kernels
    (csrc/small_step.cuh); the same cases are also forced through the general tiled engine."""
    from diff_icp_b200 import ops, shooting
    old = ops.small_enabled
    ops.small_enabled = request.param == "small_support_stage_kernels"
    shooting.ShootPlan._cache.clear()
    shooting.ClosurePlan._cache.clear()
    yield request.param
    ops.small_enabled = old
    shooting.ShootPlan._cache.clear()
    shooting.ClosurePlan._cache.clear()


def test_lddmm_shoot_and_gradients_match_reference_gold(golden, shoot_path):
    from diff_icp_b200.core.LDDMM import LDDMMModel
    g = golden("lddmm")
    spec = {"device": dev(), "dtype": torch.float32}
    for tag in g["cases"]:
        tag = str(tag)
        D, Nq, Nx, nt, sig, lam = g[f"{tag}_meta"]
        _, version, scheme, xm = tag.split("_")
        LM = LDDMMModel(sigma=float(sig), D=int(D), lambd=float(lam), spec=spec, version=version, scheme=scheme, nt=int(nt))
        q = cu(g[f"{tag}_in_q0"]).requires_grad_(True)
        p = cu(g[f"{tag}_in_p0"]).requires_grad_(True)
        x = cu(g[f"{tag}_in_x0"]).requires_grad_(True) if xm == "x" else None
        y, sig2 = cu(g[f"{tag}_in_y"]), cu(g[f"{tag}_in_sig2"])
        # single right-hand side
        ode = LM.ODE(q.detach(), p.detach(), torch.zeros(1, device=dev()), None if x is None else x.detach())
        assert relerr(ode[0].cpu().numpy(), g[f"{tag}_gold_ode_vq"]) < 1e-5, tag
        assert relerr(ode[1].cpu().numpy(), g[f"{tag}_gold_ode_dp"]) < 1e-5, tag
        dc = float(g[f"{tag}_gold_ode_dcost"])
        assert abs(float(ode[2].sum()) - dc) < 1e-5 * max(1.0, abs(dc)), tag
        if x is not None:
            assert relerr(ode[3].cpu().numpy(), g[f"{tag}_gold_ode_vx"]) < 1e-5, tag
        # shooting
        sh = LM.Shoot(q, p, x)
        assert len(sh) == int(nt) + 1 and len(sh[-1]) == (4 if x is not None else 3) and sh[-1][2].shape == (1,)
        assert relerr(sh[-1][0].detach().cpu().numpy(), g[f"{tag}_gold_q1"]) < 5e-5, tag
        assert relerr(sh[-1][1].detach().cpu().numpy(), g[f"{tag}_gold_p1"]) < 5e-5, tag
        assert relerr(sh[int(nt) // 2][0].detach().cpu().numpy(), g[f"{tag}_gold_qmid"]) < 5e-5, tag
        c1 = float(g[f"{tag}_gold_cost1"][0])
        assert abs(float(sh[-1][2]) - c1) < 5e-5 * max(1.0, abs(c1)), tag
        if x is not None:
            assert relerr(sh[-1][3].detach().cpu().numpy(), g[f"{tag}_gold_x1"]) < 5e-5, tag
        tl = LM.trajloss(sh)
        assert abs(float(tl) - float(g[f"{tag}_gold_trajloss"])) < 2e-5 * max(1.0, abs(float(g[f"{tag}_gold_trajloss"]))), tag
        moved = sh[-1][3] if x is not None else sh[-1][0]
        L = tl + ((moved - y) ** 2 / (2 * sig2[:, None])).sum()
        assert abs(float(L) - float(g[f"{tag}_gold_loss"])) < 2e-5 * abs(float(g[f"{tag}_gold_loss"])), tag
        L.backward()
        assert relerr(p.grad.cpu().numpy(), g[f"{tag}_gold_gp0"]) < 2e-4, (tag, relerr(p.grad.cpu().numpy(), g[f"{tag}_gold_gp0"]))
        assert relerr(q.grad.cpu().numpy(), g[f"{tag}_gold_gq0"]) < 2e-4, (tag, relerr(q.grad.cpu().numpy(), g[f"{tag}_gold_gq0"]))
        if x is not None:
            assert relerr(x.grad.cpu().numpy(), g[f"{tag}_gold_gx0"]) < 2e-4, tag


def test_cuda_graph_shoot_is_bit_identical_to_eager(shoot_path):
    from diff_icp_b200.core.LDDMM import LDDMMModel
    spec = {"device": dev(), "dtype": torch.float32}
    g = torch.Generator().manual_seed(1)
    q0, p0, x0 = torch.rand(300, 2, generator=g), 0.2 * torch.randn(300, 2, generator=g), torch.rand(900, 2, generator=g)
    res = []
    for use_graph in (False, True, True):
        LM = LDDMMModel(sigma=0.2, D=2, lambd=10.0, spec=spec, version="hybrid", scheme="Ralston", nt=6)
        LM.use_cuda_graph = use_graph
        p = p0.to(dev()).requires_grad_(True)
        sh = LM.Shoot(q0.to(dev()), p, x0.to(dev()))
        L = LM.trajloss(sh) + (sh[-1][3] ** 2).sum()
        L.backward()
        res.append((L.detach().cpu(), p.grad.cpu(), sh[-1][3].detach().cpu()))
    for r in res[1:]:
        assert torch.equal(r[0], res[0][0]) and torch.equal(r[1], res[0][1]) and torch.equal(r[2], res[0][2])


def test_hamiltonian_is_conserved_and_flow_inverts(shoot_path):
    """Known-answer identities of SURVEY.md §4: H drift along a Ralston shoot, backward(apply(X)) = X for eta = 0."""
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.core.registrations import LDDMMRegistration
    spec = {"device": dev(), "dtype": torch.float32}
    g = torch.Generator().manual_seed(2)
    q0 = torch.rand(200, 3, generator=g).to(dev())
    p0 = (0.05 * torch.randn(200, 3, generator=g)).to(dev())
    X = torch.rand(500, 3, generator=g).to(dev())
    LM = LDDMMModel(sigma=0.3, D=3, lambd=10.0, spec=spec, version="classic", scheme="Ralston", nt=20)
    sh = LM.Shoot(q0, p0)
    H0 = float(LM.Hamiltonian(q0, p0))
    H1 = float(LM.Hamiltonian(sh[-1][0], sh[-1][1]))
    assert abs(float(sh.H0) - H0) < 1e-6 * abs(H0)            # fused Hamiltonian pieces == standalone reductions
    assert abs(H1 - H0) < 1e-3 * abs(H0)                      # Ralston keeps H to O(dt^2)
    from oracle.lddmm import LDDMMOracle
    OR = LDDMMOracle(sigma=0.3, D=3, lambd=10.0, version="classic", scheme="Ralston", nt=20)
    so = OR.shoot(q0.cpu().double(), p0.cpu().double())
    drift_o = float(OR.hamiltonian(so[-1][0], so[-1][1]) - OR.hamiltonian(q0.cpu().double(), p0.cpu().double()))
    assert abs((H1 - H0) - drift_o) < 2e-5 * abs(H0)          # and drifts exactly like the fp64 oracle does
    reg = LDDMMRegistration(LM, q0, p0)
    fwd = reg.apply(X)
    back = reg.backward(fwd)
    assert float((back - X).abs().max()) < 2e-4               # inverse flow by re-shooting (q1,-p1): exact up to O(dt^2)
    xo = so = OR.shoot(q0.cpu().double(), p0.cpu().double(), X.cpu().double())
    assert float((fwd.cpu().double() - xo[-1][3]).abs().max()) < 1e-5
    bo = OR.shoot(xo[-1][0], -xo[-1][1], xo[-1][3])[-1][3]
    assert float((back.cpu().double() - bo).abs().max()) < 1e-5   # same round trip as the fp64 oracle


def test_library_fails_loudly_on_cpu_tensors(GK):
    K = GK(0.2, 2)
    with pytest.raises(ValueError):
        K.KRed(torch.rand(4, 2), torch.rand(5, 2), torch.rand(5, 2))
