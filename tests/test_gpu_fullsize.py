"""GPU: the CUDA path at BASELINE.json's full sizes (20k x 20k, 3-D), checked through size-independent properties and
row-subset comparisons with the oracle; plus a mid-size (5k) full gradient check that exercises the column-split grid."""
import numpy as np
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu
M_FULL = 20000


def dev():
    return torch.device("cuda:0")


def spec():
    return {"device": dev(), "dtype": torch.float32}


def test_kernel_sum_properties_at_20k():
    from diff_icp_b200.tools.kernel import GaussKernel
    from oracle.kernels import GaussOracle
    g = torch.Generator().manual_seed(0)
    x = torch.rand(M_FULL, 3, generator=g)
    b1, b2 = torch.randn(M_FULL, 3, generator=g), torch.randn(M_FULL, 3, generator=g)
    sig = 0.1
    K = GaussKernel(sig, 3, spec=spec())
    xd, b1d, b2d = x.to(dev()), b1.to(dev()), b2.to(dev())
    r1, r2 = K.KRed(xd, xd, b1d), K.KRed(xd, xd, b2d)
    r12 = K.KRed(xd, xd, 0.7 * b1d + b2d)
    assert float((r12 - (0.7 * r1 + r2)).abs().max() / r12.abs().max()) < 1e-5                    # linearity in b
    lhs, rhs = float((b2d * r1).sum()), float((b1d * r2).sum())
    assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), float((b2d.abs() * r1.abs()).sum()) * 1e-3)      # symmetry of K
    rows = torch.arange(0, M_FULL, 313)                                                           # 64 rows vs all columns
    ref = GaussOracle(sig, 3).KRed(x[rows].double(), x.double(), b1.double())
    assert relerr(r1[rows.to(dev())].cpu().numpy(), ref.numpy()) < 1e-5


@pytest.mark.parametrize("version", ["classic", "hybrid", "logdet"])
def test_rhs_properties_at_20k(version):
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from oracle.lddmm import LDDMMOracle
    g = torch.Generator().manual_seed(1)
    q = torch.rand(M_FULL, 3, generator=g)
    p = 1e-2 * torch.randn(M_FULL, 3, generator=g)
    LM = LDDMMModel(sigma=0.1, D=3, lambd=50.0, spec=spec(), version=version, scheme="Euler", nt=4)
    vq, dp, dcost = LM.ODE(q.to(dev()), p.to(dev()), torch.zeros(1, device=dev()))
    if version != "logdet":
        # Newton's third law of the classic interaction: sum_i dp_i = 0 (the pair term is antisymmetric)
        assert float(dp.sum(0).abs().max()) < 1e-5 * float(dp.abs().sum(0).max())
    # linearity of vq in p for eta = 0, affine for logdet: vq(2p) - vq(p) = KRed(q,q,p)
    vq2, _, _ = LM.ODE(q.to(dev()), (2 * p).to(dev()), torch.zeros(1, device=dev()))
    kr = LM.Kernel.KRed(q.to(dev()), q.to(dev()), p.to(dev()))
    assert float((vq2 - vq - kr).abs().max() / kr.abs().max()) < 2e-5
    # a row subset against the oracle evaluated on those rows (x = subset of q, same formulas as v / mdivsum)
    rows = torch.arange(0, M_FULL, 625)
    OR = LDDMMOracle(sigma=0.1, D=3, lambd=50.0, version=version)
    ref_v = OR.v(q[rows].double(), q.double(), p.double())
    assert relerr(vq[rows.to(dev())].cpu().numpy(), ref_v.numpy()) < 1e-5
    ref_G = OR.K.GenDKRed(q[rows].double(), q.double(), p.double(), p[rows].double())
    if version == "logdet":
        ref_G = ref_G - OR.eta * OR.K.HessKRed(q[rows].double(), q.double(), p.double(), p[rows].double()) \
            - OR.eta ** 2 * OR.K.GradLapKRed(q[rows].double(), q.double())
    assert relerr(dp[rows.to(dev())].cpu().numpy(), (-ref_G).numpy()) < 1e-5


def test_em_row_pass_at_20k_by_20k():
    """E step of the two-set configuration (20k points x 20k frozen centroids): targets are convex combinations of the
    centroids, sum_c gamma = 1, and a row subset matches the oracle."""
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from oracle.gmm import GMMOracle
    g = torch.Generator().manual_seed(2)
    xB = torch.rand(M_FULL, 3, generator=g)
    xA = xB[torch.randperm(M_FULL, generator=g)] + 0.02 * torch.randn(M_FULL, 3, generator=g)
    G = GaussianMixtureUnif(xB.to(dev()), sigma=0.05, spec=spec())
    G.to_optimize = {"mu": False, "sigma": True, "w": False, "eta0": False}
    Y, Cfe, FE = G.EM_step(xA.to(dev()))
    assert bool(((Y >= -1e-4) & (Y <= 1 + 1e-4)).all())
    rows = torch.arange(0, M_FULL, 400)
    O = GMMOracle(xB.double(), 0.05, to_optimize={"mu": False, "sigma": False, "w": False})
    Yo, _, _ = O.em_step(xA[rows].double(), skip_M=True)
    assert relerr(Y[rows.to(dev())].cpu().numpy(), Yo.numpy()) < 2e-5
    assert 0.01 < G.sigma < 0.05


@pytest.mark.parametrize("version,scheme", [("hybrid", "Ralston"), ("logdet", "Euler")])
def test_midsize_gradient_vs_oracle(version, scheme):
    """5000 support points (column-split grid, several tiles, odd count): loss and full gradient against the fp64 oracle."""
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from oracle.lddmm import LDDMMOracle
    g = torch.Generator().manual_seed(3)
    M = 4999
    q = torch.rand(M, 3, generator=g)
    p = 2e-3 * torch.randn(M, 3, generator=g)
    y = q + 0.02 * torch.randn(M, 3, generator=g)
    LM = LDDMMModel(sigma=0.15, D=3, lambd=20.0, spec=spec(), version=version, scheme=scheme, nt=3)
    pd = p.to(dev()).requires_grad_(True)
    sh = LM.Shoot(q.to(dev()), pd)
    L = LM.trajloss(sh) + ((sh[-1][0] - y.to(dev())) ** 2).sum() * 50.0
    L.backward()
    OR = LDDMMOracle(sigma=0.15, D=3, lambd=20.0, version=version, scheme=scheme, nt=3, chunk=1024)
    po = p.double().requires_grad_(True)
    Lo, _ = OR.loss(q.double(), po, None, y.double(), 50.0)
    (go,) = torch.autograd.grad(Lo, [po])
    assert abs(float(L) - float(Lo)) < 2e-5 * abs(float(Lo))
    assert relerr(pd.grad.cpu().numpy(), go.numpy()) < 2e-4


def test_stress_one_million_points_forward():
    """BASELINE configs[4] size (3-D, 10^6 control points): one fused right-hand side (10^12 pairs, in-kernel finish path,
    64-bit partial indexing), checked by Newton's third law, by sum_i p_i.vq_i = A, and on a row subset against the oracle."""
    from diff_icp_b200 import ops
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from oracle.kernels import GaussOracle
    M = 1_000_000
    g = torch.Generator().manual_seed(5)
    q = torch.rand(M, 3, generator=g)
    p = 1e-3 * torch.randn(M, 3, generator=g)
    LM = LDDMMModel(sigma=0.05, D=3, lambd=100.0, spec=spec(), version="classic", scheme="Ralston", nt=10)
    qd, pd = q.to(dev()), p.to(dev())
    vq, dp, dcost = LM.ODE(qd, pd, torch.zeros(1, device=dev()))
    assert float(dp.sum(0).abs().max()) < 2e-5 * float(dp.abs().sum(0).max())
    rows = torch.arange(0, M, 31250)                                           # 32 rows against all 10^6 columns
    ref = GaussOracle(0.05, 3).KRed(q[rows].double(), q.double(), p.double())
    assert relerr(vq[rows.to(dev())].cpu().numpy(), ref.numpy()) < 2e-5
    assert float(dcost.abs().sum()) == 0.0


def test_multi_structure_3d_grid_support_atlas_runs():
    """BASELINE configs[3] in miniature: several frames x 3 structures, 3-D, grid support in 3-D (an extension: the
    reference's grid is 2-D only), hybrid model; the free energy must decrease and every structure keeps its own GMM."""
    from diff_icp_b200.api.ICP_atlas import ICP_atlas
    g = torch.Generator().manual_seed(6)
    K, S = 4, 3
    cents = [torch.rand(5, 3, generator=g) + 2.0 * s for s in range(S)]
    x0 = [[(cents[s][torch.randint(0, 5, (600 + 50 * k,), generator=g)] + 0.05 * torch.randn(600 + 50 * k, 3, generator=g)).to(dev())
           for s in range(S)] for k in range(K)]
    torch.manual_seed(0)
    fes = []
    PSR, evol = ICP_atlas(x0, GMM_parameters={"init_components": 5},
                          registration_parameters={"type": "diffeomorphic", "lambda_LDDMM": 100.0, "sigma_LDDMM": 0.5},
                          numerical_options={"compspec": spec(), "dataspec": spec(), "support_LDDMM": {"scheme": "grid", "rho": 1.5}},
                          optim_options={"max_iterations": 3, "max_repeat_GMM": 5},
                          callback_function=lambda P, before: fes.append(P.FE), printstuff=False)
    assert PSR.S == 3 and PSR.K == 4 and PSR.q0[0].shape[1] == 3 and len(PSR.GMMi) == 3
    assert PSR.x1[3, 2].shape == (750, 3)
    assert fes[-1] < fes[0] and np.isfinite(PSR.FE)


def test_empty_and_tiny_inputs():
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.tools.kernel import GaussKernel
    K = GaussKernel(0.3, 2, spec=spec())
    x, y, b = torch.rand(0, 2, device=dev()), torch.rand(7, 2, device=dev()), torch.rand(7, 2, device=dev())
    assert K.KRed(x, y, b).shape == (0, 2) and K.KBase(x, y).shape == (0,)
    LM = LDDMMModel(sigma=0.3, D=2, lambd=5.0, spec=spec(), version="hybrid", scheme="Euler", nt=3)
    q, p = torch.rand(1, 2, device=dev()), torch.rand(1, 2, device=dev())
    sh = LM.Shoot(q, p)                                         # a single support point: K = 1, dp = 0, straight line
    assert torch.allclose(sh[-1][0], q + p, atol=1e-6) and torch.allclose(sh[-1][1], p, atol=1e-7)
    assert LM.v(torch.rand(0, 2, device=dev()), q, p).shape == (0, 2)


def test_groupwise_iteration_at_atlas_size_lockstep():
    """configs[2] size: 64 frames x 10k points, 2-D, C = 50, hybrid, grid support -- the lock-step registration of all
    frames.  Size-independent properties: the free energy never increases over the alternation, two identical runs agree
    bit for bit, every frame's stored shoot is the shoot of its stored momenta, and the uncovered-point counts of the
    batched coverage kernel equal GaussKernel.check_coverage on a frame."""
    import math
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from groupwise_iteration import spiral_frames
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.core.PSR import DiffPSR

    def run():
        frames = spiral_frames(64, 10000)
        torch.manual_seed(1234)
        G = GaussianMixtureUnif(torch.zeros(50, 2), spec=spec())
        LM = LDDMMModel(sigma=0.2, D=2, lambd=500.0, version="hybrid", scheme="Euler", nt=10, spec=spec())
        LM.use_cuda_graph = True
        P = DiffPSR([f.to(dev()) for f in frames], G, LM, dataspec=spec(), compspec=spec())
        P.printstuff = False
        P.set_support_scheme("grid", rho=math.sqrt(2))
        P.reinitialize_GMM()
        fes = [P.FE]
        for _ in range(2):
            P.GMM_opt(max_iterations=10, tol=1e-3)
            fes.append(P.FE)
            P.Reg_opt(tol=1e-3, nmax=1)
            fes.append(P.FE)
        return P, fes

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Pa, fa = run()
        Pb, fb = run()
    assert Pa._batched_plan() is not None                       # the lock-step path was the one that ran
    assert all(b <= a + 1e-6 * abs(a) for a, b in zip(fa, fa[1:])), fa
    assert fa == fb
    for k in (0, 31, 63):
        assert torch.equal(Pa.a0[k], Pb.a0[k])
        sh = Pa.LMi.Shoot(Pa.q0[k], Pa.a0[k], Pa.allx0[k])
        assert torch.allclose(sh[-1][3], Pa.x1[k, 0], atol=2e-6)
        mine = sum(int(Pa.LMi.Kernel.check_coverage(st[-1], st[0], 2.0).sum()) for st in Pa.shoot[k])
        ref = sum(int(Pa.LMi.Kernel.check_coverage(st[-1], st[0], 2.0).sum()) for st in sh)
        assert abs(mine - ref) <= 2
