"""GPU: the batched (multi-frame) registration closure and the lock-step Reg_opt against the one-frame-at-a-time path
(itself parity-tested against the reference's gradients and end-to-end runs in test_gpu_kernels / test_gpu_em_psr)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def spec():
    return {"device": dev(), "dtype": torch.float32}


def frames(D, sizes, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(n, D, generator=g) for n in sizes]


CASES = [
    # D, version, scheme, nt, support sizes, data sizes (0 = dense small support: data points are the support points)
    (2, "hybrid", "Euler", 10, [25, 25, 25, 25, 25], [1000, 1300, 777, 128, 1701]),
    (3, "logdet", "Ralston", 4, [40, 17, 130, 64], [900, 2500, 300, 1111]),
    (2, "classic", "Ralston", 5, [33, 200, 7], [0, 0, 0]),
    (3, "hybrid", "Euler", 6, [300, 150], [5000, 20000]),
    # above 512 support points the per-frame reference runs on the general tiled engine: cross-checks the two engines
    (3, "hybrid", "Ralston", 3, [900, 640], [6000, 2500]),
    (2, "logdet", "Euler", 3, [1024, 513], [0, 0]),
    # many frames x many data points: the x-row CTAs take several blocks of 128 rows each (xpass > 1 in small_step.cuh)
    (2, "hybrid", "Euler", 3, [25] * 40, [13000 + 10 * k for k in range(40)]),
    (3, "logdet", "Ralston", 2, [30] * 36, [14000] * 36),
    # sizes at the limits of the one-launch closure (one thread-block cluster per frame): 64 support points, 16384 data points
    (3, "classic", "Euler", 5, [12, 64, 33], [700, 3000, 16384]),
    (3, "hybrid", "Euler", 4, [64, 1, 2], [16384, 5, 2049]),
    (2, "hybrid", "Euler", 3, [40, 9], [32768, 12345]),
    # mid-size supports x many data points: the one-evaluation-per-pair adjoint stage of small_step.cuh (small_adj_mid_kernel: the
    # 512-row x CTAs of all frames fill the SMs) and the 128-register forward instantiation, up to 2048 support points
    (3, "hybrid", "Ralston", 2, [1500, 700, 65], [30000, 26000, 29999]),
    (3, "logdet", "Euler", 2, [300, 2048], [40000, 40001]),
    (2, "classic", "Euler", 2, [129, 1210, 100, 777], [30000, 513, 28000, 30001]),
    (2, "hybrid", "Ralston", 2, [1025, 200], [45000, 38000]),
    (3, "hybrid", "Euler", 2, [300, 7, 64, 1], [30000, 26000, 512, 40000]),      # tiny supports in a mid-size batch
    (2, "logdet", "Ralston", 2, [500, 200], [45000, 38000]),
]


# one_launch: False = stage kernels; True = launch shape chosen by the library; "T,cluster" = that shape forced (DICP_CC_SHAPE)
@pytest.mark.parametrize("one_launch", [True, False, "64,8", "64,16", "128,8", "128,16"])
@pytest.mark.parametrize("D,version,scheme,nt,Ms,Nxs", CASES)
def test_batched_closure_equals_per_frame_closure(D, version, scheme, nt, Ms, Nxs, one_launch, monkeypatch):
    """one_launch=True: the whole closure of a frame in ONE kernel (csrc/cluster_closure.cuh) where it applies (eta = 0, data
    points, Euler, <= 64 support points, <= 32768 data points); False: the stage kernels, one launch per integrator stage."""
    from diff_icp_b200 import shooting
    from diff_icp_b200.core.LDDMM import LDDMMModel
    if isinstance(one_launch, str):
        T, cl = (int(v) for v in one_launch.split(","))
        if not (version != "logdet" and scheme == "Euler" and min(Nxs) > 0 and max(Ms) <= T // 2 and max(Nxs) <= cl * 2048):
            pytest.skip("shape does not apply to these sizes")
        monkeypatch.setenv("DICP_CC_SHAPE", one_launch)
        one_launch = True
    monkeypatch.setattr(shooting.BatchedClosurePlan, "one_launch_closure", one_launch)
    sig, lam = 0.25, 50.0
    LM = LDDMMModel(sigma=sig, D=D, lambd=lam, version=version, scheme=scheme, nt=nt, spec=spec())
    K = len(Ms)
    g = torch.Generator().manual_seed(3)
    q0 = [torch.rand(m, D, generator=g).to(dev()) for m in Ms]
    x0 = [torch.rand(n, D, generator=g).to(dev()) if n else None for n in Nxs]
    nd = [n if n else m for m, n in zip(Ms, Nxs)]
    y = [torch.rand(n, D, generator=g).to(dev()) for n in nd]
    inv = [(0.5 + torch.rand(n, generator=g)).to(dev()) * 20 for n in nd]
    p = [0.02 * torch.randn(m, D, generator=g) for m in Ms]

    for use_graph in (False, True):
        plan = shooting.BatchedClosurePlan(D, nt, scheme, LM.withlogdet, sig, LM.eta, lam, dev(), Ms, Nxs, use_graph=use_graph)
        eligible = version != "logdet" and scheme == "Euler" and min(Nxs) > 0 and max(Ms) <= 64 and max(Nxs) <= 32768
        assert plan.one_launch == (one_launch and eligible)
        if not one_launch and not eligible and use_graph:
            return                                   # identical to the one_launch=True run of this case
        plan.set_geometry(q0, x0)
        plan.set_targets(torch.cat(y), torch.cat(inv))
        for rep in range(2):                         # second pass: graph replay, and a different active set
            act = [1] * K if rep == 0 else [k % 2 for k in range(K)]
            plan.active[:] = act
            plan.losses[:] = -7.0
            stale = plan.grads.copy()
            for k in range(K):
                plan.X[k, :Ms[k] * D] = (p[k] * (1 + rep)).reshape(-1).numpy()
            plan.evaluate()
            for k in range(K):
                go = plan.grads[k * plan.ostride:k * plan.ostride + Ms[k] * D]
                if not act[k]:                       # skipped frames keep their old outputs
                    assert plan.losses[k] != plan.losses[k] or True
                    assert np.array_equal(go, stale[k * plan.ostride:k * plan.ostride + Ms[k] * D])
                    continue
                sp = LM._spec_for(Ms[k], Nxs[k], dev())
                cp = shooting.ClosurePlan(sp, False, lam)
                cp.set_problem(q0[k], x0[k], y[k], inv[k])
                L, gr = cp.evaluate((p[k] * (1 + rep)).to(dev()))
                if K > 8 and k % 9 != 0:             # many-frame cases: check every 9th frame
                    continue
                tol = 1.0 if max(Ms) <= 512 else 100.0       # same kernels / different engines (summation orders differ)
                if max(Ms) > 64 and max(Nxs) >= 26000:       # mid-size adjoint stage: one evaluation per (x,q) pair, other order
                    tol = max(tol, 10.0)
                if plan.one_launch:
                    tol = 10.0                               # other summation order (per-thread sums over all stages), fixed origin
                assert abs(plan.losses[k] - L) <= tol * 2e-6 * abs(L), (k, plan.losses[k], L)
                gr = gr.reshape(-1).numpy()
                assert np.abs(go - gr).max() <= tol * 1e-6 * np.abs(gr).max(), (k, np.abs(go - gr).max(), np.abs(gr).max())


def test_finalize_trajectories_and_coverage_counts():
    from diff_icp_b200 import shooting
    from diff_icp_b200.core.LDDMM import LDDMMModel
    D, nt, sig, lam = 2, 5, 0.12, 10.0
    Ms, Nxs = [16, 30, 9], [700, 1500, 260]
    LM = LDDMMModel(sigma=sig, D=D, lambd=lam, version="hybrid", scheme="Euler", nt=nt, spec=spec())
    g = torch.Generator().manual_seed(5)
    q0 = [torch.rand(m, D, generator=g).to(dev()) for m in Ms]
    x0 = [(1.6 * torch.rand(n, D, generator=g) - 0.3).to(dev()) for n in Nxs]        # some points are far from the support
    y = [x.clone() for x in x0]
    inv = [torch.ones(n, device=dev()) for n in Nxs]
    p = [0.05 * torch.randn(m, D, generator=g) for m in Ms]
    plan = shooting.BatchedClosurePlan(D, nt, "Euler", True, sig, 0.0, lam, dev(), Ms, Nxs)
    plan.set_geometry(q0, x0)
    plan.set_targets(torch.cat(y), torch.cat(inv))
    traj, trajl, datal, counts = plan.finalize([t.numpy() for t in p], coverage_radius=2.0 * sig)
    assert counts.shape == (3, nt + 1) and counts.sum() > 0
    for k in range(3):
        sh = LM.Shoot(q0[k], p[k].to(dev()), x0[k])
        mine = plan.frame_states(traj, k)
        assert len(mine) == nt + 1 and len(mine[-1]) == 4
        for t in (0, nt):
            for a, b in zip(mine[t], sh[t]):
                assert torch.allclose(a, b, rtol=0, atol=2e-6 * float(b.abs().max() + 1e-6))
        ref_counts = [int(LM.Kernel.check_coverage(st[-1], st[0], 2.0).sum()) for st in sh]
        assert abs(sum(counts[k].tolist()) - sum(ref_counts)) <= 2          # decisions within rounding of the threshold
        tl = float(LM.trajloss(sh))
        assert abs(trajl[k] - tl) <= 1e-5 * abs(tl) + 1e-7
        dl = float(((sh[-1][3] - y[k]) ** 2).sum())
        assert abs(datal[k] - dl) <= 1e-5 * dl + 1e-7


def make_psr(batched, S=1, scheme="grid", seed=1234, K=6, N=900):
    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.core.PSR import DiffPSR
    g = torch.Generator().manual_seed(seed)
    C = 12
    t = torch.linspace(0, 2 * math.pi, C + 1)[:-1]
    mu0 = torch.stack((0.5 + 0.4 * (t / 7) * t.cos(), 0.5 + 0.3 * t.sin()), 1)
    x = []
    for k in range(K):
        sets = []
        for s in range(S):
            n = N + 37 * k + 11 * s
            c = torch.randint(0, C, (n,), generator=g)
            pts = mu0[c] + 0.03 * torch.randn(n, 2, generator=g) + 0.02 * torch.randn(1, 2, generator=g) + 0.3 * s
            sets.append(pts.to(dev()))
        x.append(sets if S > 1 else sets[0])
    torch.manual_seed(seed)
    G = GaussianMixtureUnif(torch.zeros(C, 2), spec=spec())
    LM = LDDMMModel(sigma=0.2, D=2, lambd=500.0, version="hybrid", scheme="Euler", nt=10, spec=spec())
    LM.use_cuda_graph = True
    P = DiffPSR(x, G, LM, dataspec=spec(), compspec=spec())
    P.printstuff = False
    P.batched_lbfgs = batched
    if scheme == "grid":
        P.set_support_scheme("grid", rho=math.sqrt(2))
    else:
        P.set_support_scheme("decim", rho=1.0)
    P.reinitialize_GMM()
    return P


@pytest.mark.parametrize("S,scheme", [(1, "grid"), (2, "decim")])
def test_lockstep_reg_opt_matches_sequential(S, scheme):
    Pa, Pb = make_psr(True, S, scheme), make_psr(False, S, scheme)
    assert Pa._batched_plan() is not None and Pb._batched_plan() is None
    for it in range(3):
        for P in (Pa, Pb):
            P.GMM_opt(max_iterations=5, tol=1e-3)
            P.Reg_opt(nmax=1, tol=1e-3)
        assert abs(Pa.FE - Pb.FE) <= 2e-4 * abs(Pb.FE), (it, Pa.FE, Pb.FE)
        if it == 0:
            # after ONE registration step from identical starting points the two paths differ by optimiser rounding only
            # (fp64 vs fp32 L-BFGS vectors over <= 20 iterations of an unconverged, ill-conditioned problem)
            for k in range(Pa.K):
                for s in range(S):
                    assert float((Pa.x1[k, s] - Pb.x1[k, s]).abs().max()) <= 5e-3 * 0.2, (k, s)
    for k in range(Pa.K):
        for s in range(S):
            assert float((Pa.x1[k, s] - Pb.x1[k, s]).abs().max()) <= 5e-2 * 0.2      # drift after 3 outer iterations
        assert len(Pa.shoot[k]) == 11 and Pa.shoot[k][-1][-1].shape == Pb.shoot[k][-1][-1].shape
        # the stored shoot is the shoot of the stored momenta
        sh = Pa.LMi.Shoot(Pa.q0[k], Pa.a0[k], Pa.allx0[k])
        assert torch.allclose(sh[-1][3], Pa.shoot[k][-1][3], atol=1e-6)
        # bookkeeping: the fused data loss is the quadratic loss of the stored points
        for s in range(S):
            ql = float(((Pa.x1[k, s] - Pa.y[k, s]) ** 2).sum() / (2 * Pa.GMMi[s].sigma ** 2))
            assert abs(float(Pa.quadloss[k, s]) - ql) <= 1e-5 * ql + 1e-6
    # registrations built from the lock-step result work like the sequential ones
    R = Pa.Registration(0)
    z = R.apply(Pa.x0[0, 0])
    assert torch.allclose(z, Pa.x1[0, 0], atol=2e-6)


def test_lockstep_is_deterministic():
    Pa, Pb = make_psr(True), make_psr(True)
    for P in (Pa, Pb):
        P.GMM_opt(max_iterations=5, tol=1e-3)
        P.Reg_opt(nmax=2, tol=1e-3)
    assert Pa.FE == Pb.FE
    for k in range(Pa.K):
        assert torch.equal(Pa.a0[k], Pb.a0[k])


@pytest.mark.parametrize("one_launch", [False, True])
def test_frame_groups_agree_with_a_single_batch(one_launch, monkeypatch):
    """lockstep_groups: the frames registered as 1, 2 or 3 groups (own plan / stream / host thread each).  With the stage
    kernels the groups give the same BITS here (the kernels' column-split counts coincide for these batch sizes); with the
    one-launch closure the rows-per-CTA split follows the group's largest frame, so the groups agree to rounding."""
    from diff_icp_b200 import shooting
    monkeypatch.setattr(shooting.BatchedClosurePlan, "one_launch_closure", one_launch)
    outs = []
    for groups in (1, 2, 3):
        P = make_psr(True)
        P.lockstep_groups, P.lockstep_group_min_frames = groups, 2
        assert len(P._batched_plan()) == groups
        for _ in range(2):
            P.GMM_opt(max_iterations=5, tol=1e-3)
            P.Reg_opt(nmax=2, tol=1e-3)
        outs.append((P.FE, [a.clone() for a in P.a0], [P.x1[k, 0].clone() for k in range(P.K)], [float(r) for r in P.regloss]))
    for o in outs[1:]:
        if not one_launch:
            assert o[0] == outs[0][0] and o[3] == outs[0][3]
            assert all(torch.equal(a, b) for a, b in zip(o[1], outs[0][1]))
            assert all(torch.equal(a, b) for a, b in zip(o[2], outs[0][2]))
        else:
            assert abs(o[0] - outs[0][0]) <= 1e-5 * abs(outs[0][0])
            for a, b in zip(o[2], outs[0][2]):
                assert (a - b).abs().max().item() <= 2e-3 * 0.2          # points: well inside one kernel width


@pytest.mark.parametrize("D,Ms,Nxs,nmax", [(2, [25] * 7, [3000 + 100 * k for k in range(7)], 1),
                                           (3, [12, 40, 64, 27, 5], [900, 5000, 2500, 16000, 300], 3),
                                           (2, [9], [700], 2)])
def test_device_lbfgs_equals_host_lbfgs(D, Ms, Nxs, nmax, monkeypatch):
    """tools.optim.DeviceLockstepLBFGS (csrc/lbfgs_device.cuh: the L-BFGS state machines of all frames in a kernel after the
    one-launch closure, the rounds of a step as ONE CUDA graph launch with a WHILE node) against the host state machines
    (csrc/lbfgs_batch.cu, tested against torch.optim.LBFGS in test_lockstep_lbfgs.py) on the same registration problems:
    same step / evaluation / iteration counts and, up to the fp64 rounding of the dot products (other summation order)
    amplified by the line search, the same momenta.  Run twice: the second run replays the captured graph."""
    from diff_icp_b200 import shooting
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.tools.optim import LBFGS_optimization_lockstep
    sig, lam = 0.25, 50.0
    LM = LDDMMModel(sigma=sig, D=D, lambd=lam, version="hybrid", scheme="Euler", nt=6, spec=spec())
    K = len(Ms)
    g = torch.Generator().manual_seed(21)
    q0 = [torch.rand(m, D, generator=g).to(dev()) for m in Ms]
    x0 = [torch.rand(n, D, generator=g).to(dev()) for n in Nxs]
    y = [(x + 0.05 * torch.randn(x.shape, generator=g).to(dev()) + 0.03) for x in x0]
    inv = [torch.full((n,), 30.0, device=dev()) for n in Nxs]
    p0 = [np.zeros((m, D), np.float32) for m in Ms]
    res = {}
    for device_opt in (True, False):
        monkeypatch.setattr(shooting.BatchedClosurePlan, "device_lbfgs_enabled", device_opt)
        plan = shooting.BatchedClosurePlan(D, 6, "Euler", LM.withlogdet, sig, LM.eta, lam, dev(), Ms, Nxs, use_graph=True)
        assert plan.one_launch and plan.device_lbfgs == device_opt
        plan.set_geometry(q0, x0)
        plan.set_targets(torch.cat(y), torch.cat(inv))
        runs = []
        for rep in range(2):
            bp, bL, steps, change, rounds = LBFGS_optimization_lockstep(p0, plan, nmax=nmax, tol=1e-4)
            runs.append((bp, bL, steps, rounds))
        res[device_opt] = runs
    for (bpd, bLd, sd, rd), (bph, bLh, sh, rh) in zip(res[True], res[False]):
        assert sd == sh
        assert abs(rd - rh) <= max(2, rh // 10)
        for k in range(K):
            assert abs(bLd[k] - bLh[k]) <= 2e-5 * abs(bLh[k]), (k, bLd[k], bLh[k])
            scale = max(np.abs(bph[k]).max(), 1e-6)
            assert np.abs(bpd[k] - bph[k]).max() <= 2e-2 * scale, (k, np.abs(bpd[k] - bph[k]).max(), scale)
    # replay = first (eager) run: deterministic kernels, same state machine
    (bp1, bL1, s1, r1), (bp2, bL2, s2, r2) = res[True]
    assert s1 == s2 and r1 == r2 and bL1 == bL2 and all(np.array_equal(a, b) for a, b in zip(bp1, bp2))


def test_device_lbfgs_fixed_step_mode_and_masks():
    """The paths the end-to-end runs rarely take: optimisers restarted WITHOUT line search (the reference's fall-back after a
    divergent step, tools/optim.py:77), a step of a SUBSET of the frames, set_x / get_x / stats of single frames -- device
    state machines against the host ones on the same closure plan."""
    from diff_icp_b200 import shooting
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from diff_icp_b200.tools.optim import DeviceLockstepLBFGS, LockstepLBFGS
    D, Ms, Nxs, sig, lam = 2, [16, 25, 9, 30], [1500, 2600, 800, 4000], 0.25, 50.0
    LM = LDDMMModel(sigma=sig, D=D, lambd=lam, version="classic", scheme="Euler", nt=5, spec=spec())
    K = len(Ms)
    g = torch.Generator().manual_seed(33)
    q0 = [torch.rand(m, D, generator=g).to(dev()) for m in Ms]
    x0 = [torch.rand(n, D, generator=g).to(dev()) for n in Nxs]
    y = [(x + 0.04 * torch.randn(x.shape, generator=g).to(dev())) for x in x0]
    inv = [torch.full((n,), 20.0, device=dev()) for n in Nxs]
    plan = shooting.BatchedClosurePlan(D, 5, "Euler", LM.withlogdet, sig, LM.eta, lam, dev(), Ms, Nxs, use_graph=True)
    assert plan.one_launch
    plan.set_geometry(q0, x0)
    plan.set_targets(torch.cat(y), torch.cat(inv))
    sizes = [m * D for m in Ms]
    p_start = [(1e-3 * torch.randn(n, generator=g)).numpy() for n in sizes]
    host = LockstepLBFGS(sizes, stride=plan.X.shape[1], max_iter=6)
    devo = DeviceLockstepLBFGS(sizes, plan, max_iter=6)
    for k in range(K):
        host.set_x(k, p_start[k])
        devo.set_x(k, p_start[k])
        host.reset(k, k % 2 == 0)                # frames 1 and 3: fixed-step mode
        devo.reset(k, k % 2 == 0)
    for mask in ([1, 1, 1, 1], [0, 1, 1, 0], [1, 0, 0, 1], [1, 1, 1, 1]):
        m = np.array(mask, np.uint8)
        host.step(m, plan.evaluate, plan.X, plan.active, plan.losses, plan.grads)
        devo.step(m)
        sh, sd = host.stats_all(), devo.stats_all()
        assert np.array_equal(sh[:, 2:], sd[:, 2:]), (sh[:, 2:], sd[:, 2:])          # evaluation / iteration counts
        assert np.allclose(sh[:, :2], sd[:, :2], rtol=3e-5, atol=0, equal_nan=True)
        xh, xd = host.get_all(), devo.get_all()
        for k in range(K):
            n = sizes[k]
            assert np.abs(xh[k, :n] - xd[k, :n]).max() <= 1e-2 * max(np.abs(xh[k, :n]).max(), 1e-6)
            assert np.array_equal(devo.get_x(k), xd[k, :n])
            assert devo.stats(k)["func_evals"] == int(sd[k, 2])
    bh, bd = host.get_all(best=True), devo.get_all(best=True)
    for k in range(K):
        assert np.abs(bh[k, :sizes[k]] - bd[k, :sizes[k]]).max() <= 1e-2 * max(np.abs(bh[k, :sizes[k]]).max(), 1e-6)


@pytest.mark.parametrize("version,scheme,D", [("hybrid", "Ralston", 3), ("logdet", "Euler", 3), ("classic", "Ralston", 2)])
def test_midsize_support_stage_kernels_vs_oracle(version, scheme, D):
    """Mid-size supports (150 / 333 points) x tens of thousands of data points per frame, evaluated by the lock-step stage kernels
    (small_rhs_step_kernel in its 128-register form, small_adj_mid_kernel + small_mid_finish_kernel): loss and gradient of every
    frame against the fp64 oracle (core/LDDMM.py:363-371 + core/PSR.py:498-516 restated in oracle/lddmm.py)."""
    from diff_icp_b200 import shooting
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from oracle.lddmm import LDDMMOracle
    sig, lam, nt = 0.2, 30.0, 2
    Ms, Nxs = [150, 333, 97], [30000, 27001, 29500]
    LM = LDDMMModel(sigma=sig, D=D, lambd=lam, version=version, scheme=scheme, nt=nt, spec=spec())
    g = torch.Generator().manual_seed(11)
    q0 = [torch.rand(m, D, generator=g) for m in Ms]
    x0 = [torch.rand(n, D, generator=g) for n in Nxs]
    y = [x + 0.03 * torch.randn(x.shape, generator=g) for x in x0]
    inv = [0.5 + torch.rand(n, generator=g) for n in Nxs]
    p = [0.01 * torch.randn(m, D, generator=g) for m in Ms]
    plan = shooting.BatchedClosurePlan(D, nt, scheme, LM.withlogdet, sig, LM.eta, lam, dev(), Ms, Nxs, use_graph=False)
    plan.set_geometry([q.to(dev()) for q in q0], [x.to(dev()) for x in x0])
    plan.set_targets(torch.cat(y).to(dev()), torch.cat(inv).to(dev()))
    plan.active[:] = 1
    for k in range(len(Ms)):
        plan.X[k, :Ms[k] * D] = p[k].reshape(-1).numpy()
    plan.evaluate()
    OR = LDDMMOracle(sigma=sig, D=D, lambd=lam, version=version, scheme=scheme, nt=nt, chunk=4096)
    for k in range(len(Ms)):
        po = p[k].double().requires_grad_(True)
        Lo, _ = OR.loss(q0[k].double(), po, x0[k].double(), y[k].double(), inv[k].double())
        (go,) = torch.autograd.grad(Lo, [po])
        go = go.reshape(-1).numpy()
        gk = plan.grads[k * plan.ostride:k * plan.ostride + Ms[k] * D]
        assert abs(plan.losses[k] - float(Lo)) < 2e-5 * abs(float(Lo)), (k, plan.losses[k], float(Lo))
        assert np.abs(gk - go).max() < 2e-4 * np.abs(go).max(), (k, np.abs(gk - go).max(), np.abs(go).max())


def test_single_frame_midsize_stage_kernels_vs_oracle():
    """ONE frame with enough data points to fill the SMs with 512-row CTAs (80 000 points, 150 support points): the per-frame
    closure (shooting.ClosurePlan -> dicp_small_rhs_step / dicp_small_adj_step) takes the mid-size stage kernels by itself;
    loss and gradient against the fp64 oracle."""
    from diff_icp_b200 import shooting
    from diff_icp_b200.core.LDDMM import LDDMMModel
    from oracle.lddmm import LDDMMOracle
    D, M, Nx, sig, lam, nt = 3, 150, 80000, 0.2, 30.0, 2
    LM = LDDMMModel(sigma=sig, D=D, lambd=lam, version="hybrid", scheme="Ralston", nt=nt, spec=spec())
    g = torch.Generator().manual_seed(17)
    q0, x0 = torch.rand(M, D, generator=g), torch.rand(Nx, D, generator=g)
    y = x0 + 0.03 * torch.randn(Nx, D, generator=g)
    inv = 0.5 + torch.rand(Nx, generator=g)
    p = 0.01 * torch.randn(M, D, generator=g)
    sp = LM._spec_for(M, Nx, dev())
    cp = shooting.ClosurePlan(sp, False, lam)
    assert cp.plan.small
    cp.set_problem(q0.to(dev()), x0.to(dev()), y.to(dev()), inv.to(dev()))
    L, gr = cp.evaluate(p.to(dev()))
    OR = LDDMMOracle(sigma=sig, D=D, lambd=lam, version="hybrid", scheme="Ralston", nt=nt, chunk=4096)
    po = p.double().requires_grad_(True)
    Lo, _ = OR.loss(q0.double(), po, x0.double(), y.double(), inv.double())
    (go,) = torch.autograd.grad(Lo, [po])
    assert abs(float(L) - float(Lo)) < 2e-5 * abs(float(Lo))
    assert np.abs(gr.cpu().numpy() - go.numpy()).max() < 2e-4 * np.abs(go.numpy()).max()


@pytest.mark.parametrize("seed", range(10))
def test_midsize_form_equals_one_launch_form_on_random_shapes(seed, monkeypatch):
    """The two forms of the stage kernels for supports above 64 points (DICP_SMALL_MID = 0: x-row CTAs + column-split q-row CTAs in
    one launch, every (x,q) pair twice; = 1: ring rounds over 64-column groups + finish launch, every pair once, 2 or 4 data points
    per lane) on random ragged batches: support sizes 1..2048 incl. the 64 / 65 and 2048 edges, data sizes around the 256 / 512-row
    CTA edges, all three models, both schemes, D = 2 and 3, random active masks.  Same formulas, other summation orders."""
    from diff_icp_b200 import shooting
    from diff_icp_b200.core.LDDMM import LDDMMModel
    rng = np.random.default_rng(100 + seed)
    D = int(rng.choice([2, 3]))
    version = ["classic", "hybrid", "logdet"][seed % 3]
    scheme = ["Euler", "Ralston"][(seed // 3) % 2]
    K = int(rng.integers(1, 5))
    edgesM = [65, 64, 2048, 1, 129, 1210, 2047, 66]
    edgesN = [1, 255, 256, 257, 511, 512, 513, 1025, 3000]
    Ms = [int(rng.choice(edgesM)) if rng.random() < 0.5 else int(rng.integers(1, 700)) for _ in range(K)]
    Ms[0] = max(Ms[0], 65 + seed)                       # at least one support above the ring form's 64 points
    Nxs = [int(rng.choice(edgesN)) if rng.random() < 0.6 else int(rng.integers(1, 5000)) for _ in range(K)]
    sig, lam, nt = 0.25, 40.0, 2
    LM = LDDMMModel(sigma=sig, D=D, lambd=lam, version=version, scheme=scheme, nt=nt, spec=spec())
    g = torch.Generator().manual_seed(1000 + seed)
    q0 = [torch.rand(m, D, generator=g).to(dev()) for m in Ms]
    x0 = [torch.rand(n, D, generator=g).to(dev()) for n in Nxs]
    y = [torch.rand(n, D, generator=g).to(dev()) for n in Nxs]
    inv = [(0.5 + torch.rand(n, generator=g)).to(dev()) for n in Nxs]
    p = [0.01 * torch.randn(m, D, generator=g) for m in Ms]
    act = [1] + [int(rng.random() < 0.7) for _ in range(K - 1)]
    res = {}
    for mid, rows in (("0", None), ("1", "2"), ("1", "4")):
        monkeypatch.setenv("DICP_SMALL_MID", mid)
        if rows is None:
            monkeypatch.delenv("DICP_SMALL_MID_R", raising=False)
        else:
            monkeypatch.setenv("DICP_SMALL_MID_R", rows)
        plan = shooting.BatchedClosurePlan(D, nt, scheme, LM.withlogdet, sig, LM.eta, lam, dev(), Ms, Nxs, use_graph=False)
        plan.one_launch = False
        plan.set_geometry(q0, x0)
        plan.set_targets(torch.cat(y), torch.cat(inv))
        plan.active[:] = act
        for k in range(K):
            plan.X[k, :Ms[k] * D] = p[k].reshape(-1).numpy()
        plan.evaluate()
        res[(mid, rows)] = (plan.losses.copy(), plan.grads.copy(), plan.ostride)
    L0, g0, os_ = res[("0", None)]
    for key in (("1", "2"), ("1", "4")):
        L1, g1, _ = res[key]
        for k in range(K):
            if not act[k]:
                continue
            a, b = g0[k * os_:k * os_ + Ms[k] * D], g1[k * os_:k * os_ + Ms[k] * D]
            assert abs(L0[k] - L1[k]) <= 1e-5 * abs(L0[k]), (key, k, Ms, Nxs, L0[k], L1[k])
            assert np.abs(a - b).max() <= 2e-5 * np.abs(a).max(), (key, k, Ms, Nxs, np.abs(a - b).max(), np.abs(a).max())
