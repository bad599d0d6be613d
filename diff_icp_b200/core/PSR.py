"""Outer alternation of diffICP (GMM fit <-> LDDMM registration) on the B200: drop-in for the reference's
``diffICP/core/PSR.py`` classes ``MultiPSR`` and ``DiffPSR`` (core/PSR.py:42-569).

Same storage (object arrays x0 / x1 / y indexed [frame k, structure s]; lists a0[k], q0[k], shoot[k]; GMMi[s]) and the
same public methods.  The arithmetic runs in the CUDA kernels through ``GaussianMixtureUnif`` and ``LDDMMModel``.

Groupwise multi-GPU mode: construct one DiffPSR per rank on that rank's frames and pass ``comm`` (a
``diff_icp_b200.dist.StatsComm``).  Registration of a frame uses only that frame's data (core/PSR.py:528), so Reg_opt is
purely local; the GMM fit couples frames through O(C) sufficient statistics that the GMM all-reduces (the only
collective); the free energy and the re-initialisation statistics are summed over ranks here.
"""

from __future__ import annotations

import copy
import warnings

import numpy as np
import torch

from .GMM import GaussianMixtureUnif
from .LDDMM import LDDMMModel, ShootResult
from .registrations import LDDMMRegistration
from ..tools.in_out import read_point_sets
from ..tools.point_sets import decimate
from ..tools.spec import defspec


def get_bounds(*xlist, relmargin=0.2, comm=None):
    """Per-dimension (min, max) over several point sets, enlarged by a relative margin
    (reference: visualization/visu.py:35-50, 2-D there; any D here).  With `comm`: over the point sets of all ranks."""
    mins = torch.stack([x.min(0).values for x in xlist if len(x) > 0]).min(0).values
    maxs = torch.stack([x.max(0).values for x in xlist if len(x) > 0]).max(0).values
    if comm is not None:
        both = comm.max(torch.cat((-mins, maxs)))
        mins, maxs = -both[:mins.numel()], both[mins.numel():]
    mins, maxs = mins.cpu().numpy(), maxs.cpu().numpy()
    return (1 + relmargin) * mins - relmargin * maxs, (1 + relmargin) * maxs - relmargin * mins


class MultiPSR:
    """Base class: multiple point-set registration to per-structure GMMs (reference: core/PSR.py:42-345)."""

    def __init__(self, x, GMMi, dataspec=defspec, compspec=defspec, comm=None):
        self.dataspec, self.compspec = dataspec, compspec
        self.comm = comm
        self.printstuff = True
        x, self.K, self.S, self.D = read_point_sets(x)

        self.x0 = np.empty((self.K, self.S), dtype=object)     # unregistered point sets
        self.x1 = np.empty((self.K, self.S), dtype=object)     # registered (warped) point sets
        self.y = np.empty((self.K, self.S), dtype=object)      # quadratic targets from the GMM E step
        for k in range(self.K):
            for s in range(self.S):
                self.x0[k, s] = x[k][s].contiguous().detach().to(**self.dataspec)
                self.x1[k, s] = self.x0[k, s].clone()
                self.y[k, s] = self.x0[k, s].clone()
        self.N = np.array([[self.x0[k, s].shape[0] for s in range(self.S)] for k in range(self.K)]).reshape(self.K, self.S)

        if isinstance(GMMi, GaussianMixtureUnif):
            self.GMMi = [copy.deepcopy(GMMi) for _ in range(self.S)]
        else:
            if not isinstance(GMMi, list) or len(GMMi) != self.S:
                raise ValueError("GMMi should be a single GMM model, or a list with S GMM models")
            self.GMMi = [copy.deepcopy(gmm) for gmm in GMMi]
        if any(gmm.spec != compspec for gmm in self.GMMi):
            raise ValueError("Spec (dtype+device) error : GMM 'spec' and multiPSR 'compspec' attributes should be the same")
        for gmm in self.GMMi:
            gmm.comm = comm
        if comm is not None:
            self._broadcast_GMMs()

        # full EM free energy  F = sum_{k,s} quadloss[k,s] + sum_k regloss[k] + sum_s Cfe[s]   (core/PSR.py:114-121)
        self.Cfe = [None] * self.S
        self.regloss = [0] * self.K
        self.quadloss = torch.zeros(self.K, self.S, **self.compspec)
        self.FE = None
        self.update_GMM_targets()
        self.shoot = [None] * self.K

    def _broadcast_GMMs(self):
        """Multi-GPU: the merged column statistics are centred on mu_old (mu' = mu_old + B/S0), which is only valid if every
        rank holds bit-identical GMM parameters.  Initial models built from a rank's own frame (init_components = ("set", k)
        or {"set", "C"}) or from random centroids differ per rank: rank 0's parameters are made everyone's."""
        for gmm in self.GMMi:
            cc = torch.tensor([float(gmm.C), -float(gmm.C)], dtype=torch.float64, device=gmm.mu.device)
            cc = self.comm.max(cc)
            if cc[0] != -cc[1]:
                raise ValueError("multi-GPU mode: every rank must build its initial GMM with the same number of components "
                                 f"(this rank: {gmm.C}, over ranks: {int(-cc[1])}..{int(cc[0])}); pass an explicit model list")
            head = torch.tensor([float(gmm.sigma), float(gmm.outliers["eta0"]) if gmm.outliers is not None else 0.0,
                                 float(gmm.outliers["vol0"]) if gmm.outliers is not None and gmm.outliers["vol0"] is not None
                                 else float("nan")], dtype=torch.float64, device=gmm.mu.device)
            head = self.comm.broadcast(head)
            gmm.mu = self.comm.broadcast(gmm.mu.contiguous())
            gmm.w = self.comm.broadcast(gmm.w.contiguous())
            gmm.sigma = float(head[0])
            if gmm.outliers is not None:
                gmm.outliers["eta0"] = float(head[1])
                gmm.outliers["vol0"] = None if bool(torch.isnan(head[2])) else float(head[2])

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.dataspec = defspec
        self.compspec = defspec

    def __getstate__(self):
        state = dict(self.__dict__)
        state["comm"] = None
        state.pop("_bplan", None)          # device buffers + captured CUDA graph of the batched registration
        state.pop("_bplan_key", None)
        state.pop("_a0_host", None)
        return state

    # ------------------------------------------------------------------------------------------------------
    def _cat_structure(self, arr, s):
        return torch.cat(tuple(arr[:, s]), dim=0).to(**self.compspec) if self.K > 0 else torch.empty(0, self.D, **self.compspec)

    def reinitialize_GMM(self, s=None, do_mu=True, do_sigma=True):
        """Centroids near the centre of mass of all unwarped points, sigma = std/4 (reference: core/PSR.py:143-167).
        With `comm`, mean and std are those of the points of ALL ranks, and the random draw is made identical on every
        rank by drawing on rank 0's generator state broadcast through `comm`."""
        for s in (range(self.S) if s is None else [s]):
            allx0s = self._cat_structure(self.x0, s)
            if self.comm is None:
                mean, std = allx0s.mean(dim=0), allx0s.std().item()
            else:
                # global mean / std (over all coordinates, unbiased like torch's .std()) from per-rank sums in fp64
                a = allx0s.double()
                red = self.comm.sum(torch.cat((torch.tensor([float(a.numel())], dtype=torch.float64, device=a.device),
                                               a.sum(0), (a ** 2).sum().reshape(1))))
                n, sm, sq = float(red[0]), red[1:1 + self.D], float(red[1 + self.D])
                mean = (sm / (n / self.D)).to(**self.compspec)
                tot = float(sm.sum())
                std = max((sq - tot * tot / n) / (n - 1), 0.0) ** 0.5
            gmm = self.GMMi[s]
            if do_mu and gmm.to_optimize["mu"]:
                noise = torch.randn(gmm.C, self.D, **self.dataspec)
                if self.comm is not None:
                    noise = self.comm.broadcast(noise)
                gmm.mu = (mean + 0.05 * std * noise).to(**self.compspec)
            if do_sigma and gmm.to_optimize["sigma"]:
                gmm.sigma = 0.25 * std
        self.update_GMM_targets()

    # accessors (reference: core/PSR.py:174-190)
    def get_data_points(self, k=0, s=0):
        return self.x0[k, s]

    def get_warped_data_points(self, k=0, s=0):
        return self.x1[k, s]

    def get_template(self, s=0):
        return self.GMMi[s].mu

    # ------------------------------------------------------------------------------------------------------
    def _scatter_targets(self, allys, s, allx1s=None):
        """y[k,s] <- the frame's slice of the targets of structure s; quadloss[:, s] refreshed.  allx1s: the concatenated
        warped points the targets were computed from (saves K small reductions: one segmented sum instead)."""
        last = 0
        for k in range(self.K):
            first, last = last, last + self.N[k, s]
            self.y[k, s] = allys[first:last].to(**self.dataspec)
        n = int(self.N[0, s]) if self.K > 0 else 0
        if allx1s is not None and n > 0 and all(int(self.N[k, s]) == n for k in range(self.K)):
            d2 = ((allx1s - allys) ** 2).view(self.K, n * self.D).sum(1)        # equal-sized frames: one reduction
            self.quadloss[:, s] = d2.to(**self.compspec) / (2 * self.GMMi[s].sigma ** 2)
        else:
            for k in range(self.K):
                self.update_quadloss(k, s)

    def update_GMM_targets(self):
        """Recompute y, Cfe, quadloss, FE without any GMM parameter update (reference: core/PSR.py:197-213)."""
        for s in range(self.S):
            allx1s = self._cat_structure(self.x1, s)
            allys, self.Cfe[s], _ = self.GMMi[s].EM_step(allx1s, skip_M=True)
            self._scatter_targets(allys, s, allx1s)
        self.update_FE()

    def update_quadloss(self, k, s):
        self.quadloss[k, s] = ((self.x1[k, s].to(**self.compspec) - self.y[k, s].to(**self.compspec)) ** 2).sum() \
            / (2 * self.GMMi[s].sigma ** 2)

    def update_FE(self, message=None):
        """F = sum Cfe + sum regloss + sum quadloss (reference: core/PSR.py:226-236); regloss and quadloss are per-frame
        quantities and are summed over ranks in multi-GPU mode, Cfe is already global."""
        local = float(sum(self.regloss)) + self.quadloss.sum().item()
        if self.comm is not None:
            local = float(self.comm.sum(torch.tensor([local], dtype=torch.float64, device=self.compspec["device"]))[0])
        FE = float(sum(float(c) for c in self.Cfe)) + local
        if self.printstuff and message is not None:
            print(message.ljust(70) + f"Total free energy = {FE:.8}")
        if self.FE is not None and FE > self.FE:
            print("WARNING: measured increase in free energy ! Should not happen.")
        self.FE = FE

    # ------------------------------------------------------------------------------------------------------
    def GMM_opt(self, max_iterations=100, tol=1e-5):
        """GMM part of the alternation, one structure at a time on the points of all frames
        (reference: core/PSR.py:242-271)."""
        for s in range(self.S):
            allx1s = self._cat_structure(self.x1, s)
            allys, self.Cfe[s], _, i = self.GMMi[s].EM_optimization(allx1s, max_iterations=max_iterations, tol=tol)
            self._scatter_targets(allys, s, allx1s)
            message = f"GMM optim (structure {s}) : {i} EM steps"
            if self.GMMi[s].outliers:
                p0 = 1 / (1 + np.exp(-self.GMMi[s].outliers["eta0"]))
                message += f", p_outlier={p0:.4}"
            else:
                message += "."
            self.update_FE(message=message)

    def Reg_opt(self, tol=1e-5):
        raise NotImplementedError("function Reg_opt must be written in derived classes.")

    def Registration(self, k=0):
        """Registration object of frame k (reference: core/PSR.py:294-304)."""
        if isinstance(self, DiffPSR):
            return LDDMMRegistration(self.LMi, self.q0[k], self.a0[k])
        raise NotImplementedError("only diffeomorphic registrations are part of the B200 hot path")


class DiffPSR(MultiPSR):
    """MultiPSR with diffeomorphic (LDDMM) registrations (reference: core/PSR.py:354-569)."""

    def __init__(self, x, GMMi, LMi: LDDMMModel, dataspec=defspec, compspec=defspec, comm=None):
        super().__init__(x, GMMi, dataspec=dataspec, compspec=compspec, comm=comm)
        if LMi.Kernel.spec != compspec:
            raise ValueError("Spec (dtype+device) error : LDDMMmodel kernel 'spec' and diffPSR 'compspec' attributes should be the same")
        self.LMi = LMi
        self.allx0 = [torch.cat(tuple(self.x0[k, :]), dim=0).to(**self.compspec).contiguous() for k in range(self.K)]
        self.support_scheme, self.rho = None, None
        self.q0 = self.allx0                       # dense scheme by default: support points = data points
        self.a0 = [None] * self.K
        self.initialize_a0()

    def initialize_a0(self, **v2p_args):
        """Momenta giving (approximately) zero initial speeds (reference: core/PSR.py:406-413)."""
        for k in range(self.K):
            v0 = torch.zeros(self.q0[k].shape, **self.compspec)
            self.a0[k] = self.LMi.v2p(self.q0[k], v0, **v2p_args)

    def update_a0(self, q0_prev, a0_prev=None, **v2p_args):
        """Project v(q0_prev, a0_prev) on the span of the new support points (reference: core/PSR.py:415-425)."""
        if a0_prev is None:
            a0_prev = self.a0
        for k in range(self.K):
            v0 = self.LMi.v(self.q0[k], q0_prev[k], a0_prev[k])
            self.a0[k] = self.LMi.v2p(self.q0[k], v0, **v2p_args)

    def set_support_scheme(self, scheme="decim", rho=1.0, xticks=None, yticks=None, q0=None, zticks=None):
        """Choose the LDDMM support points: "decim" (greedy covering of the data), "grid" (regular grid of step
        rho*sigma over the data bounds; the reference's grid is 2-D only, core/PSR.py:472-482 -- here any D) or
        "custom" (reference: core/PSR.py:430-493)."""
        self.rho = rho
        Rcover = rho * self.LMi.Kernel.sigma
        self.support_scheme = scheme
        q0_prev = self.q0

        if scheme == "decim":
            self.q0 = [None] * self.K
            for k in range(self.K):
                ids = [decimate(self.x0[k, s], Rcover)[0] for s in range(self.S)]
                nd = sum(len(i) for i in ids)
                if self.printstuff:
                    print(f"Decimation, frame {k} : {nd} support points ({nd / sum(self.N[k, :]):.0%} of original sets)")
                self.q0[k] = torch.cat(tuple(self.x0[k, s][ids[s]] for s in range(self.S)), dim=0).to(**self.compspec).contiguous()

        elif scheme == "grid":
            given = [xticks, yticks, zticks][:self.D]
            if any(t is None for t in given):
                lo, hi = get_bounds(*self.allx0, relmargin=0.1, comm=self.comm)
            ticks = [np.arange(lo[d] - Rcover / 2, hi[d] + Rcover / 2, Rcover) if given[d] is None else np.asarray(given[d])
                     for d in range(self.D)]
            if self.D == 2:     # same point ordering as the reference (meshgrid 'xy' + Fortran reshape)
                pts = np.stack(np.meshgrid(ticks[0], ticks[1]), axis=2).reshape((-1, 2), order="F")
            else:
                pts = np.stack(np.meshgrid(*ticks, indexing="ij"), axis=-1).reshape(-1, self.D)
            grid = torch.tensor(pts, **self.compspec).contiguous()
            self.q0 = [grid] * self.K

        elif scheme == "custom":
            assert q0 is not None, "For a custom support scheme, please specify argument q0"
            self.q0 = [q0.clone().detach().to(**self.compspec).contiguous()] * self.K

        else:
            raise ValueError(f"Unknown value of support point scheme : {scheme}. Only values available are 'decim', 'grid' and 'custom'.")

        self.update_a0(q0_prev, rcond=1e-1)

    def QuadLossFunctor(self, k):
        """x -> sum_n |x_n - y_n|^2 / (2 sigma_s(n)^2) over all points of frame k (reference: core/PSR.py:498-516)."""
        y = torch.cat(tuple(self.y[k, :]), dim=0).to(**self.compspec).contiguous()
        inv = torch.cat(tuple(torch.full((int(self.N[k, s]),), 1.0 / (2 * self.GMMi[s].sigma ** 2)) for s in range(self.S))
                        ).to(**self.compspec).contiguous()

        def dataloss_func(x):
            return (((x - y) ** 2) * inv[:, None]).sum()
        dataloss_func.targets, dataloss_func.inv2sig2 = y, inv       # lets LDDMMModel.Optimize fuse the whole closure
        return dataloss_func

    # number of frames registered concurrently on the frame-by-frame path (one Python thread + one CUDA stream each). Frames
    # are independent (core/PSR.py:528), and every frame's computation is deterministic, so results do not depend on this
    # setting (asserted bit for bit in the tests).  None = automatic.  MEASURED on B200: with 64 frames x 10k points and 25
    # support points it does not pay off (closures of microseconds: the per-frame L-BFGS is bound by host Python time, which
    # threads serialise on the GIL; that regime is served by the lock-step path anyway); with configs[3]-shaped frames (50k
    # points, 1210 support points, 3-D, Ralston: closures of ~3 ms of which ~45 % are latency-bound launches of a few
    # microseconds) a second frame in flight fills those gaps: Reg_opt of 16 frames 1.37 -> 1.13 s with 2 workers, 0.99-1.18 s
    # with 4.  Automatic = 3 workers when a frame's closure sweeps >= 2e7 (point, support point) pairs per stage, else 1.
    frame_workers = None
    frame_workers_big_pairs = 2.0e7

    def _auto_frame_workers(self):
        if self.frame_workers is not None:
            return int(self.frame_workers)
        if self.K < 2:
            return 1
        pairs = sorted(float(self.q0[k].shape[0]) * float(self.allx0[k].shape[0]) for k in range(self.K))
        return 3 if pairs[len(pairs) // 2] >= self.frame_workers_big_pairs else 1

    def _register_frame(self, k, nmax, tol):
        """Optimise a0[k] and collect everything Reg_opt's bookkeeping needs (no shared state is written here)."""
        x0 = None if self.support_scheme is None else self.allx0[k]
        a0, shoot, regloss, datal, isteps, change = \
            self.LMi.Optimize(self.QuadLossFunctor(k), self.q0[k], self.a0[k], x0, tol=tol, nmax=nmax)
        allx1k = shoot[-1][0] if x0 is None else shoot[-1][-1]
        counts = None
        if self.support_scheme is not None:
            # coverage of the warped data points by the support points at every time step (core/PSR.py:559-566);
            # one host read for the whole trajectory
            counts = torch.stack([self.LMi.Kernel.check_coverage(st[-1], st[0], 2.0).sum() for st in shoot]).tolist()
        return dict(a0=a0, shoot=shoot, regloss=regloss, datal=datal, isteps=isteps, change=change, x1=allx1k, counts=counts)

    # Lock-step registration of all frames: with a small support set (grid / decimated / custom
    # scheme) one frame's closure is microseconds of device work, so the K independent L-BFGS runs of Reg_opt advance
    # together and every round evaluates all pending closures with ONE launch sequence (shooting.BatchedClosurePlan +
    # tools.optim.LBFGS_optimization_lockstep).  Per frame the algorithm is unchanged; iterates agree with the
    # one-frame-at-a-time path up to floating-point rounding.  Set False for the sequential path.
    batched_lbfgs = True

    # Frame groups of the lock-step registration: the K frames can be split into `lockstep_groups` contiguous groups, each
    # with its own BatchedClosurePlan, CUDA stream, L-BFGS state machines and host thread.  Per frame the algorithm is
    # unchanged; values agree with the single batch up to the summation order of the kernels' row / column splits (chosen
    # from the group's sizes) and are bit-identical when those coincide (tests/test_gpu_batched.py).  MEASURED on B200
    # (64 frames x 10k points, 25 support points, Reg_opt of one iteration): with the L-BFGS state machines on the HOST and the
    # one-launch closure, a second group hides the host work of a round behind the other group's kernel (16.2 / 14.7 / 14.9 /
    # 16.2 ms with 1 / 2 / 3 / 4 groups); with the state machines on the DEVICE (the default where the one-launch closure
    # applies) there is no host work per round left to hide: 12.8 / 12.7 / 16.0 ms with 1 / 2 / 4 groups.
    # With MID-SIZE supports (more than 64 points: stage kernels of 0.2-1.3 ms) groups pay for another reason: every
    # stage launch of a group ends in a partly filled wave of long CTAs, and the last lock-step rounds have few active frames;
    # the other groups' launches fill those holes.  Reg_opt of the configs[3]-shaped atlas (1210 support points, 50k data points
    # per frame), 1 / 2 / 4 / 8 groups: 64 frames 3490 / 3230 / 3123 / 3349 ms, 16 frames 923 / 871 / 799 / 857 ms, 8 frames
    # 406 / 368 / 354 ms, 4 frames 205 / 184 ms.
    # Default None = 1 group for small supports, min(4, K // 2) groups for mid-size supports; an integer forces the count
    # (subject to lockstep_group_min_frames).
    lockstep_groups = None
    lockstep_group_min_frames = 8          # forced counts: below 2 x this many frames a second group is never formed

    def _frame_groups(self):
        if self.lockstep_groups is None:
            mid = max(int(q.shape[0]) for q in self.q0) > 64
            G = max(1, min(4, self.K // 2)) if mid else 1
        else:
            G = max(1, int(self.lockstep_groups))
            while G > 1 and self.K < G * self.lockstep_group_min_frames:
                G -= 1
        bounds = [round(g * self.K / G) for g in range(G + 1)]
        return [list(range(bounds[g], bounds[g + 1])) for g in range(G)]

    def _batched_plan(self):
        """The BatchedClosurePlans (one per frame group) of the current frames / support points, or None when the lock-step
        path does not apply (dense or large supports, CPU tensors, generic model options)."""
        from .. import ops, shooting
        LM = self.LMi
        dev = torch.device(self.compspec["device"])
        if not (self.batched_lbfgs and LM.fused_closure and dev.type == "cuda" and self.K >= 1):
            return None
        has_x = self.support_scheme is not None
        if LM.withlogdet and LM.gradcomponent and LM.try_trajcost_optim and not has_x:
            return None
        Ms = [int(q.shape[0]) for q in self.q0]
        if min(Ms) < 1 or not ops.use_small_path(max(Ms), batched=True):
            return None
        Nxs = [int(x.shape[0]) if has_x else 0 for x in self.allx0]
        if has_x and min(Nxs) < 1:                 # a frame without data points: leave it to the per-frame path
            return None
        groups = self._frame_groups()
        key = (tuple(Ms), tuple(Nxs), LM.D, LM.nt, LM.scheme, LM.withlogdet, float(LM.Kernel.sigma), float(LM.eta),
               float(LM.lam), str(dev), bool(LM.use_cuda_graph), len(groups),
               tuple((q.data_ptr(), q._version) for q in self.q0), tuple((x.data_ptr(), x._version) for x in self.allx0))
        if getattr(self, "_bplan_key", None) != key:
            plans = []
            for frames in groups:
                plan = shooting.BatchedClosurePlan(LM.D, LM.nt, LM.scheme, LM.withlogdet, LM.Kernel.sigma, LM.eta, LM.lam, dev,
                                                   [Ms[k] for k in frames], [Nxs[k] for k in frames], use_graph=LM.use_cuda_graph)
                plan.set_geometry([self.q0[k] for k in frames], [self.allx0[k] if has_x else None for k in frames])
                plan.frames = frames
                plan.stream = torch.cuda.Stream(dev) if len(groups) > 1 else None
                plans.append(plan)
            self._bplan, self._bplan_key = plans, key
        return self._bplan

    def _register_group_lockstep(self, plan, nmax, tol):
        """Lock-step registration of the frames of one group (runs on the group's stream / thread)."""
        from ..tools.optim import LBFGS_optimization_lockstep
        frames, D = plan.frames, self.D
        # targets and weights of the group's data points, concatenated in frame order (QuadLossFunctor, core/PSR.py:498-516)
        y_cat = torch.cat([self.y[k, s] for k in frames for s in range(self.S)], dim=0).to(**self.compspec)
        table = torch.tensor([1.0 / (2 * self.GMMi[s].sigma ** 2) for s in range(self.S)], **self.compspec)
        if getattr(plan, "struct_id", None) is None:
            plan.struct_id = torch.cat([torch.full((int(self.N[k, s]),), s, dtype=torch.long)
                                        for k in frames for s in range(self.S)]).to(table.device)
        plan.set_targets(y_cat, table[plan.struct_id])
        # starting momenta on the host: the previous result of this path if a0[k] is still that tensor, else one download
        cache = getattr(self, "_a0_host", None)
        p0 = []
        for k in frames:
            c = cache[k] if cache is not None else None
            if c is not None and c[0] is self.a0[k] and c[1] == self.a0[k]._version:
                p0.append(c[2])
            else:
                p0.append(self.a0[k].detach().cpu().numpy())
        best_p, _, steps, change, _ = LBFGS_optimization_lockstep(p0, plan, nmax=nmax, tol=tol)
        radius = 2.0 * self.LMi.Kernel.sigma if self.support_scheme is not None else None
        traj, trajl, datal, counts = plan.finalize(best_p, coverage_radius=radius)
        results = []
        for j, k in enumerate(frames):
            shoot = plan.frame_states(traj, j)              # lazy: state tuples are built on first access
            shoot.__class__ = ShootResult
            M, MD, Nx = plan.Ms[j], plan.Ms[j] * D, plan.Nxs[j]
            a0 = traj[0, j, MD:2 * MD].view(M, D)
            x1 = traj[-1, j, 2 * MD:2 * MD + Nx * D].view(Nx, D) if Nx else traj[-1, j, :MD].view(M, D)
            results.append((k, (a0, a0._version, best_p[j]),
                            dict(lockstep=True, a0=a0, shoot=shoot, regloss=float(trajl[j]), datal=float(datal[j]), isteps=steps[j],
                                 change=change[j], x1=x1, counts=None if counts is None else counts[j].tolist())))
        return results

    def _register_all_lockstep(self, plans, nmax, tol):
        K = self.K
        if len(plans) == 1:
            per_group = [self._register_group_lockstep(plans[0], nmax, tol)]
        else:
            import threading
            dev = torch.device(self.compspec["device"])
            main = torch.cuda.current_stream(dev)
            per_group, errors = [None] * len(plans), []

            def work(g):
                try:
                    torch.cuda.set_device(dev)
                    with torch.cuda.stream(plans[g].stream):
                        plans[g].stream.wait_stream(main)
                        per_group[g] = self._register_group_lockstep(plans[g], nmax, tol)
                except BaseException as e:          # surfaced on the main thread
                    errors.append(e)
            threads = [threading.Thread(target=work, args=(g,)) for g in range(len(plans))]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            for pl in plans:
                main.wait_stream(pl.stream)
            if errors:
                raise errors[0]
        results, self._a0_host = [None] * K, [None] * K
        for group in per_group:
            for k, host, r in group:
                results[k], self._a0_host[k] = r, host
        return results

    def _register_all(self, nmax, tol):
        K = self.K
        dev = torch.device(self.compspec["device"])
        plan = self._batched_plan()
        if plan is not None:
            return self._register_all_lockstep(plan, nmax, tol)
        W = min(self._auto_frame_workers(), K)
        if W <= 1 or dev.type != "cuda":
            return [self._register_frame(k, nmax, tol) for k in range(K)]
        import threading
        from .. import shooting
        if self.LMi.use_cuda_graph:          # capture every plan up front, sequentially, on this thread
            for k in range(K):
                sp = self.LMi._spec_for(self.q0[k].shape[0], 0 if self.support_scheme is None else self.allx0[k].shape[0], dev)
                shooting.ShootPlan.get(sp, True, slot=1 + k % W).ensure_captured()
        main = torch.cuda.current_stream(dev)
        streams = [torch.cuda.Stream(dev) for _ in range(W)]
        results, errors = [None] * K, []

        def work(wid):
            try:
                shooting.set_plan_slot(1 + wid)
                torch.cuda.set_device(dev)
                with torch.cuda.stream(streams[wid]):
                    streams[wid].wait_stream(main)
                    for k in range(wid, K, W):
                        results[k] = self._register_frame(k, nmax, tol)
            except BaseException as e:          # surfaced on the main thread
                errors.append(e)

        threads = [threading.Thread(target=work, args=(w,)) for w in range(W)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for st in streams:
            main.wait_stream(st)
        if errors:
            raise errors[0]
        return results

    def Reg_opt(self, nmax=10, tol=1e-3):
        """LDDMM registration of every (local) frame to its current targets (reference: core/PSR.py:521-569)."""
        results = self._register_all(nmax, tol)
        lockstep = bool(results) and results[0].get("lockstep", False)
        for k, r in enumerate(results):
            self.a0[k], self.shoot[k], self.regloss[k] = r["a0"], r["shoot"], r["regloss"]
            last = 0
            for s in range(self.S):
                first, last = last, last + self.N[k, s]
                self.x1[k, s] = r["x1"][first:last].to(**self.dataspec)
            if lockstep and self.S == 1:
                pass        # one structure: quadloss[:, 0] = the frames' fused data losses, set for all frames below
            else:
                for s in range(self.S):
                    self.update_quadloss(k, s)
            if r["counts"] is not None:
                for t, c in enumerate(r["counts"]):
                    if c > 0:
                        print(f"WARNING : shooting, time step {t} : {c} uncovered points ({c / self.allx0[k].shape[0]:.2%})")
                        warnings.warn("Uncovered points during LDDMM shooting. Choose a smaller rho when defining the support scheme.", RuntimeWarning)
            chg = r["change"]
            chg = f"{chg:.4}" if isinstance(chg, float) else str(chg)
            message = f"Frame {k} : {r['isteps']} optim steps, loss={r['regloss'] + r['datal']:.4}, change={chg}."
            if self.comm is None and not lockstep:
                self.update_FE(message=message)
            elif self.printstuff:
                print(message)
        if lockstep and self.S == 1:
            # one structure: quadloss[k,0] is the frame's data loss sum_n |x1_n - y_n|^2 / (2 sigma^2), already reduced
            # (deterministically) by the fused closure -- one host-to-device copy for all frames
            self.quadloss[:, 0] = torch.tensor([r["datal"] for r in results], dtype=self.quadloss.dtype).to(self.quadloss.device)
        if self.comm is not None or lockstep:   # all frames were registered together: ONE free-energy update (and, with
            self.update_FE(message="Registration of all frames done.")      # ranks, ONE collective) per Reg_opt
