"""Isotropic-uniform Gaussian mixture with a fused EM step on the B200: drop-in for the reference's
``diffICP/core/GMM.py`` (live class, core/GMM.py:40-721).

Model: centroids mu (C,D), ONE shared scalar sigma, component scores w (pi = softmax(w)), optional uniform outlier
component encoded by a log-odds ratio eta0 and a reference volume vol0 (core/GMM.py:42-109).

EM step (core/GMM.py:236-325 torch twin, :402-529 KeOps formulation), per call:
  * mu, w frozen or skip_M=True : ONE sweep over points x components (dicp_em_rowpass) produces the row
    log-sum-exp, the targets Y, and the four sums from which sigma', Cfe and FE follow;
  * otherwise                    : row LSE sweep -> column sweep (log-domain sufficient statistics of every
    component, dicp_em_colstats) -> tiny M step on (C, D+3) numbers -> full row sweep (old gamma, new theta).
One host synchronisation per EM step (the reference has >= 3 ``.item()`` calls).

``computversion`` keeps the reference's two SEMANTIC variants (they differ when the M step runs):
  "keops" / "b200" : sigma' from squared distances to the NEW centroids, Gaussian normalisation from the NEW sigma
                     (core/GMM.py:453-455, :483) -- the reference's default and GPU behaviour;
  "torch"          : sigma' from distances to the OLD centroids, normalisation from the OLD sigma (:263-264, :296, :314).
Both run on the same CUDA kernels; nothing here computes on the CPU.

Multi-GPU (groupwise mode, points sharded by frame): set ``self.comm`` to a ``diff_icp_b200.dist.StatsComm``; the
column statistics and the four sums are then all-reduced (MAX on the exponents, SUM on the rescaled sums) -- the only
collective of the whole algorithm.
"""

from __future__ import annotations

import copy
import math

import numpy as np
import torch
from torch.nn import Module
from torch.nn.functional import log_softmax, softmax, softplus

from .. import em_ops
from ..tools.point_sets import intrinsic_scale
from ..tools.spec import defspec

_LOG2E = 1.4426950408889634
_LN2 = 0.6931471805599453
_ACCEPTED = ("b200", "keops", "torch")


class GaussianMixtureUnif(Module):

    def __init__(self, mu, sigma=None, use_outliers=False, spec=defspec, computversion="keops"):
        super().__init__()
        self.params = {}
        self.spec = spec
        self.mu = mu.clone().detach().to(**spec)
        self.C, self.D = self.mu.shape
        self.sigma = sigma
        if self.sigma is None:
            # ad hoc: 0.1 x "typical radius" of the volume owned by one centroid (reference: core/GMM.py:83-88)
            r = self.mu.var(0).sum().sqrt().item()
            self.sigma = max(0.1 * (r / self.C ** (1 / self.D)), 1e-6)
        self.w = torch.zeros(self.C, **spec)
        self.to_optimize = {"sigma": True, "mu": True, "w": True, "eta0": True}
        self.outliers = {"vol0": None, "eta0": 0.0} if use_outliers else None
        self.ensure_continuum = False
        self.comm = None                      # multi-GPU statistics reducer (diff_icp_b200.dist.StatsComm) or None
        self.set_computversion(computversion)

    # ------------------------------------------------------------------------------------------------------
    def __deepcopy__(self, memo):
        G2 = GaussianMixtureUnif(self.mu, spec=self.spec, computversion=self.computversion)
        G2.sigma = self.sigma
        G2.w = self.w.clone().detach()
        G2.to_optimize = copy.deepcopy(self.to_optimize)
        G2.outliers = copy.deepcopy(self.outliers)
        G2.ensure_continuum = self.ensure_continuum
        G2.comm = self.comm
        return G2

    def set_computversion(self, version):
        if version not in _ACCEPTED:
            raise ValueError(f"unkown computversion : {version}. Choices are 'b200', 'keops' or 'torch'")
        self.computversion = version
        self.EM_step = self.EM_step_b200
        return self

    def fix(self):
        self.to_optimize = {"sigma": False, "mu": False, "w": False, "eta0": False}
        return self

    def set_vol0(self, X: torch.Tensor):
        """Reference volume of the outlier distribution = bounding box of the data points (reference: core/GMM.py:163-172).
        With `comm` (multi-GPU): the bounding box of the points of ALL ranks, so that every rank uses the same volume as the
        single-process run (a rank without points contributes nothing)."""
        if self.outliers is not None:
            if self.comm is None:
                self.outliers["vol0"] = (X.max(dim=0)[0] - X.min(dim=0)[0]).prod().item()
            else:
                if X.shape[0] > 0:
                    both = torch.cat((-X.min(dim=0)[0], X.max(dim=0)[0])).to(**self.spec)
                else:
                    both = torch.full((2 * self.D,), -float("inf"), **self.spec)
                both = self.comm.max(both)
                self.outliers["vol0"] = (both[self.D:] + both[:self.D]).prod().item()
        return self

    def __str__(self):
        s = super().__str__()
        s += ": Gaussian Mixture with Uniform covariances. Parameters:\n"
        s += "    C [# components] : " + str(self.C) + "\n"
        s += "    sigma [unif. std] : " + str(self.sigma) + "\n"
        s += "    mu_c [centroids] :" + str(self.mu) + "\n"
        s += "    w_c [component scores]:" + str(self.w) + "\n"
        if self.outliers is not None:
            s += "    vol0 [ref. volume for outliers]:" + str(self.outliers["vol0"]) + "\n"
            s += "    eta0 [outlier vs GMM log-ratio]:" + str(self.outliers["eta0"]) + "\n"
        return s

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.spec = defspec
        self.EM_step = self.EM_step_b200

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("EM_step", None)            # bound method: rebuilt on load
        state.pop("_lpi_cache", None)
        state.pop("_em_loop", None)          # device buffers + captured CUDA graph of the EM loop
        state.pop("_m_ref", None)
        state["comm"] = None
        return state

    # ------------------------------------------------------------------------------------------------------
    def log_ratio_to_proba(self, eta):
        """(log p, log q) of a Bernoulli with log-odds eta = log(p/q) (reference: core/GMM.py:205-217)."""
        if not isinstance(eta, torch.Tensor):
            eta = torch.tensor(eta, **self.spec)
        Z = softplus(eta)
        return eta - Z, -Z

    def log_responsibilities(self, X):
        """(N,C) log gamma_nc without outliers (reference: core/GMM.py:221-232)."""
        lgam, _ = em_ops.log_resp(self.sigma, X.detach().contiguous(), self.mu.contiguous(), self.w.contiguous())
        return lgam

    def hard_assignments(self, X):
        """argmax_c gamma_nc, (N,) int64, first index on ties (what the reference computes as
        log_responsibilities(X).argmax(dim=1), core/GMM.py:677-680) without materialising the (N,C) matrix."""
        _, amax = em_ops.log_resp(self.sigma, X.detach().contiguous(), self.mu.contiguous(), self.w.contiguous(),
                                  want_lgam=False, want_argmax=True)
        return amax

    # ------------------------------------------------------------------------------------------------------
    def _lgn(self, sigma):
        return self.D * (math.log(sigma) + 0.5 * math.log(2 * math.pi))

    # ---- multi-GPU: the exponent every rank rescales its column statistics to (one SUM all-reduce, no MAX round) ----
    def _agreed_exponent(self):
        """(C,) log2 column masses of the previous merged statistics -- identical bits on every rank because they are
        computed from all-reduced numbers -- or None when there is none yet (first step, changed model size)."""
        m = getattr(self, "_m_ref", None)
        return m if (m is not None and m.shape[0] == self.C and m.device == self.mu.device) else None

    def _remember_exponent(self, merged):
        # round(): rescaling by an integer power of two is exact
        self._m_ref = torch.round(merged[:, 0] + torch.log2(torch.clamp(merged[:, 1], min=1e-30)))

    def EM_step_b200(self, X, skip_M=False, _safe_merge=False):
        """One (E step, M step) alternation; returns (Y, Cfe, FE) like the reference (core/GMM.py:236-325, :501-529).
        Y (N,D): quadratic targets sum_c gamma_nc mu_c;  Cfe: free-energy offset;  FE = Cfe + sum_n (1-gamma0_n)|x_n-y_n|^2/(2 sigma^2)."""
        X = X.detach().contiguous()
        N_local = X.shape[0]
        D, C = self.D, self.C
        opt = self.to_optimize
        keops_sem = self.computversion != "torch"
        do_M = not skip_M
        do_mu, do_w, do_sig = do_M and opt["mu"], do_M and opt["w"], do_M and opt["sigma"]
        use_out = self.outliers is not None
        comm = self.comm

        sigma_old = float(self.sigma)
        lgn_old = self._lgn(sigma_old)
        mu_old = self.mu.contiguous()
        w_old = self.w.contiguous()
        # log pi of the current weights: kept from the previous M step when self.w is still that tensor
        cache = getattr(self, "_lpi_cache", None)
        if cache is not None and cache[0] is self.w and cache[1] == self.w._version:
            lpi_old = cache[2]
        else:
            lpi_old = w_old - torch.logsumexp(w_old, 0)
        wl2 = ((lpi_old - lgn_old) * _LOG2E).contiguous()

        # ---- M step for mu / w (and the column form of sigma) from log-domain column statistics, ONE launch --
        ms = None
        mu_new, w_new, lpi_new = mu_old, w_old, lpi_old
        merge_check = None
        if do_mu or do_w:
            stats = em_ops.lse_colstats(sigma_old, X, mu_old, wl2) if N_local > 0 else _empty_stats(C, D, X.device)
            if comm is not None:
                m_ref = None if _safe_merge else self._agreed_exponent()
                if m_ref is None:
                    stats = comm.merge_colstats(stats)                      # MAX round + SUM round
                else:
                    stats, _, flag = comm.merge_colstats_ref(stats, m_ref)  # ONE round, checked at the step's host read
                    merge_check = torch.stack((flag, (stats[:, 1] < 1e-30).any().to(flag.dtype)))
                self._remember_exponent(stats)
            sig_mode = 0 if not do_sig else (1 if (keops_sem and do_mu) else 2)
            mu_new, w_new, lpi_new, ms = em_ops.mstep(stats, mu_old, w_old, do_mu, do_w, sig_mode)
        nds2 = ms[0] if (ms is not None and do_sig) else None
        lpi_new = lpi_new.contiguous()

        # ---- full row pass: old responsibilities, new centroids / weights ------------------------------------
        T2, Y, scal, rowP, rowQ, sq = em_ops.rowpass(sigma_old, X, mu_old, wl2, mu_new, lpi_new, per_point=use_out)
        parts = [scal, scal.new_full((1,), float(N_local))]
        if nds2 is not None:
            parts.append(nds2.reshape(1))
        if use_out:
            if self.outliers["vol0"] is None:
                self.set_vol0(X)
            logJ0 = -math.log(self.outliers["vol0"])
            eta_n = self.outliers["eta0"] + logJ0 - T2 * _LN2
            lg0, lgT = self.log_ratio_to_proba(eta_n)
            if do_M and opt["eta0"]:
                # log-domain sums over n (reference: core/GMM.py:290 / :450)
                l0, lT = torch.logsumexp(lg0, 0), torch.logsumexp(lgT, 0)
                if comm is not None:
                    l0, lT = comm.logsumexp_pair(l0, lT)
                parts.append(torch.stack((l0, lT)))
        vec = torch.cat(parts)
        if comm is not None:
            head = comm.sum(vec[:5].clone())                    # P, Q, SQ, DS, N are plain sums over frames
            vec = torch.cat((head, vec[5:]))
        if merge_check is not None:
            vec = torch.cat((vec, merge_check))
        vals = vec.tolist()                                     # the ONE host synchronisation of this EM step
        if merge_check is not None:
            if vals[-2] > 0 or vals[-1] > 0:                    # exponent drifted too far from the agreed one (every rank
                self._m_ref = None                              # sees the same flags): redo the step with the MAX round
                return self.EM_step_b200(X, skip_M=skip_M, _safe_merge=True)
            vals = vals[:-2]
        P, Q, SQ, DS, N = vals[:5]
        if comm is None:
            N = float(N_local)
        k = 5
        if do_sig:
            nd = vals[k] if nds2 is not None else DS
            k += 1 if nds2 is not None else 0
            self.sigma = math.sqrt(max(nd, 0.0) / (D * N))
            if self.ensure_continuum:
                self.sigma = max(self.sigma, intrinsic_scale(mu_new))
        if use_out and do_M and opt["eta0"]:
            self.outliers["eta0"] = vals[k] - vals[k + 1]
        if do_mu:
            self.mu = mu_new
        if do_w:
            self.w = w_new
        self._lpi_cache = (self.w, self.w._version, lpi_new)

        sig = float(self.sigma)
        lgn = self._lgn(sig) if keops_sem else lgn_old
        inv2s2 = 1.0 / (2 * sig * sig)
        if not use_out:
            Cfe_val = P * inv2s2 + Q + N * lgn
            FE_val = Cfe_val + SQ * inv2s2
            # 0-d fp32 tensors like the reference's return values; kept on the host (they are only ever read back as
            # Python numbers -- EM_optimization's stop test, MultiPSR.update_FE -- and a device copy costs a transfer each)
            Cfe = torch.tensor(Cfe_val, dtype=self.spec["dtype"])
            FE = torch.tensor(FE_val, dtype=self.spec["dtype"])
            return Y, Cfe, FE
        g0, gT = lg0.exp(), lgT.exp()
        lpi0, lpiT = self.log_ratio_to_proba(self.outliers["eta0"])
        cfe_n = rowP * inv2s2 + rowQ + lgn
        tot = torch.stack(((gT * (cfe_n + lgT - lpiT) + g0 * (-logJ0 + lg0 - lpi0)).sum(), (gT * sq).sum()))
        if comm is not None:
            tot = comm.sum(tot)
        c_out, q_out = tot.tolist()
        return Y, c_out, c_out + q_out * inv2s2

    # ------------------------------------------------------------------------------------------------------
    def EM_optimization(self, X, max_iterations=100, tol=1e-5):
        """Repeat EM steps until the relative change of FE is below tol (reference: core/GMM.py:330-357).
        Returns (Y, Cfe, FE, number of steps)."""
        if X.shape[0] == 0 and self.comm is None:
            return torch.empty(X.shape, **self.spec), torch.tensor(0.0), torch.tensor(0.0), 0
        if self._pipelined_applies():
            return self._EM_optimization_pipelined(X, max_iterations, tol)
        if self._graph_loop_applies(X, max_iterations):
            return self._EM_optimization_graph(X, max_iterations, tol)
        Y = Cfe = FE = last_FE = None
        for i in range(max_iterations):
            Y, Cfe, FE = self.EM_step(X)
            if last_FE is not None and tol is not None and abs(FE - last_FE) < tol * abs(last_FE):
                return Y, Cfe, FE, i + 1
            last_FE = FE
        print(f"GMM optimization - reached maximum number of iterations : {max_iterations}")
        return Y, Cfe, FE, i + 1

    # ---- one GPU, few components: the whole loop as one CUDA graph replay (em_loop.py) ----------------------------------------
    graph_em_loop = True

    def _graph_loop_applies(self, X, max_iterations):
        opt = self.to_optimize
        return (self.graph_em_loop and self.comm is None and self.outliers is None and X.is_cuda and X.shape[0] > 0
                and self.EM_step == self.EM_step_b200 and self.C <= em_ops.SMALL_C and (opt["mu"] or opt["w"])
                and not (opt["sigma"] and self.ensure_continuum) and 1 <= max_iterations <= 1000
                and X.dtype == torch.float32 and self.mu.is_cuda)

    def _EM_optimization_graph(self, X, max_iterations, tol):
        from ..em_loop import EMLoopGraph
        opt = self.to_optimize
        keops_sem = self.computversion != "torch"
        sig_mode = 0 if not opt["sigma"] else (1 if (keops_sem and opt["mu"]) else 2)
        X = X.detach().contiguous()
        key = (X.shape[0], self.C, self.D, str(X.device), bool(opt["mu"]), bool(opt["w"]), sig_mode, keops_sem)
        loop = getattr(self, "_em_loop", None)
        if loop is None or loop[0] != key:
            loop = (key, EMLoopGraph(X.shape[0], self.C, self.D, X.device, opt["mu"], opt["w"], sig_mode, keops_sem))
            self._em_loop = loop
        w_old = self.w.contiguous()
        cache = getattr(self, "_lpi_cache", None)
        if cache is not None and cache[0] is self.w and cache[1] == self.w._version:
            lpi_old = cache[2]
        else:
            lpi_old = w_old - torch.logsumexp(w_old, 0)
        Y, mu, w, lpi, sigma, Cfe_val, FE_val, steps, stopped = loop[1].run(X, self.mu.contiguous(), w_old, lpi_old,
                                                                            float(self.sigma), tol, int(max_iterations))
        if opt["sigma"]:
            self.sigma = sigma
        if opt["mu"]:
            self.mu = mu
        if opt["w"]:
            self.w = w
        self._lpi_cache = (self.w, self.w._version, lpi)
        if not stopped:
            print(f"GMM optimization - reached maximum number of iterations : {max_iterations}")
        return (Y, torch.tensor(Cfe_val, dtype=self.spec["dtype"]), torch.tensor(FE_val, dtype=self.spec["dtype"]), steps)

    # ---- multi-GPU: ONE all-reduce per EM step ------------------------------------------------------------------------
    pipelined_allreduce = True

    def _pipelined_applies(self):
        """The pipelined loop needs the next step's sigma right after the merged column statistics: the KeOps ordering with
        mu optimised (sigma' from the column statistics, core/GMM.py:453-455), or a fixed sigma.  Other settings (torch
        ordering with sigma from the row sums, outliers, frozen mu and w) keep the step-by-step loop."""
        opt = self.to_optimize
        return (self.comm is not None and self.pipelined_allreduce and self.outliers is None
                and self.EM_step == self.EM_step_b200 and (opt["mu"] or opt["w"])
                and (not opt["sigma"] or (self.computversion != "torch" and opt["mu"])))

    def _EM_optimization_pipelined(self, X, max_iterations, tol):
        """EM_optimization with the points sharded over ranks and ONE all-reduce per EM step.

        Step i needs two global reductions: the column statistics (before its M step) and the four free-energy sums
        [P, Q, SQ, DS] of its row pass (after it).  The sums of step i are only needed on the host -- for FE_i and the stop
        test -- so they ride on the all-reduce of step i+1's column statistics: that step's first half (row LSE + column
        statistics with the new parameters, ~50 us) is issued SPECULATIVELY; if FE_i then says "stop", its result is
        dropped and the model stays at step i's parameters, exactly what the step-by-step loop returns.  The buffer is
        [S0, B, A rescaled to the rank-agreed exponents | overflow flag | P, Q, SQ, DS, N]; a final all-reduce of the five
        sums alone closes the loop when the step limit is reached.  Same arithmetic as EM_step_b200 on every rank."""
        comm, D, C = self.comm, self.D, self.C
        opt = self.to_optimize
        do_mu, do_w, do_sig = opt["mu"], opt["w"], opt["sigma"]
        X = X.detach().contiguous()
        N_local = X.shape[0]
        dev = self.mu.device
        sig_mode = 1 if do_sig else 0

        def first_half(sigma, mu, lpi):
            wl2 = ((lpi - self._lgn(sigma)) * _LOG2E).contiguous()
            if N_local == 0:
                return wl2, _empty_stats(C, D, dev)
            return wl2, em_ops.lse_colstats(sigma, X, mu, wl2)

        def values(sums, sigma_new):
            P, Q, SQ, DS, N = sums
            inv2s2 = 1.0 / (2 * sigma_new * sigma_new)
            Cfe_val = P * inv2s2 + Q + N * self._lgn(sigma_new)
            return Cfe_val, Cfe_val + SQ * inv2s2, N

        sigma = float(self.sigma)
        mu, w = self.mu.contiguous(), self.w.contiguous()
        lpi = w - torch.logsumexp(w, 0)
        wl2, stats = first_half(sigma, mu, lpi)
        pending = None            # (Y, local sums, sigma after that step) of the last completed step, FE not yet known
        last_FE, done, N_glob = None, None, None
        zeros5 = torch.zeros(5, dtype=torch.float32, device=dev)
        i = 0
        while True:
            # ---- the ONE collective of step i: its column statistics + the previous step's sums ------------------------
            extra = pending[1] if pending is not None else torch.cat((zeros5[:4], zeros5.new_full((1,), float(N_local))))
            m_ref = self._agreed_exponent()
            if m_ref is None:
                merged = comm.merge_colstats(stats)
                red = comm.sum(extra.clone())
                mu_new, w_new, lpi_new, ms = em_ops.mstep(merged, mu, w, do_mu, do_w, sig_mode)
                host = torch.cat((ms[:1], red, zeros5[:2])).tolist()
                self._remember_exponent(merged)
            else:
                # pack (one launch) -> all-reduce -> M step on the reduced buffer (one launch) -> ONE host read
                buf = comm.sum(em_ops.reduce_pack(stats, m_ref, extra))
                mu_new, w_new, lpi_new, m_next, hostv = em_ops.mstep_merged(buf, m_ref, mu, w, do_mu, do_w, sig_mode, 5)
                host = hostv.tolist()
                if host[-2] > 0 or host[-1] > 0:                        # exponents drifted: repeat with the MAX round
                    self._m_ref = None
                    continue
                self._m_ref = m_next
            N_glob = host[5]
            if pending is not None:                                     # FE of step i-1 is now known
                Cfe_val, FE_val, _ = values(host[1:6], pending[2])
                if last_FE is not None and tol is not None and abs(FE_val - last_FE) < tol * abs(last_FE):
                    done = (pending[0], Cfe_val, FE_val, i)             # step i's speculative first half is dropped
                    break
                last_FE = FE_val
            if i == max_iterations:
                break
            # ---- second half of step i: new parameters, row pass with the OLD responsibilities -------------------------
            sigma_new = math.sqrt(max(host[0], 0.0) / (D * N_glob)) if do_sig else sigma
            if do_sig and self.ensure_continuum:
                sigma_new = max(sigma_new, intrinsic_scale(mu_new))
            lpi_new = lpi_new.contiguous()
            if N_local > 0:
                _, Y, scal, _, _, _ = em_ops.rowpass(sigma, X, mu, wl2, mu_new, lpi_new)
            else:
                Y, scal = torch.empty(0, D, dtype=torch.float32, device=dev), zeros5[:4]
            pending = (Y, torch.cat((scal, scal.new_full((1,), float(N_local)))), sigma_new)
            sigma, mu, w, lpi = sigma_new, (mu_new if do_mu else mu), (w_new if do_w else w), lpi_new
            self.sigma, self.mu, self.w = sigma, mu, w
            self._lpi_cache = (self.w, self.w._version, lpi)
            i += 1
            if i < max_iterations:
                wl2, stats = first_half(sigma, mu, lpi)                 # speculative first half of the next step
            else:                                                       # step limit: only the sums remain to be reduced
                red = comm.sum(pending[1].clone()).tolist()
                Cfe_val, FE_val, _ = values(red, pending[2])
                print(f"GMM optimization - reached maximum number of iterations : {max_iterations}")
                done = (pending[0], Cfe_val, FE_val, i)
                break
        Y, Cfe_val, FE_val, steps = done
        return Y, torch.tensor(Cfe_val, dtype=self.spec["dtype"]), torch.tensor(FE_val, dtype=self.spec["dtype"]), steps

    @staticmethod
    def get_GMM_model(X, C, fixed_sigma=None, optimize_w=False, use_outliers=False, max_iterations=100, tol=1e-5,
                      spec=defspec, computversion="keops"):
        """GMM with C components fitted to X from C random data points (reference: core/GMM.py:361-383)."""
        mu = X[torch.randint(0, X.shape[0], (C,)), :]
        GMM = GaussianMixtureUnif(mu, use_outliers=use_outliers, spec=spec, computversion=computversion)
        GMM.to_optimize = {"mu": True, "sigma": True, "w": optimize_w, "eta0": True}
        if fixed_sigma is not None:
            GMM.to_optimize["sigma"] = False
            GMM.sigma = fixed_sigma
        GMM.EM_optimization(X, max_iterations=max_iterations, tol=tol)
        return GMM

    # ------------------------------------------------------------------------------------------------------
    def pi(self):
        return softmax(self.w, dim=0)

    def get_sample(self, N):
        """N random points from the mixture, without the outlier term (reference: core/GMM.py:543-550)."""
        samp = self.sigma * torch.randn(N, self.D, **self.spec)
        c = torch.distributions.categorical.Categorical(logits=self.w).sample((N,))
        return samp + self.mu[c, :]

    def update_covariances(self):
        self.params["gamma"] = (torch.eye(self.D, **self.spec) * self.sigma ** (-2))[None, :, :].repeat([self.C, 1, 1]).view(self.C, self.D ** 2)

    def weights(self):
        return softmax(self.w, 0) / self.sigma ** self.D

    def weights_log(self):
        return log_softmax(self.w, 0) - self.D * math.log(self.sigma)

    def log_likelihoods(self, sample):
        """Log-density sampled on a point cloud, with the reference's normalisation (core/GMM.py:714-721:
        weights_log already carries -D ln sigma and loggaussnorm is subtracted on top of it)."""
        sample = sample.to(**self.spec).contiguous()
        wl2 = ((self.weights_log() - self._lgn(self.sigma)) * _LOG2E).contiguous()
        T2 = em_ops.rowpass(self.sigma, sample, self.mu.contiguous(), wl2)
        return T2 * _LN2

    def likelihoods(self, sample):
        """Density sampled on a point cloud (reference: core/GMM.py:706-712)."""
        return self.log_likelihoods(sample).exp()


def _empty_stats(C, D, device):
    s = torch.zeros(C, D + 3, dtype=torch.float32, device=device)
    s[:, 0] = -3.0e38
    return s
