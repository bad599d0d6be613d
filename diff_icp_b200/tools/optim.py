"""L-BFGS driver used by LDDMMModel.Optimize: host-side control flow, same behaviour as the reference's
``LBFGS_optimization`` (tools/optim.py:10-110):

* ``torch.optim.LBFGS(max_iter=20, max_eval=100, history_size=100, line_search_fn="strong_wolfe")`` (:26);
* at most ``nmax`` optimizer steps, stop when the RMS change of every parameter is below ``tol`` x its RMS (:98-104);
* divergence guard: NaN / loss above ``errthresh`` / loss increase -> fall back to the best parameters seen so far,
  or perturb by 1 % noise, and restart L-BFGS *without* line search (:60-97);
* returns the BEST parameters seen over all closure evaluations, not the last ones (:108-110).

It stays on the host on purpose (parity of iterates with the reference); every closure evaluation is one fused
shoot + one adjoint sweep on the device, and one ``.item()`` synchronisation as in the reference (:39).
"""

import math

import torch


def LBFGS_optimization(p0, lossfunc, nmax=10, tol=1e-3, errthresh=1e8, lossgrad=None):
    """`lossgrad` (optional, B200 build): callable returning (loss, [gradients]) directly -- used by LDDMMModel.Optimize when
    the whole closure (shoot + quadratic loss + adjoint) runs as one captured launch sequence; the optimiser then sees
    exactly the same loss / gradient values as through `lossfunc` + `.backward()`, without building an autograd graph."""
    params = [t.clone().contiguous().detach().requires_grad_(True) for t in p0]

    def new_optimizer(line_search):
        return torch.optim.LBFGS(params, max_iter=20, max_eval=100, history_size=100, line_search_fn=line_search)

    optimizer = new_optimizer("strong_wolfe")
    history = []
    best = {"L": math.inf, "p": None}

    def closure():
        optimizer.zero_grad()
        if lossgrad is not None:
            val, grads = lossgrad(*[t.detach() for t in params])
            val = float(val)
            loss = torch.tensor(val)
        else:
            loss = lossfunc(*params)
            val = loss.detach().item()
        history.append(val)
        if val < best["L"]:
            best["L"] = val
            best["p"] = [t.clone().detach() for t in params]
        if lossgrad is not None:
            for t, g in zip(params, grads):
                t.grad = g.to(device=t.device, dtype=t.dtype).clone()
        else:
            loss.backward()
        return loss

    step, go_on, L, change = 0, True, math.inf, None
    while step < nmax and go_on:
        step += 1
        before = [t.clone().detach() for t in params]
        optimizer.step(closure)
        L_before, L = L, history[-1]

        if L > L_before or L > errthresh or math.isnan(L):
            if math.isnan(L):
                print("WARNING: NaN value for loss L during L-BFGS optimization.")
            elif L > errthresh:
                print("WARNING: Aberrantly large value for loss L during L-BFGS optimization.")
            else:
                print("WARNING: Increase of loss L during L-BGFS optimization.")
            if best["L"] < L_before:
                params = [t.clone() for t in best["p"]]
                L = best["L"]
                print("L-BFGS optimization. Found an intermediate 'best_p' value for this iteration.")
            else:
                rmod = 0.01
                params = [t + rmod * t.std() * torch.randn(t.shape, dtype=t.dtype, device=t.device) for t in best["p"]]
                L = lossfunc(*params) if lossgrad is None else lossgrad(*params)[0]
                print(f"L-BFGS optimization. Trying a random perturbation of parameter from its current value, with relative strength {rmod}.")
            change = "None (divergent iteration step)"
            params = [t.requires_grad_(True) for t in params]
            optimizer = new_optimizer(None)
        else:
            deltas = [((t - b) ** 2).mean().sqrt().detach().cpu().numpy() for t, b in zip(params, before)]
            scales = [(b ** 2).mean().sqrt().detach().cpu().numpy() for b in before]
            go_on = any(dl > tol * sc for dl, sc in zip(deltas, scales))
            change = max(deltas)

    return [t.detach() for t in best["p"]], best["L"], step, change


# ---------------------------------------------------------------------------------------------------------------------
# Lock-step form for K independent problems (B200 build)

class LockstepLBFGS:
    """K independent torch.optim.LBFGS-like optimisers (max_iter=20, max_eval=100, history_size=100, strong-Wolfe line
    search: the settings of tools/optim.py:26) advanced together: each round, every optimiser that needs a closure value
    writes its trial point into row k of `X`, ONE call of `evaluate` serves all of them, and every optimiser consumes its
    own (loss, gradient).  The per-frame state machines are native host code (csrc/lbfgs_batch.cu, C ABI
    dicp_lbfgs_*); this class only owns the buffers."""

    def __init__(self, sizes, stride=None, max_iter=20, max_eval=100, history_size=100,
                 tolerance_grad=1e-7, tolerance_change=1e-9):
        import ctypes
        import numpy as np
        from .._lib import load
        self._lib = load()
        self.K = len(sizes)
        self.sizes = [int(n) for n in sizes]
        self.stride = int(stride) if stride is not None else max(self.sizes)
        arr = (ctypes.c_int64 * self.K)(*self.sizes)
        self._h = self._lib.dicp_lbfgs_create(self.K, arr, self.stride, max_iter, max_eval, history_size,
                                              float(tolerance_grad), float(tolerance_change))
        if not self._h:
            raise ValueError("LockstepLBFGS: bad sizes")
        self._np = np
        self._ct = ctypes

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.dicp_lbfgs_destroy(h)

    def _fp(self, a):
        return a.ctypes.data_as(self._ct.c_void_p)

    def set_x(self, k, x):
        a = self._np.ascontiguousarray(x, dtype=self._np.float32).reshape(-1)
        assert a.size == self.sizes[k]
        self._lib.dicp_lbfgs_set_x(self._h, k, self._fp(a))

    def get_x(self, k, best=False):
        a = self._np.empty(self.sizes[k], dtype=self._np.float32)
        if self._lib.dicp_lbfgs_get_x(self._h, k, self._fp(a), int(best)) != 0:
            raise RuntimeError("LockstepLBFGS.get_x: no value")
        return a

    def reset(self, k, line_search=True):
        """A fresh optimiser for frame k (tools/optim.py:77 restarts without line search after a divergent step)."""
        self._lib.dicp_lbfgs_reset(self._h, k, int(bool(line_search)))

    def stats(self, k):
        out = (self._ct.c_double * 4)()
        self._lib.dicp_lbfgs_stats(self._h, k, out)
        return {"last": out[0], "best": out[1], "func_evals": int(out[2]), "n_iter": int(out[3])}

    def get_all(self, best=False):
        """(K, stride) fp32 array of every frame's current (or best-so-far) parameters; row k is valid in [:sizes[k]]."""
        X = self._np.zeros((self.K, self.stride), dtype=self._np.float32)
        if self._lib.dicp_lbfgs_get_all(self._h, self._fp(X), int(best)) != 0:
            raise RuntimeError("LockstepLBFGS.get_all: no value")
        return X

    def stats_all(self):
        """(K, 4) fp64 array: last closure value, best closure value, closure evaluations, L-BFGS iterations."""
        out = self._np.zeros((self.K, 4), dtype=self._np.float64)
        self._lib.dicp_lbfgs_stats_all(self._h, self._fp(out))
        return out

    def step(self, mask, evaluate, X, active, losses, grads):
        """One optimizer.step(closure) of every frame selected by `mask` (uint8 (K,)).  X / grads: (K, stride) fp32 numpy
        arrays, active: (K,) uint8, losses: (K,) fp32 -- the caller's (pinned) buffers; `evaluate()` must fill losses and
        grads for the frames flagged in `active` from the rows of X.  Returns the number of evaluation rounds."""
        m = self._np.ascontiguousarray(mask, dtype=self._np.uint8)
        if self._lib.dicp_lbfgs_begin_step(self._h, self._fp(m)) < 0:
            raise RuntimeError("LockstepLBFGS.step: an optimiser step is already in progress")
        rounds = 0
        while self._lib.dicp_lbfgs_pending(self._h, self._fp(X), self._fp(active)) > 0:
            evaluate()
            rounds += 1
            self._lib.dicp_lbfgs_feed(self._h, self._fp(losses), self._fp(grads))
        return rounds


class DeviceLockstepLBFGS:
    """LockstepLBFGS with the optimiser state on the DEVICE (csrc/lbfgs_device.cuh): the per-frame state machines run in a
    kernel right after the one-launch closure of `plan` (shooting.BatchedClosurePlan with plan.one_launch), one warp per frame,
    and the rounds of one optimizer.step() of all frames are ONE CUDA graph launch (a WHILE conditional node whose condition
    "some frame still waits for a closure value" is set on the device) -- no host round trip per round.  Same algorithm and
    settings as LockstepLBFGS; iterates agree with it to fp64 rounding of the dot products (other summation order)."""

    MAX_ROUNDS = 400                      # per optimizer.step(): max_eval = 100 closure values per frame leaves a wide margin

    def __init__(self, sizes, plan, max_iter=20, max_eval=100, history_size=100, tolerance_grad=1e-7, tolerance_change=1e-9):
        import numpy as np
        from .. import _lib
        self._np, self._libmod = np, _lib
        self._lib = _lib.load()
        self.plan = plan
        self.K = K = len(sizes)
        self.sizes = [int(n) for n in sizes]
        self.stride = int(plan.pstride)
        if max(self.sizes) > self.stride or K != plan.K:
            raise ValueError("DeviceLockstepLBFGS: sizes do not match the closure plan")
        dev = plan.device
        self.history = int(history_size)
        NI, ND, NV = 24, 24, 12
        ints = np.zeros((K, NI), np.int32)
        ints[:, 0] = self.sizes
        ints[:, 1] = 1
        dbl = np.zeros((K, ND), np.float64)
        dbl[:, 1] = 1.0
        dbl[:, 18] = np.nan
        dbl[:, 19] = np.inf
        self._ints0, self._dbl0 = torch.from_numpy(ints).to(dev), torch.from_numpy(dbl).to(dev)
        self.ints = self._ints0.clone()
        self.dbl = self._dbl0.clone()
        self.vec = torch.zeros(K, NV, self.stride, dtype=torch.float64, device=dev)
        self.best_x = torch.zeros(K, self.stride, dtype=torch.float32, device=dev)
        self.dirs = torch.zeros(K, self.history, self.stride, dtype=torch.float64, device=dev)
        self.stps = torch.zeros(K, self.history, self.stride, dtype=torch.float64, device=dev)
        self.ro = torch.zeros(K, self.history, dtype=torch.float64, device=dev)
        self.al = torch.zeros(K, self.history, dtype=torch.float64, device=dev)
        self.counters = torch.zeros(4, dtype=torch.int32, device=dev)
        self.mask = torch.zeros(K, dtype=torch.uint8, device=dev)
        self.h_counters = torch.zeros(4, dtype=torch.int32).pin_memory()
        L = _lib.LbfgsDev()
        L.K, L.stride, L.history = K, self.stride, self.history
        L.max_iter, L.max_eval, L.max_ls = int(max_iter), int(max_eval), 25
        L.tol_grad, L.tol_change, L.lr, L.c1, L.c2 = float(tolerance_grad), float(tolerance_change), 1.0, 1e-4, 0.9
        L.ints, L.dbl, L.vec, L.best_x = self.ints.data_ptr(), self.dbl.data_ptr(), self.vec.data_ptr(), self.best_x.data_ptr()
        L.dirs, L.stps, L.ro, L.al = self.dirs.data_ptr(), self.stps.data_ptr(), self.ro.data_ptr(), self.al.data_ptr()
        L.counters = self.counters.data_ptr()
        self.L = L
        self.graph = None
        self._cache, self._pin = None, None

    def __del__(self):
        g, self.graph = getattr(self, "graph", None), None
        if g:
            self._lib.dicp_lbfgs_dev_loop_destroy(g)

    def fresh(self):
        """Back to the state of a newly created optimiser (new LockstepLBFGS object in the host version)."""
        self.ints.copy_(self._ints0)
        self.dbl.copy_(self._dbl0)
        self._cache = None

    # ---- parameters / state ---------------------------------------------------------------------------------------------------
    def set_x(self, k, x):
        a = torch.from_numpy(self._np.ascontiguousarray(x, dtype=self._np.float32).reshape(-1).astype(self._np.float64))
        assert a.numel() == self.sizes[k]
        self.vec[k, 0, :self.sizes[k]] = a.to(self.vec.device)
        self._cache = None

    def set_all(self, xs):
        X = self._np.zeros((self.K, self.stride), self._np.float64)
        for k, x in enumerate(xs):
            X[k, :self.sizes[k]] = self._np.asarray(x, dtype=self._np.float32).reshape(-1)
        self.vec[:, 0, :] = torch.from_numpy(X).to(self.vec.device)
        self._cache = None

    def reset(self, k, line_search=True):
        # dicp_lbfgs_reset: fresh optimiser memory, counters and step size; best / last closure values persist
        idx = torch.tensor([1, 2, 6, 7, 17, 18], device=self.ints.device)
        self.ints[k, idx] = torch.tensor([int(bool(line_search)), 0, 0, 0, 0, 0], dtype=torch.int32, device=self.ints.device)
        self.dbl[k, :2] = torch.tensor([0.0, 1.0], dtype=torch.float64, device=self.dbl.device)
        self._cache = None

    def reset_all(self, line_search=True):
        idx = torch.tensor([1, 2, 6, 7, 17, 18], device=self.ints.device)
        self.ints[:, idx] = torch.tensor([int(bool(line_search)), 0, 0, 0, 0, 0], dtype=torch.int32, device=self.ints.device)
        self.dbl[:, :2] = torch.tensor([0.0, 1.0], dtype=torch.float64, device=self.dbl.device)
        self._cache = None

    def _fetch(self):
        """One read of everything the host driver looks at (state integers / scalars, current and best parameters): four
        asynchronous copies into pinned buffers and ONE synchronisation; cached until the state changes."""
        if self._cache is None:
            if self._pin is None:
                self._pin = (torch.zeros_like(self.ints, device="cpu").pin_memory(), torch.zeros_like(self.dbl, device="cpu").pin_memory(),
                             torch.zeros(self.K, self.stride, dtype=torch.float32).pin_memory(),
                             torch.zeros(self.K, self.stride, dtype=torch.float32).pin_memory())
            hi, hd, hx, hb = self._pin
            hi.copy_(self.ints, non_blocking=True)
            hd.copy_(self.dbl, non_blocking=True)
            hx.copy_(self.vec[:, 0, :].float(), non_blocking=True)
            hb.copy_(self.best_x, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self._cache = (hi.numpy().copy(), hd.numpy().copy(), hx.numpy().copy(), hb.numpy().copy())
        return self._cache

    def get_all(self, best=False):
        i, _, x, b = self._fetch()
        if best:
            if not (i[:, 16] != 0).all():
                raise RuntimeError("DeviceLockstepLBFGS.get_all: no value")
            return b
        return x

    def get_x(self, k, best=False):
        return self.get_all(best)[k, :self.sizes[k]].copy()

    def stats_all(self):
        i, d, _, _ = self._fetch()
        out = self._np.zeros((self.K, 4), self._np.float64)
        out[:, 0], out[:, 1], out[:, 2], out[:, 3] = d[:, 18], d[:, 19], i[:, 6], i[:, 7]
        return out

    def stats(self, k):
        s = self.stats_all()[k]
        return {"last": s[0], "best": s[1], "func_evals": int(s[2]), "n_iter": int(s[3])}

    # ---- one optimizer.step() of the selected frames ----------------------------------------------------------------------------
    def step(self, mask, evaluate=None, X=None, active=None, losses=None, grads=None):
        from ..shooting import ShootPlan
        from .._lib import check, stream_ptr
        np, plan, lib = self._np, self.plan, self._lib
        self._cache = None
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        self.mask.copy_(torch.from_numpy(m), non_blocking=False)
        st = stream_ptr()
        check(lib.dicp_lbfgs_dev_begin(self.L, self.mask.data_ptr(), plan.d_X.data_ptr(), plan.ostride, plan.d_active.data_ptr(), st),
              "dicp_lbfgs_dev_begin")
        if self.graph is None:
            # first step: the rounds one by one with a host read of the waiting count (this also loads the kernels); then the
            # WHILE graph is built for all later steps
            rounds = 0
            while True:
                check(lib.dicp_lbfgs_dev_round(self.L, *plan.cluster_args(), st), "dicp_lbfgs_dev_round")
                rounds += 1
                self.h_counters.copy_(self.counters)
                if int(self.h_counters[3]) == 0 or rounds >= self.MAX_ROUNDS:
                    break
            with ShootPlan._lock:
                self.graph = lib.dicp_lbfgs_dev_loop_create(self.L, *plan.cluster_args(), self.MAX_ROUNDS)
            if not self.graph:
                raise RuntimeError("dicp_lbfgs_dev_loop_create failed")
        else:
            check(lib.dicp_lbfgs_dev_loop_launch(self.graph, st), "dicp_lbfgs_dev_loop_launch")
            self.h_counters.copy_(self.counters)               # synchronises
            rounds = int(self.h_counters[2])
        if int(self.h_counters[3]) != 0:
            raise RuntimeError("DeviceLockstepLBFGS.step: frames still waiting after MAX_ROUNDS rounds")
        plan.evaluations += rounds
        return rounds


def LBFGS_optimization_lockstep(p0, evaluator, nmax=10, tol=1e-3, errthresh=1e8):
    """`LBFGS_optimization` (tools/optim.py:10-110) for K independent problems advanced in lock step.

    p0: list of K host arrays (any shape).  evaluator: object with numpy buffers `X` (K, stride) fp32, `active` (K,)
    uint8, `losses` (K,) fp32, `grads` (K, stride) fp32 and a method `evaluate()` that fills losses / grads for the active
    frames.  Per frame, same control flow as the sequential driver: at most nmax optimiser steps, stop when the RMS change
    is below tol x RMS, divergence guard with fall-back to the best point / a 1 % perturbation and a restart without line
    search, and the BEST parameters over all closure evaluations are returned.
    Returns (list of best parameter arrays, list of best losses, list of step counts, list of changes, rounds)."""
    import numpy as np
    K = len(p0)
    shapes = [np.shape(p) for p in p0]
    sizes = [int(np.prod(s)) for s in shapes]
    if getattr(evaluator, "device_lbfgs", False):
        # optimiser state on the device, rounds of a step as one CUDA graph launch (csrc/lbfgs_device.cuh)
        opt = evaluator.device_optimizer(sizes)         # fresh state: line search on, empty memory
        opt.set_all(p0)
    else:
        opt = LockstepLBFGS(sizes, stride=evaluator.X.shape[1])
        for k in range(K):
            opt.set_x(k, np.asarray(p0[k], dtype=np.float32))
            opt.reset(k, True)
    steps = [0] * K
    go_on = [True] * K
    L = [math.inf] * K
    change = [None] * K
    rounds = 0
    while True:
        mask = np.array([1 if (go_on[k] and steps[k] < nmax) else 0 for k in range(K)], dtype=np.uint8)
        if not mask.any():
            break
        before = opt.get_all()
        rounds += opt.step(mask, evaluator.evaluate, evaluator.X, evaluator.active, evaluator.losses, evaluator.grads)
        redo = []
        st_all, now_all = opt.stats_all(), opt.get_all()
        # RMS change and RMS size of every frame's parameters in one go (rows are zero beyond the frame's size)
        szs = np.asarray(sizes, dtype=np.float64)
        delta_all = np.sqrt(((now_all.astype(np.float32) - before.astype(np.float32)) ** 2).sum(1, dtype=np.float32) / szs)
        scale_all = np.sqrt((before.astype(np.float32) ** 2).sum(1, dtype=np.float32) / szs)
        for k in np.flatnonzero(mask):
            k = int(k)
            steps[k] += 1
            st = {"last": float(st_all[k, 0]), "best": float(st_all[k, 1])}
            L_before, L[k] = L[k], st["last"]
            if L[k] > L_before or L[k] > errthresh or math.isnan(L[k]):
                if math.isnan(L[k]):
                    print("WARNING: NaN value for loss L during L-BFGS optimization.")
                elif L[k] > errthresh:
                    print("WARNING: Aberrantly large value for loss L during L-BFGS optimization.")
                else:
                    print("WARNING: Increase of loss L during L-BGFS optimization.")
                best = opt.get_x(k, best=True)
                if st["best"] < L_before:
                    opt.set_x(k, best)
                    L[k] = st["best"]
                    print("L-BFGS optimization. Found an intermediate 'best_p' value for this iteration.")
                else:
                    rmod = 0.01
                    t = torch.from_numpy(best)
                    opt.set_x(k, (t + rmod * t.std() * torch.randn(t.shape)).numpy())
                    redo.append(k)
                    print(f"L-BFGS optimization. Trying a random perturbation of parameter from its current value, with relative strength {rmod}.")
                change[k] = "None (divergent iteration step)"
                opt.reset(k, False)
            else:
                delta, scale = float(delta_all[k]), float(scale_all[k])
                go_on[k] = delta > tol * scale
                change[k] = delta
        if redo:                                   # loss at the perturbed points (tools/optim.py:73), outside the optimisers
            evaluator.active[:] = 0
            for k in redo:
                evaluator.active[k] = 1
                evaluator.X[k, :sizes[k]] = opt.get_x(k)
            evaluator.evaluate()
            rounds += 1
            for k in redo:
                L[k] = float(evaluator.losses[k])
    best_all, st_all = opt.get_all(best=True), opt.stats_all()
    best_p = [best_all[k, :sizes[k]].reshape(shapes[k]) for k in range(K)]
    best_L = [float(st_all[k, 1]) for k in range(K)]
    return best_p, best_L, steps, change, rounds
