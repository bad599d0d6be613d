"""Geodesic shooting on the device: fused right-hand side + Euler / Ralston steps + hand-written discrete adjoint.

Replaces the Python integrator loop of the reference (tools/integrators.py:20-51) driving LDDMMModel.ODE
(core/LDDMM.py:176-227), and the reverse-mode autograd pass through that loop (tools/optim.py:34-47), by
back-to-back launches of the fused kernels (dicp_rhs_forward / dicp_rhs_adjoint / dicp_axpy) on the current
stream.  The adjoint is the exact transpose of the discrete Euler / Ralston step sequence, so gradients match
"autograd through the loop" to rounding.  With `use_graph` the whole launch sequence of one shoot (and of one
adjoint sweep) is captured once per problem shape into a CUDA graph and replayed (north_star item 3).

State layout (one flat fp32 vector of S = 2*M*D + Nx*D + 1 floats):   [ q (M,D) | p (M,D) | x (Nx,D) | cost ]
Right-hand-side buffers carry 3 extra floats: [ vq | dp | vx | dcost | A | B | C ]  (A,B,C: Hamiltonian pieces).
"""

from __future__ import annotations

import torch

import os
import threading

from . import ops
from . import _lib

# Concurrent registration of several frames (DiffPSR.Reg_opt with frame_workers > 1) runs one Python thread + one CUDA
# stream per worker; each worker uses its own plan instances (buffers, captured graphs), selected by this slot.
_tls = threading.local()


def set_plan_slot(slot):
    _tls.slot = slot


def _plan_slot():
    return getattr(_tls, "slot", 0)


class ShootSpec:
    """Static description of one shooting problem (shapes + model), hashable: the key of the graph cache."""

    __slots__ = ("D", "M", "Nx", "nt", "scheme", "withlogdet", "sigma", "eta", "device")

    def __init__(self, D, M, Nx, nt, scheme, withlogdet, sigma, eta, device):
        self.D, self.M, self.Nx, self.nt = int(D), int(M), int(Nx), int(nt)
        self.scheme, self.withlogdet = scheme, bool(withlogdet)
        self.sigma, self.eta, self.device = float(sigma), float(eta), device

    def key(self):
        return (self.D, self.M, self.Nx, self.nt, self.scheme, self.withlogdet, self.sigma, self.eta, str(self.device))

    @property
    def S(self):
        return 2 * self.M * self.D + self.Nx * self.D + 1


def _views(spec, flat):
    """(q, p, x, cost) views of a flat state / cotangent vector."""
    MD = spec.M * spec.D
    q = flat[0:MD].view(spec.M, spec.D)
    p = flat[MD:2 * MD].view(spec.M, spec.D)
    x = flat[2 * MD:2 * MD + spec.Nx * spec.D].view(spec.Nx, spec.D) if spec.Nx else None
    cost = flat[spec.S - 1:spec.S]
    return q, p, x, cost


def _rhs(spec, state, F, ws):
    q, p, x, _ = _views(spec, state)
    vq, dp, vx, _ = _views(spec, F)
    ops.rhs_forward(spec.D, spec.withlogdet, spec.sigma, spec.eta, q, p, x, vq, dp, vx, F[spec.S - 1:], ws)


def _vjp(spec, state, lam, G, ws):
    """G = J_F(state)^T lam  (G's cost entry stays 0: the right-hand side does not depend on cost)."""
    q, p, x, _ = _views(spec, state)
    a, u, wx, gc = _views(spec, lam)
    gq, gp, gx, _ = _views(spec, G)
    ops.rhs_adjoint(spec.D, spec.withlogdet, spec.sigma, spec.eta, q, p, x, a, u, wx, gc, gq, gp, gx, ws)


def forward_sweep(spec, traj, mid, F1, F2, F0, ws, sws=None):
    """traj[0] holds the initial state; fills traj[1..nt] (and mid[0..nt-1] for Ralston), F0 = rhs(traj[0]).
    With a small support (sws given) every stage is ONE fused launch (csrc/small_step.cuh)."""
    h = 1.0 / spec.nt
    S = spec.S
    if sws is not None:
        a = (spec.D, spec.withlogdet, spec.sigma, spec.eta, spec.M, spec.Nx)
        for t in range(spec.nt):
            Fa = F0 if t == 0 else F1
            if spec.scheme == "Euler":
                ops.small_rhs_step(*a, traj[t], traj[t], None, h, 0.0, traj[t + 1], Fa, sws)
            else:
                ops.small_rhs_step(*a, traj[t], traj[t], None, 2.0 * h / 3.0, 0.0, mid[t], Fa, sws)
                ops.small_rhs_step(*a, mid[t], traj[t], Fa, 0.75 * h, 0.25 * h, traj[t + 1], F2, sws)
        return
    for t in range(spec.nt):
        Fa = F0 if t == 0 else F1
        _rhs(spec, traj[t], Fa, ws)
        if spec.scheme == "Euler":
            ops.axpy(traj[t + 1], traj[t], h, Fa, n=S)
        else:
            ops.axpy(mid[t], traj[t], 2.0 * h / 3.0, Fa, n=S)
            _rhs(spec, mid[t], F2, ws)
            ops.axpy(traj[t + 1], traj[t], 0.25 * h, Fa, 0.75 * h, F2, n=S)


def adjoint_sweep(spec, traj, mid, gtraj, lam, mu, G1, G2, ws, sws=None, lam2=None):
    """lam <- d(loss)/d(traj[0]) given gtraj[t] = d(loss)/d(traj[t]) (direct dependence on every stored state).
    With a small support (sws, lam2 given) every stage is ONE fused launch; lam / lam2 ping-pong so that the result
    lands in lam."""
    h = 1.0 / spec.nt
    S = spec.S
    if sws is not None:
        a = (spec.D, spec.withlogdet, spec.sigma, spec.eta, spec.M, spec.Nx)
        cur, nxt = (lam, lam2) if spec.nt % 2 == 0 else (lam2, lam)
        cur.copy_(gtraj[spec.nt])
        for t in range(spec.nt - 1, -1, -1):
            if spec.scheme == "Euler":
                ops.small_adj_step(*a, traj[t], cur, cur, None, gtraj[t], h, 0.0, nxt, G1, sws)
            else:
                ops.small_adj_step(*a, mid[t], cur, cur, None, None, 2.0 * h, 0.0, mu, G2, sws)
                ops.small_adj_step(*a, traj[t], mu, cur, G2, gtraj[t], 0.25 * h, 0.75 * h, nxt, G1, sws)
            cur, nxt = nxt, cur
        return
    lam.copy_(gtraj[spec.nt])
    for t in range(spec.nt - 1, -1, -1):
        if spec.scheme == "Euler":
            _vjp(spec, traj[t], lam, G1, ws)
            ops.axpy(lam, lam, h, G1, 1.0, gtraj[t], n=S)
        else:
            _vjp(spec, mid[t], lam, G2, ws)
            ops.axpy(mu, lam, 2.0 * h, G2, n=S)
            _vjp(spec, traj[t], mu, G1, ws)
            ops.axpy(lam, lam, 0.75 * h, G2, 0.25 * h, G1, n=S)
            ops.axpy(lam, lam, 1.0, gtraj[t], n=S)


class _PlanCache:
    """LRU cache of plans bounded by the BYTES of device memory they hold (a plan of a dense 1M-point frame keeps ~1 GB of
    trajectory / adjoint buffers; a count-bound cache could exhaust the GPU with ragged large frames)."""

    def __init__(self, max_bytes):
        from collections import OrderedDict
        self.max_bytes = int(max_bytes)
        self.items = OrderedDict()
        self.bytes = 0

    def get(self, key):
        plan = self.items.get(key)
        if plan is not None:
            self.items.move_to_end(key)
        return plan

    def __len__(self):
        return len(self.items)

    def __setitem__(self, key, plan):
        self.items[key] = plan
        self.bytes += plan.nbytes
        while self.bytes > self.max_bytes and len(self.items) > 1:
            _, old = self.items.popitem(last=False)
            self.bytes -= old.nbytes

    def clear(self):
        self.items.clear()
        self.bytes = 0


def _tensor_bytes(obj):
    return sum(t.numel() * t.element_size() for t in vars(obj).values() if isinstance(t, torch.Tensor) and t.is_cuda)


PLAN_CACHE_BYTES = int(float(os.environ.get("DICP_PLAN_CACHE_GB", "8")) * (1 << 30))


class ShootPlan:
    """Buffers (and, optionally, captured CUDA graphs) for one ShootSpec. Cached per spec (LRU, bounded by bytes)."""

    _cache = _PlanCache(PLAN_CACHE_BYTES)

    def __init__(self, spec: ShootSpec, use_graph: bool):
        self.spec = spec
        dev = spec.device
        S, nt = spec.S, spec.nt
        f32 = dict(dtype=torch.float32, device=dev)
        self.traj = torch.zeros(nt + 1, S, **f32)
        self.mid = torch.zeros(nt, S, **f32) if spec.scheme == "Ralston" else None
        self.F0 = torch.zeros(S + 3, **f32)
        self.F1 = torch.zeros(S + 3, **f32)
        self.F2 = torch.zeros(S + 3, **f32)
        self.gtraj = torch.zeros(nt + 1, S, **f32)
        self.lam = torch.zeros(S, **f32)
        self.mu = torch.zeros(S, **f32)
        self.G1 = torch.zeros(S, **f32)
        self.G2 = torch.zeros(S, **f32)
        rows = max(spec.M, spec.Nx)
        self.small = ops.use_small_path(spec.M)
        if self.small:
            self.sws = ops.alloc_small_workspace(spec.M, spec.Nx, dev)
            self.lam2 = torch.zeros(S, **f32)
            self.ws = torch.zeros(8192, dtype=torch.uint8, device=dev)       # quad-loss partials only
        else:
            self.sws, self.lam2 = None, None
            self.ws = ops.alloc_workspace(rows, rows, dev)
        self.version = 0
        self.use_graph = use_graph
        self.fwd_graph = None
        self.bwd_graph = None
        self.nbytes = _tensor_bytes(self)

    _lock = threading.Lock()

    @classmethod
    def get(cls, spec: ShootSpec, use_graph: bool, slot=None):
        key = spec.key() + (bool(use_graph), _plan_slot() if slot is None else slot)
        plan = cls._cache.get(key)
        if plan is None:
            with cls._lock:
                plan = cls._cache.get(key)
                if plan is None:
                    plan = cls(spec, use_graph)
                    cls._cache[key] = plan
        return plan

    def ensure_captured(self):
        """Capture the forward and adjoint graphs now (on the calling thread / current stream), so that later replays
        from worker threads never capture concurrently."""
        if not self.use_graph:
            return
        if self.fwd_graph is None:
            self.traj[0].zero_()
            self._capture_forward()
        if self.bwd_graph is None:
            self.gtraj.zero_()
            self._capture_backward()

    def _capture_forward(self):
        self._fwd_body()                      # warm-up outside capture (lazy init, occupancy queries)
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self._fwd_body()
        self.fwd_graph = g

    def _capture_backward(self):
        self._bwd_body()
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self._bwd_body()
        self.bwd_graph = g

    # -- forward ---------------------------------------------------------------------------------------
    def _fwd_body(self):
        forward_sweep(self.spec, self.traj, self.mid, self.F1, self.F2, self.F0, self.ws, self.sws)

    def run_forward(self, q0, p0, x0):
        spec = self.spec
        q, p, x, cost = _views(spec, self.traj[0])
        q.copy_(q0)
        p.copy_(p0)
        if x is not None:
            x.copy_(x0)
        cost.zero_()
        if self.use_graph:
            if self.fwd_graph is None:
                with ShootPlan._lock:
                    self._capture_forward()
            self.fwd_graph.replay()
        else:
            self._fwd_body()
        self.version += 1

    # -- backward --------------------------------------------------------------------------------------
    def _bwd_body(self):
        adjoint_sweep(self.spec, self.traj, self.mid, self.gtraj, self.lam, self.mu, self.G1, self.G2, self.ws,
                      self.sws, self.lam2)

    def run_backward(self, gtraj):
        self.gtraj.copy_(gtraj)
        if self.use_graph:
            if self.bwd_graph is None:
                with ShootPlan._lock:
                    self._capture_backward()
            self.bwd_graph.replay()
        else:
            self._bwd_body()


class ClosurePlan:
    """One L-BFGS closure of the registration step -- shoot, trajectory loss lambda*H(q0,p0) + cost(1), quadratic data
    loss sum_n inv_n |x_n(1) - y_n|^2 (core/LDDMM.py:363-371 with core/PSR.py:498-516) and the adjoint sweep -- as ONE
    launch sequence on static buffers, replayed as a single CUDA graph.  Output: one contiguous buffer
    [dcost(0), A, B, C, cost(1), data loss, -, - | d loss / d p0 (M*D)] so that the host reads loss and gradient with one
    device-to-host copy (the reference synchronises once per closure too, tools/optim.py:39)."""

    _cache = _PlanCache(PLAN_CACHE_BYTES)
    NS = 8

    def __init__(self, spec: ShootSpec, use_graph: bool, lam_reg: float):
        self.spec, self.lam_reg, self.use_graph = spec, float(lam_reg), use_graph
        self.plan = ShootPlan(spec, False)
        dev = spec.device
        self.n_data = spec.Nx if spec.Nx else spec.M
        f32 = dict(dtype=torch.float32, device=dev)
        self.y = torch.zeros(self.n_data, spec.D, **f32)
        self.inv = torch.zeros(self.n_data, **f32)
        self.out = torch.zeros(self.NS + spec.M * spec.D, **f32)
        self.out_host = torch.zeros(self.NS + spec.M * spec.D, dtype=torch.float32).pin_memory() if dev.type == "cuda" \
            else torch.zeros(self.NS + spec.M * spec.D, dtype=torch.float32)
        self.plan.gtraj[spec.nt][spec.S - 1] = 1.0            # d loss / d cost(1) = 1
        self.graph = None
        self.nbytes = self.plan.nbytes + _tensor_bytes(self)

    @classmethod
    def get(cls, spec, use_graph, lam_reg):
        key = spec.key() + (bool(use_graph), float(lam_reg), _plan_slot())
        cp = cls._cache.get(key)
        if cp is None:
            with ShootPlan._lock:
                cp = cls._cache.get(key)
                if cp is None:
                    cp = cls(spec, use_graph, lam_reg)
                    cls._cache[key] = cp
        return cp

    def set_problem(self, q0, x0, y, inv):
        spec = self.spec
        q, _, x, cost = _views(spec, self.plan.traj[0])
        q.copy_(q0)
        if x is not None:
            x.copy_(x0)
        cost.zero_()
        self.y.copy_(y)
        self.inv.copy_(inv)

    def _body(self):
        spec, plan = self.spec, self.plan
        S, MD = spec.S, spec.M * spec.D
        forward_sweep(spec, plan.traj, plan.mid, plan.F1, plan.F2, plan.F0, plan.ws, plan.sws)
        end = plan.traj[spec.nt]
        q1, _, x1, _ = _views(spec, end)
        gq, _, gx, _ = _views(spec, plan.gtraj[spec.nt])
        ops.quad_loss(x1 if spec.Nx else q1, self.y, self.inv, gx if spec.Nx else gq, self.out[5:6], plan.ws)
        adjoint_sweep(spec, plan.traj, plan.mid, plan.gtraj, plan.lam, plan.mu, plan.G1, plan.G2, plan.ws,
                      plan.sws, plan.lam2)
        # d/dp0 [lambda H(q0,p0)] = lambda vq(0)   (Hamilton's equations, core/LDDMM.py:156-158)
        ops.axpy(self.out[self.NS:], plan.lam[MD:2 * MD], self.lam_reg, plan.F0[0:MD], n=MD)
        self.out[0:4].copy_(plan.F0[S - 1:S + 3])
        self.out[4:5].copy_(end[S - 1:S])

    def evaluate(self, p0):
        """p0: (M,D) tensor on any device.  Returns (loss as a Python float, gradient view into the host buffer)."""
        spec = self.spec
        _, p, _, _ = _views(spec, self.plan.traj[0])
        p.copy_(p0)
        if self.use_graph:
            if self.graph is None:
                with ShootPlan._lock:
                    self._body()
                    torch.cuda.current_stream().synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        self._body()
                    self.graph = g
            self.graph.replay()
        else:
            self._body()
        self.out_host.copy_(self.out)                        # the one host synchronisation of this closure
        s = self.out_host[:self.NS].tolist()
        H0 = 0.5 * s[1] - spec.eta * s[2] - 0.5 * spec.eta ** 2 * s[3]
        loss = self.lam_reg * H0 + s[4] + s[5]
        return loss, self.out_host[self.NS:].view(spec.M, spec.D)


class _ShootFn(torch.autograd.Function):
    """(q0, p0, x0) -> (trajectory (nt+1, S), H(q0,p0) as a 0-d tensor)."""

    @staticmethod
    def forward(ctx, spec, use_graph, q0, p0, x0):
        plan = ShootPlan.get(spec, use_graph)
        plan.run_forward(q0, p0, x0)
        traj = plan.traj.clone()
        hs = plan.F0[spec.S - 1:spec.S + 3]                  # dcost(0), A, B, C  (core/LDDMM.py:150-155)
        H0 = 0.5 * hs[1] - spec.eta * hs[2] - (0.5 * spec.eta ** 2) * hs[3]
        ctx.spec, ctx.plan, ctx.version = spec, plan, plan.version
        need = any(t is not None and t.requires_grad for t in (q0, p0, x0))
        if need:
            ctx.saved_mid = plan.mid.clone() if plan.mid is not None else None
            ctx.F0 = plan.F0[:spec.S].clone()
            ctx.save_for_backward(traj)
        ctx.has_x = x0 is not None
        return traj, H0

    @staticmethod
    def backward(ctx, g_traj, g_H):
        spec, plan = ctx.spec, ctx.plan
        (traj,) = ctx.saved_tensors
        if plan.version != ctx.version:                       # plan buffers were reused by another shoot
            plan.traj.copy_(traj)
            if plan.mid is not None:
                plan.mid.copy_(ctx.saved_mid)
            plan.version += 1
            ctx.version = plan.version
        if g_traj is None:
            g_traj = torch.zeros_like(traj)
        plan.run_backward(g_traj.contiguous())
        lam = plan.lam.clone()
        gq, gp, gx, _ = _views(spec, lam)
        if g_H is not None:
            # Hamilton's equations: dH/dp = vq(0), dH/dq = Gq(0) = -dp(0)  (core/LDDMM.py:156-158), both
            # already produced by the first right-hand-side evaluation of the forward sweep.
            vq0, dp0, _, _ = _views(spec, ctx.F0)
            gp.add_(g_H * vq0)
            gq.sub_(g_H * dp0)
        return None, None, gq, gp, (gx if ctx.has_x else None)


class LazyStates:
    """The reference's "shoot" variable -- a list of nt+1 tuples (q, p, cost[, x]) -- whose tuples are views of the
    trajectory tensor created on first access (most callers only read shoot[-1] and shoot[0])."""

    def __init__(self, spec, traj, has_x):
        self._spec, self._traj, self._has_x = spec, traj, has_x
        self._cache = {}

    def __len__(self):
        return self._spec.nt + 1

    def __getitem__(self, t):
        if isinstance(t, slice):
            return [self[i] for i in range(*t.indices(len(self)))]
        n = len(self)
        if t < 0:
            t += n
        if not 0 <= t < n:
            raise IndexError("shoot index out of range")
        st = self._cache.get(t)
        if st is None:
            q, p, x, cost = _views(self._spec, self._traj[t])
            st = (q, p, cost, x) if self._has_x else (q, p, cost)
            self._cache[t] = st
        return st

    def __iter__(self):
        return (self[t] for t in range(len(self)))


def shoot(spec: ShootSpec, q0, p0, x0=None, use_graph=False):
    """Returns (states, H(q0,p0)): `states` behaves like the reference's list of nt+1 state tuples, with autograd
    attached to every entry."""
    _lib.require_cuda(q0, p0, x0)
    q0c, p0c = q0.contiguous(), p0.contiguous()
    x0c = None if x0 is None else x0.contiguous()
    traj, H0 = _ShootFn.apply(spec, use_graph, q0c, p0c, x0c)
    return LazyStates(spec, traj, x0 is not None), H0


class BatchedClosurePlan:
    """The L-BFGS closures of K independent frames -- shoot, lambda*H(q0,p0) + cost(1), quadratic data loss, adjoint sweep
    (what ClosurePlan does for one frame) -- evaluated by ONE launch sequence with the frame index on blockIdx.y, replayed
    as a single CUDA graph that also contains the host->device copy of the trial momenta and the device->host copy of
    losses and gradients.  Small supports only (M_k <= dicp_small_max_support()).  It is the `evaluator` of
    tools.optim.LBFGS_optimization_lockstep: numpy views `X`, `active`, `losses`, `grads` of its pinned host buffers.

    Frame k's buffers are rows of (K, fstride) arrays, with the frame's own sizes (ragged frames are fine).
    Every frame's arithmetic is independent of which other frames are active, and deterministic."""

    NS = 8
    one_launch_closure = True          # class-level switch (tests compare the two forms)
    device_lbfgs_enabled = True        # with the one-launch closure: L-BFGS state machines on the device (tools/optim.py)

    def __init__(self, D, nt, scheme, withlogdet, sigma, eta, lam_reg, device, Ms, Nxs, use_graph=True):
        import numpy as np
        self.D, self.nt, self.scheme, self.withlogdet = int(D), int(nt), scheme, bool(withlogdet)
        self.sigma, self.eta, self.lam_reg, self.device = float(sigma), float(eta), float(lam_reg), device
        self.Ms, self.Nxs = [int(m) for m in Ms], [int(n) for n in Nxs]
        self.K = K = len(self.Ms)
        self.maxM, self.maxNx = max(self.Ms), max(self.Nxs)
        if self.maxM > _lib.load().dicp_small_max_support():
            raise ValueError("BatchedClosurePlan: support sets are too large for the small-support kernels")
        self.specs = [ShootSpec(D, m, n, nt, scheme, withlogdet, sigma, eta, device) for m, n in zip(self.Ms, self.Nxs)]
        self.ndata = [n if n > 0 else m for m, n in zip(self.Ms, self.Nxs)]
        self.maxNd = max(self.ndata)
        self.fstride = (2 * self.maxM * D + self.maxNx * D + 4 + 3) // 4 * 4
        self.pstride = (self.maxM * D + 3) // 4 * 4
        self.ostride = self.NS + self.pstride
        f32 = dict(dtype=torch.float32, device=device)
        fs = self.fstride
        self.traj = torch.zeros(nt + 1, K, fs, **f32)
        self.mid = torch.zeros(nt, K, fs, **f32) if scheme == "Ralston" else None
        self.F0, self.F1, self.F2 = (torch.zeros(K, fs, **f32) for _ in range(3))
        self.gend, self.lamA, self.lamB, self.mu, self.G1, self.G2 = (torch.zeros(K, fs, **f32) for _ in range(6))
        self.dims = torch.tensor([[m, n] for m, n in zip(self.Ms, self.Nxs)], dtype=torch.int32).to(device)
        # d loss / d cost(1) = 1
        idx = torch.tensor([k * fs + sp.S - 1 for k, sp in enumerate(self.specs)], dtype=torch.long, device=device)
        self.gend.view(-1)[idx] = 1.0
        self.y = torch.zeros(K, self.maxNd, D, **f32)
        self.inv = torch.zeros(K, self.maxNd, **f32)
        # destination rows of the frames' data points in the padded (K * maxNd) layout
        self.rows = torch.cat([k * self.maxNd + torch.arange(n) for k, n in enumerate(self.ndata)]).to(device)
        self.ws_frame = ops.batch_frame_ws_bytes(self.maxM, self.maxNx)
        self.ws = torch.zeros(K * self.ws_frame, dtype=torch.uint8, device=device)
        self.qws = torch.zeros(int(_lib.load().dicp_batch_quad_workspace_bytes(K)), dtype=torch.uint8, device=device)
        self.counts = torch.zeros(K, nt + 1, dtype=torch.int32, device=device)
        # host <-> device staging: [ active (K int32, stored in a float32 buffer) | X (K, ostride) ] and out (K, ostride)
        pin = device.type == "cuda"
        self.h_in = torch.zeros(K + K * self.ostride, dtype=torch.float32)
        self.h_out = torch.zeros(K, self.ostride, dtype=torch.float32)
        if pin:
            self.h_in, self.h_out = self.h_in.pin_memory(), self.h_out.pin_memory()
        self.d_in = torch.zeros(K + K * self.ostride, **f32)
        self.d_out = torch.zeros(K, self.ostride, **f32)
        self.d_active = self.d_in[:K].view(torch.int32)
        self.d_X = self.d_in[K:].view(K, self.ostride)
        self._h_active = self.h_in[:K].view(torch.int32).numpy()
        self.X = self.h_in[K:].view(K, self.ostride).numpy()
        self.active = np.zeros(K, dtype=np.uint8)
        self.losses = np.zeros(K, dtype=np.float32)
        self._out_np = self.h_out.numpy()
        self.grads = self._out_np.reshape(-1)[self.NS:]
        self.use_graph = bool(use_graph) and device.type == "cuda"
        self.graph = None
        self.evaluations = 0
        # the whole closure of every frame in ONE launch (one thread-block cluster per frame, csrc/cluster_closure.cuh) when
        # the model / sizes allow it: eta = 0, data points present, Euler, small supports
        self.one_launch = (BatchedClosurePlan.one_launch_closure and device.type == "cuda" and min(self.Nxs) > 0
                           and ops.batch_closure_cluster_rows(D, eta, scheme, self.maxM, self.maxNx, nt, K) > 0)

    @property
    def device_lbfgs(self):
        return bool(self.one_launch and BatchedClosurePlan.device_lbfgs_enabled)

    def device_optimizer(self, sizes):
        """A DeviceLockstepLBFGS bound to this plan's buffers, with fresh state (its device buffers and its captured graph are
        reused from call to call; the state arrays are re-initialised)."""
        from .tools.optim import DeviceLockstepLBFGS
        opt = getattr(self, "_dev_opt", None)
        if opt is None or opt.sizes != [int(n) for n in sizes]:
            opt = DeviceLockstepLBFGS(sizes, self)
            self._dev_opt = opt
        else:
            opt.fresh()
        return opt

    # ---- problem data -------------------------------------------------------------------------------------------------
    def set_geometry(self, q0_list, x0_list):
        """Support points and data points of every frame (constant over the outer iterations of DiffPSR)."""
        t0 = self.traj[0]
        for k, sp in enumerate(self.specs):
            q, _, x, _ = _views(sp, t0[k])
            q.copy_(q0_list[k])
            if x is not None:
                x.copy_(x0_list[k])

    def set_targets(self, y_cat, inv_cat):
        """y_cat (sum_k n_k, D), inv_cat (sum_k n_k): targets / weights 1/(2 sigma_s^2) of all frames' data points,
        concatenated in frame order."""
        self.y.view(-1, self.D).index_copy_(0, self.rows, y_cat)
        self.inv.view(-1).index_copy_(0, self.rows, inv_cat)

    # ---- launch sequences ---------------------------------------------------------------------------------------------
    def _args(self):
        return (self.D, self.withlogdet, self.sigma, self.eta, self.K, self.dims, self.d_active, self.maxM, self.maxNx,
                self.fstride)

    def _forward(self):
        a, h, nt = self._args(), 1.0 / self.nt, self.nt
        ops.batch_set_p(self.D, self.K, self.dims, self.d_active, self.maxM, self.fstride, self.d_X, self.ostride,
                        self.traj[0])
        for t in range(nt):
            Fa = self.F0 if t == 0 else self.F1
            if self.scheme == "Euler":
                ops.batch_rhs_step(*a, self.traj[t], self.traj[t], None, h, 0.0, self.traj[t + 1], Fa, self.ws, self.ws_frame)
            else:
                ops.batch_rhs_step(*a, self.traj[t], self.traj[t], None, 2.0 * h / 3.0, 0.0, self.mid[t], Fa, self.ws,
                                   self.ws_frame)
                ops.batch_rhs_step(*a, self.mid[t], self.traj[t], Fa, 0.75 * h, 0.25 * h, self.traj[t + 1], self.F2,
                                   self.ws, self.ws_frame)
        ops.batch_quad_loss(self.D, self.K, self.dims, self.d_active, self.maxNd, self.fstride, self.traj[nt], self.y,
                            self.inv, self.maxNd, self.gend, self.d_out[:, 5], self.ostride, self.qws)

    def _adjoint(self):
        a, h, nt = self._args(), 1.0 / self.nt, self.nt
        cur = self.gend
        for t in range(nt - 1, -1, -1):
            nxt = self.lamA if cur is not self.lamA else self.lamB
            if self.scheme == "Euler":
                ops.batch_adj_step(*a, self.traj[t], cur, cur, None, None, h, 0.0, nxt, self.G1, self.ws, self.ws_frame)
            else:
                ops.batch_adj_step(*a, self.mid[t], cur, cur, None, None, 2.0 * h, 0.0, self.mu, self.G2, self.ws,
                                   self.ws_frame)
                ops.batch_adj_step(*a, self.traj[t], self.mu, cur, self.G2, None, 0.25 * h, 0.75 * h, nxt, self.G1,
                                   self.ws, self.ws_frame)
            cur = nxt
        return cur

    def cluster_args(self):
        """Arguments of dicp_batch_closure_cluster up to `nscal` (also the closure part of dicp_lbfgs_dev_round / _loop_create)."""
        return (self.D, int(self.withlogdet), float(self.sigma), float(self.eta), self.K, ops.ptr(self.dims), ops.ptr(self.d_active),
                self.maxM, self.maxNx, self.fstride, self.nt, ops.ptr(self.traj), self.K * self.fstride, ops.ptr(self.d_X),
                self.ostride, ops.ptr(self.y), ops.ptr(self.inv), self.maxNd, float(self.lam_reg), ops.ptr(self.d_out), self.ostride,
                self.NS)

    def _body(self):
        self.d_in.copy_(self.h_in, non_blocking=True)
        if self.one_launch:
            ops.batch_closure_cluster(self.D, self.withlogdet, self.sigma, self.eta, self.K, self.dims, self.d_active,
                                      self.maxM, self.maxNx, self.fstride, self.nt, self.traj, self.K * self.fstride,
                                      self.d_X, self.ostride, self.y, self.inv, self.maxNd, self.lam_reg, self.d_out,
                                      self.ostride, self.NS)
            self.h_out.copy_(self.d_out, non_blocking=True)
            return
        self._forward()
        lam = self._adjoint()
        ops.batch_closure_out(self.D, self.K, self.dims, self.d_active, self.maxM, self.fstride, self.lam_reg, lam,
                              self.F0, self.traj[self.nt], self.d_out, self.ostride, self.NS)
        self.h_out.copy_(self.d_out, non_blocking=True)

    def _losses(self):
        s = self._out_np[:, :self.NS].astype("float64")
        H0 = 0.5 * s[:, 1] - self.eta * s[:, 2] - 0.5 * self.eta ** 2 * s[:, 3]
        return self.lam_reg * H0 + s[:, 4], s[:, 5]

    def evaluate(self):
        """Closure values of the frames flagged in `active` at the rows of `X` -> `losses`, `grads` (one host
        synchronisation for all frames)."""
        self._h_active[:] = self.active
        if self.use_graph:
            if self.graph is None:
                with ShootPlan._lock:
                    self._body()
                    torch.cuda.current_stream().synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        self._body()
                    self.graph = g
            self.graph.replay()
        else:
            self._body()
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        trajl, datal = self._losses()
        self.losses[:] = trajl + datal
        self.evaluations += 1

    def finalize(self, p_list, coverage_radius=None):
        """Shoot every frame from the given momenta (forward only).  Returns (traj, trajloss (K,), dataloss (K,), counts):
        traj is a private (nt+1, K, fstride) copy of the trajectories; counts (K, nt+1) numbers of data points farther
        than coverage_radius from every support point at each stored time (None without data points / radius)."""
        for k, p in enumerate(p_list):
            self.X[k, :self.Ms[k] * self.D] = p.reshape(-1)
        self.active[:] = 1
        self._h_active[:] = 1
        self.d_in.copy_(self.h_in, non_blocking=True)
        self._forward()
        ops.batch_closure_out(self.D, self.K, self.dims, self.d_active, self.maxM, self.fstride, self.lam_reg, None,
                              self.F0, self.traj[self.nt], self.d_out, self.ostride, self.NS)
        counts = None
        if coverage_radius is not None and self.maxNx > 0:
            self.counts.zero_()
            ops.batch_coverage(self.D, self.K, self.dims, self.d_active, self.maxM, self.maxNx, self.fstride, self.traj,
                               self.K * self.fstride, self.nt + 1, coverage_radius, self.counts)
            counts = self.counts.cpu().numpy()             # synchronises
        self.h_out.copy_(self.d_out)
        trajl, datal = self._losses()
        return self.traj.clone(), trajl.copy(), datal.copy(), counts

    def frame_states(self, traj, k):
        """The reference's "shoot" variable of frame k (list-like of nt+1 tuples) as views into `traj`."""
        return LazyStates(self.specs[k], traj[:, k, :], self.Nxs[k] > 0)
