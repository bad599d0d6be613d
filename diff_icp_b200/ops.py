"""Thin functional wrappers over the C ABI (include/dicp_b200.h): tensors in, tensors out, current stream.

These are the only functions that touch the shared library for the kernel-sum / LDDMM part of the path.
Every call validates device / dtype / contiguity like the reference's getspec (tools/spec.py:39-43) and
raises ValueError on mismatch; errors from the library are raised as DicpError.  No CPU path exists.
"""

from __future__ import annotations

import torch

from . import _lib
from ._lib import check, load, ptr, require_cuda, stream_ptr, workspace

# output selectors (include/dicp_b200.h)
K_BASE, K_REDSCAL, K_RED, K_GRAD, K_DD, K_GEND, K_HESS, K_LAP, K_GRADLAP, K_MINSQ, K_DOT = (
    1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)
_SLOTS = [K_BASE, K_REDSCAL, K_RED, K_GRAD, K_DD, K_GEND, K_HESS, K_LAP, K_GRADLAP, K_MINSQ, K_DOT]
_VECTOR = {K_RED, K_GRAD, K_DD, K_GEND, K_HESS, K_GRADLAP}


def _c(t):
    return None if t is None else t.detach().contiguous()


def ksum(mask: int, sigma: float, x, y, b=None, c=None, d=None, ws=None):
    """Gaussian kernel reductions selected by `mask` in one sweep. Returns {selector: tensor}."""
    x, y, b, c, d = _c(x), _c(y), _c(b), _c(c), _c(d)
    dev = require_cuda(x, y, b, c, d)
    M, D = x.shape
    N = y.shape[0]
    if y.shape[1] != D or D not in (2, 3):
        raise ValueError("point sets must be (n,2) or (n,3) with matching dimension")
    outs, args = {}, []
    for sel in _SLOTS:
        if mask & sel:
            outs[sel] = torch.empty((M, D) if sel in _VECTOR else (M,), dtype=torch.float32, device=dev)
            args.append(ptr(outs[sel]))
        else:
            args.append(None)
    if M == 0:
        return outs
    if N == 0:
        for sel, o in outs.items():
            o.fill_(float("inf") if sel == K_MINSQ else 0.0)
        return outs
    if ws is None:
        ws = workspace(M, N, dev)
    with torch.cuda.device(dev):
        rc = load().dicp_ksum(D, mask, float(sigma), ptr(x), M, ptr(y), N, ptr(b), ptr(c), ptr(d), *args,
                              ptr(ws), ws.numel(), stream_ptr())
    check(rc, "dicp_ksum")
    return outs


def rhs_forward(D, withlogdet, sigma, eta, q, p, x, vq, dp, vx, scal, ws):
    """Fused ODE right-hand side; all arguments are preallocated contiguous fp32 CUDA tensors (x, vx may be None)."""
    M = q.shape[0]
    Nx = 0 if x is None else x.shape[0]
    rc = load().dicp_rhs_forward(D, int(bool(withlogdet)), float(sigma), float(eta), ptr(q), ptr(p), M,
                                 ptr(x), Nx, ptr(vq), ptr(dp), ptr(vx), ptr(scal), ptr(ws), ws.numel(), stream_ptr())
    check(rc, "dicp_rhs_forward")


def rhs_adjoint(D, withlogdet, sigma, eta, q, p, x, a, u, wx, gc, gq, gp, gx, ws):
    M = q.shape[0]
    Nx = 0 if x is None else x.shape[0]
    rc = load().dicp_rhs_adjoint(D, int(bool(withlogdet)), float(sigma), float(eta), ptr(q), ptr(p), M,
                                 ptr(x), Nx, ptr(a), ptr(u), ptr(wx), ptr(gc), ptr(gq), ptr(gp), ptr(gx),
                                 ptr(ws), ws.numel(), stream_ptr())
    check(rc, "dicp_rhs_adjoint")


def axpy(out, a, alpha, f1, beta=0.0, f2=None, n=None):
    """out[:n] = a[:n] + alpha*f1[:n] + beta*f2[:n] on flat contiguous fp32 CUDA tensors."""
    if n is None:
        n = out.numel()
    rc = load().dicp_axpy(n, ptr(out), ptr(a), float(alpha), ptr(f1), float(beta), ptr(f2), stream_ptr())
    check(rc, "dicp_axpy")


# ---- fused integrator stages for small supports ---------------------------------------------------------------------
small_enabled = True


def use_small_path(M):
    """True when the support set is small enough for the one-launch-per-stage kernels (csrc/small_step.cuh)."""
    return small_enabled and M <= load().dicp_small_max_support()


def alloc_small_workspace(M, Nx, device):
    """Zero-initialised (ticket counters) workspace for the small-support stage kernels."""
    return torch.zeros(int(load().dicp_small_workspace_bytes(int(M), int(Nx))), dtype=torch.uint8, device=device)


def small_rhs_step(D, withlogdet, sigma, eta, M, Nx, s_eval, base, other, c_this, c_other, out, F, ws):
    rc = load().dicp_small_rhs_step(D, int(bool(withlogdet)), float(sigma), float(eta), M, Nx, ptr(s_eval), ptr(base),
                                    ptr(other), float(c_this), float(c_other), ptr(out), ptr(F), ptr(ws), ws.numel(),
                                    stream_ptr())
    check(rc, "dicp_small_rhs_step")


def small_adj_step(D, withlogdet, sigma, eta, M, Nx, s_eval, lam, base, other, add, c_this, c_other, out, G, ws):
    rc = load().dicp_small_adj_step(D, int(bool(withlogdet)), float(sigma), float(eta), M, Nx, ptr(s_eval), ptr(lam),
                                    ptr(base), ptr(other), ptr(add), float(c_this), float(c_other), ptr(out), ptr(G),
                                    ptr(ws), ws.numel(), stream_ptr())
    check(rc, "dicp_small_adj_step")


def quad_loss(x, y, inv, g, loss, ws):
    """loss[0] = sum_n inv[n]|x_n-y_n|^2, g = 2 inv (x-y); all preallocated contiguous fp32 CUDA tensors."""
    n, D = x.shape
    rc = load().dicp_quad_loss(D, ptr(x), ptr(y), ptr(inv), n, ptr(g), ptr(loss), ptr(ws), ws.numel(), stream_ptr())
    check(rc, "dicp_quad_loss")


def alloc_workspace(rows, cols, device):
    """A private workspace tensor for a long-lived plan (e.g. a CUDA-graph-captured shoot)."""
    return torch.empty(int(load().dicp_pair_workspace_bytes(int(rows), int(cols))), dtype=torch.uint8, device=device)


def pipe_probe(which: int, blocks: int, iters: int, out):
    rc = load().dicp_pipe_probe(which, blocks, iters, ptr(out), stream_ptr())
    check(rc, "dicp_pipe_probe")
