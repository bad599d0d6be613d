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


ENGINE_DEFAULT, ENGINE_GENERAL, ENGINE_SYMMETRIC, ENGINE_SYMMETRIC_ALL = -1, 0, 1, 2      # include/dicp_b200.h


def rhs_forward(D, withlogdet, sigma, eta, q, p, x, vq, dp, vx, scal, ws, engine=ENGINE_DEFAULT):
    """Fused ODE right-hand side; all arguments are preallocated contiguous fp32 CUDA tensors (x, vx may be None)."""
    M = q.shape[0]
    Nx = 0 if x is None else x.shape[0]
    rc = load().dicp_rhs_forward(D, int(bool(withlogdet)), float(sigma), float(eta), ptr(q), ptr(p), M,
                                 ptr(x), Nx, ptr(vq), ptr(dp), ptr(vx), ptr(scal), ptr(ws), ws.numel(), stream_ptr(),
                                 int(engine))
    check(rc, "dicp_rhs_forward")


def rhs_adjoint(D, withlogdet, sigma, eta, q, p, x, a, u, wx, gc, gq, gp, gx, ws, engine=ENGINE_DEFAULT):
    M = q.shape[0]
    Nx = 0 if x is None else x.shape[0]
    rc = load().dicp_rhs_adjoint(D, int(bool(withlogdet)), float(sigma), float(eta), ptr(q), ptr(p), M,
                                 ptr(x), Nx, ptr(a), ptr(u), ptr(wx), ptr(gc), ptr(gq), ptr(gp), ptr(gx),
                                 ptr(ws), ws.numel(), stream_ptr(), int(engine))
    check(rc, "dicp_rhs_adjoint")


def axpy(out, a, alpha, f1, beta=0.0, f2=None, n=None):
    """out[:n] = a[:n] + alpha*f1[:n] + beta*f2[:n] on flat contiguous fp32 CUDA tensors."""
    if n is None:
        n = out.numel()
    rc = load().dicp_axpy(n, ptr(out), ptr(a), float(alpha), ptr(f1), float(beta), ptr(f2), stream_ptr())
    check(rc, "dicp_axpy")


# ---- fused integrator stages for small supports ---------------------------------------------------------------------
small_enabled = True


SMALL_SINGLE_MAX = 512        # one frame at a time: above this the general tiled engine is at least as fast


def use_small_path(M, batched=False):
    """True when the support set is small enough for the one-launch-per-stage kernels (csrc/small_step.cuh): up to
    dicp_small_max_support() points when many frames are evaluated together, SMALL_SINGLE_MAX for a single frame."""
    cap = load().dicp_small_max_support()
    return small_enabled and M <= (cap if batched else min(cap, SMALL_SINGLE_MAX))


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


# ---- batched (multi-frame) closure for small supports: blockIdx.y = frame (csrc/batch_closure.cuh) -------------------
def batch_frame_ws_bytes(maxM, maxNx):
    b = int(load().dicp_small_workspace_bytes(int(maxM), int(maxNx)))
    return (b + 255) // 256 * 256


def batch_rhs_step(D, withlogdet, sigma, eta, K, dims, active, maxM, maxNx, fstride, s_eval, base, other, c_this, c_other,
                   out, F, ws, ws_frame):
    rc = load().dicp_batch_rhs_step(D, int(bool(withlogdet)), float(sigma), float(eta), K, ptr(dims), ptr(active), maxM,
                                    maxNx, fstride, ptr(s_eval), ptr(base), ptr(other), float(c_this), float(c_other),
                                    ptr(out), ptr(F), ptr(ws), ws_frame, stream_ptr())
    check(rc, "dicp_batch_rhs_step")


def batch_adj_step(D, withlogdet, sigma, eta, K, dims, active, maxM, maxNx, fstride, s_eval, lam, base, other, add,
                   c_this, c_other, out, G, ws, ws_frame):
    rc = load().dicp_batch_adj_step(D, int(bool(withlogdet)), float(sigma), float(eta), K, ptr(dims), ptr(active), maxM,
                                    maxNx, fstride, ptr(s_eval), ptr(lam), ptr(base), ptr(other), ptr(add), float(c_this),
                                    float(c_other), ptr(out), ptr(G), ptr(ws), ws_frame, stream_ptr())
    check(rc, "dicp_batch_adj_step")


def batch_set_p(D, K, dims, active, maxM, fstride, X, xstride, state0):
    rc = load().dicp_batch_set_p(D, K, ptr(dims), ptr(active), maxM, fstride, ptr(X), xstride, ptr(state0), stream_ptr())
    check(rc, "dicp_batch_set_p")


def batch_quad_loss(D, K, dims, active, max_points, fstride, state_end, y, inv, ystride, g_end, loss, lstride, ws):
    rc = load().dicp_batch_quad_loss(D, K, ptr(dims), ptr(active), max_points, fstride, ptr(state_end), ptr(y), ptr(inv),
                                     ystride, ptr(g_end), ptr(loss), lstride, ptr(ws), ws.numel(), stream_ptr())
    check(rc, "dicp_batch_quad_loss")


def batch_closure_out(D, K, dims, active, maxM, fstride, lam_reg, lam, F0, state_end, out, ostride, nscal):
    rc = load().dicp_batch_closure_out(D, K, ptr(dims), ptr(active), maxM, fstride, float(lam_reg), ptr(lam), ptr(F0),
                                       ptr(state_end), ptr(out), ostride, nscal, stream_ptr())
    check(rc, "dicp_batch_closure_out")


def batch_closure_cluster_rows(D, eta, scheme, maxM, maxNx, nt, K):
    """Rows per CTA of the one-launch closure (csrc/cluster_closure.cuh) if it applies to these sizes, else 0."""
    return int(load().dicp_batch_closure_cluster_rows(int(D), float(eta), int(scheme == "Euler"), int(maxM), int(maxNx), int(nt),
                                                      int(K)))


def batch_closure_cluster(D, withlogdet, sigma, eta, K, dims, active, maxM, maxNx, fstride, nt, traj, tstride, X, xstride,
                          y, inv, ystride, lam_reg, out, ostride, nscal):
    rc = load().dicp_batch_closure_cluster(D, int(bool(withlogdet)), float(sigma), float(eta), K, ptr(dims), ptr(active), maxM,
                                           maxNx, fstride, nt, ptr(traj), tstride, ptr(X), xstride, ptr(y), ptr(inv), ystride,
                                           float(lam_reg), ptr(out), ostride, nscal, stream_ptr())
    check(rc, "dicp_batch_closure_cluster")


def batch_coverage(D, K, dims, active, maxM, maxNx, fstride, traj, tstride, ntimes, radius, counts):
    rc = load().dicp_batch_coverage(D, K, ptr(dims), ptr(active), maxM, maxNx, fstride, ptr(traj), tstride, ntimes,
                                    float(radius), ptr(counts), stream_ptr())
    check(rc, "dicp_batch_coverage")
