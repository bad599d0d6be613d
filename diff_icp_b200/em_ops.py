"""Functional wrappers over the EM entry points of the C ABI (dicp_em_rowpass / dicp_em_colstats / dicp_em_lse_colstats /
dicp_em_mstep / dicp_log_resp).

Inputs and outputs are contiguous fp32 CUDA tensors; everything is enqueued on the current stream without host
synchronisation.  No CPU path.
"""

from __future__ import annotations

import torch

from ._lib import check, load, ptr, require_cuda, stream_ptr, workspace


def rowpass(sigma_old, X, mu_old, wl2, mu_new=None, lpi_new=None, per_point=False):
    """Row pass of the EM step.  lite (mu_new is None): returns T2 (N,), log2-domain row LSE.
    full: returns (T2, Y, scal4, rowP, rowQ, sq) with the per-point arrays only when per_point=True."""
    dev = require_cuda(X, mu_old, wl2, mu_new, lpi_new)
    N, D = X.shape
    C = mu_old.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    T2 = torch.empty(N, **f32)
    lite = mu_new is None
    Y = scal = rowP = rowQ = sq = None
    if not lite:
        Y = torch.empty(N, D, **f32)
        scal = torch.zeros(4, **f32)
        if per_point:
            rowP, rowQ, sq = (torch.empty(N, **f32) for _ in range(3))
    if N > 0:
        ws = workspace(max(N, C), max(N, C), dev)
        rc = load().dicp_em_rowpass(D, int(lite), float(sigma_old), ptr(X), N, ptr(mu_old), ptr(wl2), C,
                                    ptr(mu_new), ptr(lpi_new), ptr(T2), ptr(Y), ptr(rowP), ptr(rowQ), ptr(sq),
                                    ptr(scal), ptr(ws), ws.numel(), stream_ptr())
        check(rc, "dicp_em_rowpass")
    if lite:
        return T2
    return T2, Y, scal, rowP, rowQ, sq


def colstats(sigma_old, X, T2, mu_old, wl2):
    """Column statistics (C, D+3) = [m (log2), S0, B (D), A] of every component over the points X."""
    dev = require_cuda(X, T2, mu_old, wl2)
    N, D = X.shape
    C = mu_old.shape[0]
    stats = torch.empty(C, D + 3, dtype=torch.float32, device=dev)
    ws = workspace(max(N, C), max(N, C), dev)
    rc = load().dicp_em_colstats(D, float(sigma_old), ptr(X), N, ptr(T2), ptr(mu_old), ptr(wl2), C, ptr(stats),
                                 ptr(ws), ws.numel(), stream_ptr())
    check(rc, "dicp_em_colstats")
    return stats


SMALL_C = 64        # up to this many components the first sweep is one launch / one read of X (csrc/em_col_small.cuh)


def lse_colstats(sigma_old, X, mu_old, wl2):
    """First sweep of the EM step: column statistics (C, D+3) = [m (log2), S0, B (D), A] of every component, with the row
    log-sum-exp computed on the way (rowpass(lite) + colstats as ONE call; one launch and one read of X when C <= SMALL_C)."""
    dev = require_cuda(X, mu_old, wl2)
    N, D = X.shape
    C = mu_old.shape[0]
    stats = torch.empty(C, D + 3, dtype=torch.float32, device=dev)
    T2 = torch.empty(N, dtype=torch.float32, device=dev) if C > SMALL_C else None
    ws = workspace(max(N, C), max(N, C), dev)
    rc = load().dicp_em_lse_colstats(D, float(sigma_old), ptr(X), N, ptr(mu_old), ptr(wl2), C, ptr(T2), ptr(stats),
                                     ptr(ws), ws.numel(), stream_ptr())
    check(rc, "dicp_em_lse_colstats")
    return stats


def mstep(stats, mu_old, w_old, do_mu, do_w, sig_mode):
    """M step on the (merged) column statistics in one launch.  Returns (mu_new (C,D), w_new (C,), lpi_new (C,),
    scal (2,) = [N D sigma'^2 or 0, LSE(w_new)]); sig_mode: 0 none, 1 distances to the new centroids, 2 to the old ones."""
    dev = require_cuda(stats, mu_old, w_old)
    C, D = mu_old.shape
    f32 = dict(dtype=torch.float32, device=dev)
    mu_new, w_new, lpi_new, scal = torch.empty(C, D, **f32), torch.empty(C, **f32), torch.empty(C, **f32), torch.empty(2, **f32)
    rc = load().dicp_em_mstep(D, ptr(stats), ptr(mu_old), ptr(w_old), C, int(bool(do_mu)), int(bool(do_w)), int(sig_mode),
                              ptr(mu_new), ptr(w_new), ptr(lpi_new), ptr(scal), stream_ptr())
    check(rc, "dicp_em_mstep")
    return mu_new, w_new, lpi_new, scal


def reduce_pack(stats, m_ref, extra):
    """The buffer of the EM step's one all-reduce: [S0, B, A rescaled to the exponents m_ref | overflow flag | extra]."""
    dev = require_cuda(stats, m_ref, extra)
    C, D = stats.shape[0], stats.shape[1] - 3
    n_extra = 0 if extra is None else extra.numel()
    buf = torch.empty(C * (D + 2) + 1 + n_extra, dtype=torch.float32, device=dev)
    rc = load().dicp_em_reduce_pack(D, ptr(stats), ptr(m_ref), C, ptr(extra), n_extra, ptr(buf), stream_ptr())
    check(rc, "dicp_em_reduce_pack")
    return buf


def mstep_merged(buf, m_ref, mu_old, w_old, do_mu, do_w, sig_mode, n_extra):
    """M step on the all-reduced buffer of `reduce_pack`.  Returns (mu_new, w_new, lpi_new, m_next (C,), host (3 + n_extra,) =
    [N D sigma'^2, reduced extras..., flag sum, vanished-mass flag])."""
    dev = require_cuda(buf, m_ref, mu_old, w_old)
    C, D = mu_old.shape
    f32 = dict(dtype=torch.float32, device=dev)
    mu_new, w_new, lpi_new, m_next = torch.empty(C, D, **f32), torch.empty(C, **f32), torch.empty(C, **f32), torch.empty(C, **f32)
    host = torch.empty(3 + n_extra, **f32)
    rc = load().dicp_em_mstep_merged(D, ptr(buf), ptr(m_ref), ptr(mu_old), ptr(w_old), C, int(bool(do_mu)), int(bool(do_w)),
                                     int(sig_mode), int(n_extra), ptr(mu_new), ptr(w_new), ptr(lpi_new), ptr(m_next), ptr(host),
                                     stream_ptr())
    check(rc, "dicp_em_mstep_merged")
    return mu_new, w_new, lpi_new, m_next, host


def log_resp(sigma, X, mu, w, want_lgam=True, want_argmax=False):
    dev = require_cuda(X, mu, w)
    N, D = X.shape
    C = mu.shape[0]
    lgam = torch.empty(N, C, dtype=torch.float32, device=dev) if want_lgam else None
    amax = torch.empty(N, dtype=torch.int64, device=dev) if want_argmax else None
    rc = load().dicp_log_resp(D, float(sigma), ptr(X), N, ptr(mu), ptr(w), C, ptr(lgam), ptr(amax), stream_ptr())
    check(rc, "dicp_log_resp")
    return lgam, amax
