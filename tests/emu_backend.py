"""Test-only stand-ins for diff_icp_b200.em_ops / ops that run the SAME Op arithmetic on the CPU through tests/hostemu.
Used to exercise the host-side logic (GMM M step, PSR loop, multi-rank reductions) without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import torch

from conftest import ROOT

F = ctypes.POINTER(ctypes.c_float)
_lib = None


def lib():
    global _lib
    if _lib is None:
        if os.environ.get("DICP_HOSTEMU_BUILT") != "1":          # once per test session: spawned ranks inherit the flag
            subprocess.check_call(["sh", os.path.join(ROOT, "tests", "hostemu", "build.sh")])
            os.environ["DICP_HOSTEMU_BUILT"] = "1"
        _lib = ctypes.CDLL(os.path.join(ROOT, "tests", "hostemu", "_build", "libdicp_hostemu.so"))
    return _lib


def _np(t):
    return None if t is None else np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(F)


def rowpass(sigma_old, X, mu_old, wl2, mu_new=None, lpi_new=None, per_point=False):
    N, D = X.shape
    C = mu_old.shape[0]
    lite = mu_new is None
    Xa, ma, wa, mna, lpa = _np(X), _np(mu_old), _np(wl2), _np(mu_new), _np(lpi_new)
    T2 = np.zeros(N, np.float32)
    Y = scal = rowP = rowQ = sq = None
    if not lite:
        Y = np.zeros((N, D), np.float32)
        scal = np.zeros(4, np.float32)
        if per_point:
            rowP, rowQ, sq = (np.zeros(N, np.float32) for _ in range(3))
    if N > 0:
        rc = lib().emu_em_rowpass(D, int(lite), ctypes.c_float(sigma_old), _p(Xa), ctypes.c_int64(N), _p(ma), _p(wa),
                                  ctypes.c_int64(C), _p(mna), _p(lpa), _p(T2), _p(Y), _p(rowP), _p(rowQ), _p(sq), _p(scal))
        assert rc == 0
    t = lambda a: None if a is None else torch.from_numpy(a)
    if lite:
        return t(T2)
    return t(T2), t(Y), t(scal), t(rowP), t(rowQ), t(sq)


def colstats(sigma_old, X, T2, mu_old, wl2):
    N, D = X.shape
    C = mu_old.shape[0]
    stats = np.zeros((C, D + 3), np.float32)
    Xa, Ta, ma, wa = _np(X), _np(T2), _np(mu_old), _np(wl2)
    rc = lib().emu_em_colstats(D, ctypes.c_float(sigma_old), _p(Xa), ctypes.c_int64(N), _p(Ta), _p(ma), _p(wa),
                               ctypes.c_int64(C), _p(stats))
    assert rc == 0
    return torch.from_numpy(stats)


def lse_colstats(sigma_old, X, mu_old, wl2):
    return colstats(sigma_old, X, rowpass(sigma_old, X, mu_old, wl2), mu_old, wl2)


def mstep(stats, mu_old, w_old, do_mu, do_w, sig_mode):
    """CPU stand-in of dicp_em_mstep: the same formulas with torch ops (csrc/em_col_small.cuh, em_mstep_kernel)."""
    D = mu_old.shape[1]
    m, S0, B, A = stats[:, 0], stats[:, 1], stats[:, 2:2 + D], stats[:, 2 + D]
    mu_new = (mu_old + B / S0[:, None]).contiguous() if do_mu else mu_old.clone()
    w_new = (m + torch.log2(S0)) * 0.6931471805599453 if do_w else w_old.clone()
    lse = torch.logsumexp(w_new, 0)
    if sig_mode == 1:
        nd = (torch.exp2(m) * (A - (B * B).sum(-1) / S0)).sum()
    elif sig_mode == 2:
        nd = (torch.exp2(m) * A).sum()
    else:
        nd = torch.zeros(())
    return mu_new, w_new, w_new - lse, torch.stack((nd, lse)).float()


def reduce_pack(stats, m_ref, extra):
    """CPU stand-in of dicp_em_reduce_pack (csrc/em_col_small.cuh, em_reduce_pack_kernel)."""
    d = stats[:, 0] - m_ref
    scaled = stats[:, 1:] * torch.exp2(torch.clamp(d, max=120.0))[:, None]
    flag = (d > 100.0).any().to(stats.dtype).reshape(1)
    parts = [scaled.reshape(-1), flag] + ([extra.to(stats.dtype).reshape(-1)] if extra is not None else [])
    return torch.cat(parts)


def mstep_merged(buf, m_ref, mu_old, w_old, do_mu, do_w, sig_mode, n_extra):
    """CPU stand-in of dicp_em_mstep_merged."""
    C, D = mu_old.shape
    n = C * (D + 2)
    sums = buf[:n].view(C, D + 2)
    merged = torch.cat((m_ref[:, None], sums), dim=1)
    mu_new, w_new, lpi_new, ms = mstep(merged, mu_old, w_old, do_mu, do_w, sig_mode)
    m_next = torch.round(m_ref + torch.log2(torch.clamp(sums[:, 0], min=1e-30)))
    host = torch.cat((ms[:1], buf[n + 1:n + 1 + n_extra], buf[n:n + 1], (sums[:, 0] < 1e-30).any().to(buf.dtype).reshape(1)))
    return mu_new, w_new, lpi_new, m_next, host


def install(monkeypatch):
    from diff_icp_b200 import em_ops
    monkeypatch.setattr(em_ops, "rowpass", rowpass)
    monkeypatch.setattr(em_ops, "colstats", colstats)
    monkeypatch.setattr(em_ops, "lse_colstats", lse_colstats)
    monkeypatch.setattr(em_ops, "mstep", mstep)
    monkeypatch.setattr(em_ops, "reduce_pack", reduce_pack)
    monkeypatch.setattr(em_ops, "mstep_merged", mstep_merged)


# ---- kernel sums / LDDMM right-hand side on the CPU emulation -------------------------------------------------------
_SLOTS = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
_VECTOR = {4, 8, 16, 32, 64, 256}


def ksum(mask, sigma, x, y, b=None, c=None, d=None, ws=None):
    M, D = x.shape
    N = y.shape[0]
    xa, ya, ba, ca, da = _np(x), _np(y), _np(b), _np(c), _np(d)
    outs, args = {}, []
    for sel in _SLOTS:
        if mask & sel:
            outs[sel] = np.zeros((M, D) if sel in _VECTOR else (M,), np.float32)
            args.append(_p(outs[sel]))
        else:
            args.append(None)
    if M > 0:
        rc = lib().emu_ksum(D, ctypes.c_uint(mask), ctypes.c_float(sigma), _p(xa), ctypes.c_int64(M), _p(ya),
                            ctypes.c_int64(N), _p(ba), _p(ca), _p(da), *args)
        assert rc == 0, rc
    return {k: torch.from_numpy(v) for k, v in outs.items()}


def _inplace(fn, outs):
    """Run fn on numpy copies of the output tensors, then copy the results back into the (possibly strided) views."""
    arrs = [None if o is None else np.zeros(tuple(o.shape), np.float32) for o in outs]
    fn(arrs)
    for o, a in zip(outs, arrs):
        if o is not None:
            o.copy_(torch.from_numpy(a))


def rhs_forward(D, withlogdet, sigma, eta, q, p, x, vq, dp, vx, scal, ws, engine=-1):
    M, Nx = q.shape[0], (0 if x is None else x.shape[0])
    qa, pa, xa = _np(q), _np(p), _np(x)

    def run(a):
        rc = lib().emu_rhs_forward(D, int(bool(withlogdet)), ctypes.c_float(sigma), ctypes.c_float(eta), _p(qa), _p(pa),
                                   ctypes.c_int64(M), _p(xa), ctypes.c_int64(Nx), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]))
        assert rc == 0, rc
    _inplace(run, [vq, dp, vx, scal[:4]])


def rhs_adjoint(D, withlogdet, sigma, eta, q, p, x, a, u, wx, gc, gq, gp, gx, ws, engine=-1):
    M, Nx = q.shape[0], (0 if x is None else x.shape[0])
    qa, pa, xa, aa, ua, wa, ga = _np(q), _np(p), _np(x), _np(a), _np(u), _np(wx), _np(gc)

    def run(o):
        rc = lib().emu_rhs_adjoint(D, int(bool(withlogdet)), ctypes.c_float(sigma), ctypes.c_float(eta), _p(qa), _p(pa),
                                   ctypes.c_int64(M), _p(xa), ctypes.c_int64(Nx), _p(aa), _p(ua), _p(wa), _p(ga),
                                   _p(o[0]), _p(o[1]), _p(o[2]))
        if rc != 0:
            raise RuntimeError(f"emu_rhs_adjoint rc={rc}")
    _inplace(run, [gq, gp, gx])


def axpy(out, a, alpha, f1, beta=0.0, f2=None, n=None):
    n = out.numel() if n is None else n
    r = a[:n] + np.float32(alpha) * f1[:n]
    if f2 is not None:
        r = r + np.float32(beta) * f2[:n]
    out[:n] = r


def install_all(monkeypatch):
    """Route every C-ABI wrapper of the product through the CPU emulation (tests of host logic only)."""
    from diff_icp_b200 import _lib, em_ops, ops
    install(monkeypatch)
    monkeypatch.setattr(ops, "ksum", ksum)
    monkeypatch.setattr(ops, "rhs_forward", rhs_forward)
    monkeypatch.setattr(ops, "rhs_adjoint", rhs_adjoint)
    monkeypatch.setattr(ops, "axpy", axpy)
    monkeypatch.setattr(ops, "quad_loss", quad_loss)
    monkeypatch.setattr(ops, "small_enabled", False)        # the CPU emulation covers the general engine path
    monkeypatch.setattr(ops, "alloc_workspace", lambda r, c, dev: torch.empty(16, dtype=torch.uint8))
    monkeypatch.setattr(_lib, "require_cuda", lambda *t: torch.device("cpu"))
    monkeypatch.setattr(em_ops, "log_resp", None, raising=False)
    # greedy decimation: the oracle's dense restatement stands in for the CUDA kernel (index lists are bit-exact, see
    # tests/test_gpu_pointsets.py)
    from oracle import pointsets as ops_oracle
    from diff_icp_b200.core import PSR as psr_mod
    monkeypatch.setattr(psr_mod, "decimate", ops_oracle.decimate)


def quad_loss(x, y, inv, g, loss, ws):
    r = x - y
    g.copy_(2.0 * inv[:, None] * r)
    loss[0] = float((inv[:, None].double() * r.double() ** 2).sum())


def _install_quad(monkeypatch):
    from diff_icp_b200 import ops
    monkeypatch.setattr(ops, "quad_loss", quad_loss)
