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
        subprocess.check_call(["sh", os.path.join(ROOT, "tests", "hostemu", "build.sh")])
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


def install(monkeypatch):
    from diff_icp_b200 import em_ops
    monkeypatch.setattr(em_ops, "rowpass", rowpass)
    monkeypatch.setattr(em_ops, "colstats", colstats)
