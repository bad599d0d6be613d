"""CPU: the arithmetic of the CUDA Ops (diff_icp_b200/csrc/ops_*.cuh) and the dispatch logic of the library,
executed on the CPU through tests/hostemu, must agree with the oracle (fp64) on kernel sums, the fused
Hamiltonian right-hand side and its hand-derived adjoint.  This pins the formulas, packing offsets and constants
in the build container (no GPU); the same comparisons run against the real kernels under -m gpu."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, relerr
from oracle.kernels import GaussOracle
from oracle.lddmm import LDDMMOracle

F = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def emu():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    subprocess.check_call(["sh", os.path.join(ROOT, "tests", "hostemu", "build.sh")])
    return ctypes.CDLL(os.path.join(ROOT, "tests", "hostemu", "_build", "libdicp_hostemu.so"))


def fp(a):
    return None if a is None else a.ctypes.data_as(F)


def f32(t):
    return np.ascontiguousarray(np.asarray(t, dtype=np.float32))


MASKS = {"KBase": 1, "KRedScal": 2, "KRed": 4, "GradKRed": 8, "DDKRed": 16, "GenDKRed": 32, "HessKRed": 64,
         "LapKRed": 128, "GradLapKRed": 256, "MinSq": 512}
SLOT = {"KBase": 0, "KRedScal": 1, "KRed": 2, "GradKRed": 3, "DDKRed": 4, "GenDKRed": 5, "HessKRed": 6,
        "LapKRed": 7, "GradLapKRed": 8, "MinSq": 9, "Dot": 10}


def emu_ksum(lib, D, mask, sigma, x, y, b=None, c=None, d=None):
    M, N = x.shape[0], y.shape[0]
    outs = {}
    args = []
    for name, slot in sorted(SLOT.items(), key=lambda kv: kv[1]):
        if mask >> slot & 1:
            vec = name in ("KRed", "GradKRed", "DDKRed", "GenDKRed", "HessKRed", "GradLapKRed")
            outs[name] = np.zeros((M, D) if vec else (M,), dtype=np.float32)
            args.append(fp(outs[name]))
        else:
            args.append(None)
    rc = lib.emu_ksum(D, ctypes.c_uint(mask), ctypes.c_float(sigma), fp(x), ctypes.c_int64(M), fp(y), ctypes.c_int64(N),
                      fp(b), fp(c), fp(d), *args)
    assert rc == 0
    return outs


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_emulated_kernel_sums_match_gold(emu, golden, tag):
    g = golden("kernels")
    M, N, D, sig = g[f"{tag}_meta"]
    D = int(D)
    x, y, b, c, d = (f32(g[f"{tag}_in_{k}"]) for k in "xybcd")
    for name, mask in MASKS.items():
        if name == "MinSq":
            continue
        out = emu_ksum(emu, D, mask, float(sig), x, y, b, c, d)[name]
        gold = g[f"{tag}_gold_{name}"]
        assert relerr(out, gold) < 2e-5, (name, relerr(out, gold))
    # reversed gradient sum: rows = y, cols = x with vectors c
    out = emu_ksum(emu, D, 1024, float(sig), y, x, b=c)["Dot"]
    assert relerr(out, g[f"{tag}_gold_GradKRed_rev"]) < 2e-5
    # fused subsets agree with the singles
    fused = emu_ksum(emu, D, 4 | 8 | 128, float(sig), x, y, b, c, d)
    for name in ("KRed", "GradKRed", "LapKRed"):
        assert relerr(fused[name], g[f"{tag}_gold_{name}"]) < 2e-5
    # coverage: min squared distance
    ms = emu_ksum(emu, D, 512, float(sig), x, y)["MinSq"]
    ref = GaussOracle(sig, D).min_sqdist(torch.from_numpy(x).double(), torch.from_numpy(y).double())
    assert relerr(ms, ref.numpy()) < 1e-5


def _rand_case(D, Nq, Nx, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.rand(Nq, D, generator=g)
    p = 0.3 * torch.randn(Nq, D, generator=g)
    x = torch.rand(Nx, D, generator=g) if Nx else None
    return q, p, x


@pytest.mark.parametrize("D", [2, 3])
@pytest.mark.parametrize("version", ["classic", "hybrid", "logdet"])
@pytest.mark.parametrize("with_x", [False, True])
def test_emulated_rhs_forward_matches_oracle(emu, D, version, with_x):
    sig, lam = 0.3, 5.0
    Nq, Nx = 150, (260 if with_x else 0)       # > one 128-column tile
    q, p, x = _rand_case(D, Nq, Nx, 7 + D)
    LM = LDDMMOracle(sigma=sig, D=D, lambd=lam, version=version)
    ode = LM.ode(q.double(), p.double(), torch.zeros(1, dtype=torch.float64), None if x is None else x.double())
    qa, pa = f32(q), f32(p)
    xa = f32(x) if with_x else None
    vq, dp = np.zeros_like(qa), np.zeros_like(pa)
    vx = np.zeros_like(xa) if with_x else None
    scal = np.zeros(4, dtype=np.float32)
    # plain evaluation (every ordered pair) and symmetric evaluation (every unordered (q,q) pair once, Op::pair_sym)
    for fn in (emu.emu_rhs_forward, emu.emu_rhs_forward_sym):
        vq[:], dp[:], scal[:] = 0, 0, 0
        rc = fn(D, int(LM.withlogdet), ctypes.c_float(sig), ctypes.c_float(LM.eta), fp(qa), fp(pa),
                ctypes.c_int64(Nq), fp(xa), ctypes.c_int64(Nx), fp(vq), fp(dp), fp(vx), fp(scal))
        assert rc == 0
        assert relerr(vq, ode[0].numpy()) < 2e-5
        assert relerr(dp, ode[1].numpy()) < 2e-5
        dc = float(ode[2].sum())
        assert abs(scal[0] - dc) < 2e-5 * max(1.0, abs(dc)), (scal[0], dc)
        if with_x:
            assert relerr(vx, ode[3].numpy()) < 2e-5
        H = float(LM.hamiltonian(q.double(), p.double()))
        Hk = 0.5 * scal[1] - LM.eta * scal[2] - 0.5 * LM.eta ** 2 * scal[3]
        assert abs(Hk - H) < 2e-5 * max(1.0, abs(H)), (Hk, H)


@pytest.mark.parametrize("D", [2, 3])
@pytest.mark.parametrize("version", ["classic", "hybrid", "logdet"])
@pytest.mark.parametrize("with_x", [False, True])
def test_emulated_rhs_adjoint_matches_autograd(emu, D, version, with_x):
    sig, lam = 0.3, 5.0
    Nq, Nx = 140, (200 if with_x else 0)
    q, p, x = _rand_case(D, Nq, Nx, 31 + D)
    g = torch.Generator().manual_seed(5)
    a = torch.randn(Nq, D, generator=g)
    u = torch.randn(Nq, D, generator=g)
    wx = torch.randn(Nx, D, generator=g) if with_x else None
    gc = 0.7
    LM = LDDMMOracle(sigma=sig, D=D, lambd=lam, version=version)
    qd = q.double().requires_grad_(True)
    pd = p.double().requires_grad_(True)
    xd = x.double().requires_grad_(True) if with_x else None
    ode = LM.ode(qd, pd, torch.zeros(1, dtype=torch.float64), xd)
    L = (a.double() * ode[0]).sum() + (u.double() * ode[1]).sum() + gc * ode[2].sum()
    if with_x:
        L = L + (wx.double() * ode[3]).sum()
    grads = torch.autograd.grad(L, [qd, pd] + ([xd] if with_x else []))
    qa, pa, aa, ua = f32(q), f32(p), f32(a), f32(u)
    xa = f32(x) if with_x else None
    wa = f32(wx) if with_x else None
    gq, gp = np.zeros_like(qa), np.zeros_like(pa)
    gx = np.zeros_like(xa) if with_x else None
    gca = np.array([gc], dtype=np.float32)
    # plain evaluation (every ordered pair) and symmetric evaluation (every unordered (q,q) pair once, Op::pair_sym)
    for fn in (emu.emu_rhs_adjoint, emu.emu_rhs_adjoint_sym):
        gq[:], gp[:] = 0, 0
        rc = fn(D, int(LM.withlogdet), ctypes.c_float(sig), ctypes.c_float(LM.eta), fp(qa), fp(pa),
                ctypes.c_int64(Nq), fp(xa), ctypes.c_int64(Nx), fp(aa), fp(ua), fp(wa), fp(gca),
                fp(gq), fp(gp), fp(gx))
        assert rc == 0
        assert relerr(gq, grads[0].numpy()) < 3e-5, relerr(gq, grads[0].numpy())
        assert relerr(gp, grads[1].numpy()) < 3e-5, relerr(gp, grads[1].numpy())
        if with_x:
            assert relerr(gx, grads[2].numpy()) < 3e-5, relerr(gx, grads[2].numpy())
