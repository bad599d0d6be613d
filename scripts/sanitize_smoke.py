"""Small invocations of every kernel family added late in round 1 and in round 2 (a quick smoke run; also usable under a memory
checker):
    python scripts/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

os.environ.setdefault("DICP_SMALL_MID", "1")       # forces the mid-size stage kernels wherever they can run

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_icp_b200 import em_ops, ops, shooting          # noqa: E402
from diff_icp_b200.tools.point_sets import decimate, min2_sqdist      # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def rnd(*s):
    return torch.randn(*s, generator=g).to(dev)


def uni(*s):
    return torch.rand(*s, generator=g).to(dev)


# symmetric (q,q) adjoint + general forward, ragged size
for D, eta, wld in ((3, 0.02, True), (2, 0.0, True), (3, 0.0, False)):
    M = 4097
    q, p, a, u = uni(M, D), rnd(M, D), rnd(M, D), rnd(M, D)
    gq, gp = torch.zeros_like(q), torch.zeros_like(q)
    ws = ops.alloc_workspace(M, M, dev)
    ops.rhs_adjoint(D, wld, 0.3, eta, q, p, None, a, u, None, torch.ones(1, device=dev), gq, gp, None, ws)
    vq, dp, scal = torch.zeros_like(q), torch.zeros_like(q), torch.zeros(4, device=dev)
    ops.rhs_forward(D, wld, 0.3, eta, q, p, None, vq, dp, None, scal, ws)
# fused (x,q) adjoint (rectangular ring), ragged sizes
for D, eta, wld in ((3, 0.0, True), (2, 0.03, True)):
    M, Nx = 257, 2049
    q, p, a, u, x, wx = uni(M, D), rnd(M, D), rnd(M, D), rnd(M, D), uni(Nx, D), rnd(Nx, D)
    gq, gp, gx = torch.zeros_like(q), torch.zeros_like(q), torch.zeros_like(x)
    ws = ops.alloc_workspace(Nx, Nx, dev)
    ops.rhs_adjoint(D, wld, 0.3, eta, q, p, x, a, u, wx, torch.ones(1, device=dev), gq, gp, gx, ws)
# batched small-support closure (ring adjoint stage), ragged frames, Euler and Ralston
for D, scheme, eta in ((2, "Euler", 0.0), (3, "Ralston", 0.02)):
    Ms, Nxs = [25, 7, 33], [1000, 333, 1290]
    plan = shooting.BatchedClosurePlan(D, 3, scheme, True, 0.3, eta, 10.0, dev, Ms, Nxs, use_graph=False)
    plan.set_geometry([uni(m, D) for m in Ms], [uni(n, D) for n in Nxs])
    plan.set_targets(uni(sum(Nxs), D), torch.full((sum(Nxs),), 5.0, device=dev))
    plan.active[:] = 1
    plan.evaluate()
    plan.finalize([np.zeros((m, D), np.float32) for m in Ms], coverage_radius=0.5)
# mid-size supports: 128-register forward stage, one-evaluation-per-pair adjoint stage (2 and 4 data points per lane), ragged frames
# (DICP_SMALL_MID=1 above forces the mid form at these small sizes; run once more with DICP_SMALL_MID_R=4 for the 4-point form)
for D, scheme, eta in ((3, "Ralston", 0.0), (2, "Euler", 0.02)):
    Ms, Nxs = [130, 65, 257], [1500, 513, 2050]
    plan = shooting.BatchedClosurePlan(D, 2, scheme, True, 0.3, eta, 10.0, dev, Ms, Nxs, use_graph=False)
    plan.set_geometry([uni(m, D) for m in Ms], [uni(n, D) for n in Nxs])
    plan.set_targets(uni(sum(Nxs), D), torch.full((sum(Nxs),), 5.0, device=dev))
    plan.active[:] = 1
    plan.evaluate()
# EM: packed row passes, few-component column kernel, M step
N, C, D = 5001, 13, 3
X, mu, w = uni(N, D), uni(C, D), torch.zeros(C, device=dev)
wl2 = (w - torch.logsumexp(w, 0)).contiguous()
T2 = em_ops.rowpass(0.2, X, mu, wl2)
st = em_ops.colstats(0.2, X, T2, mu, wl2)
mu2, w2, lpi, ms = em_ops.mstep(st, mu, w, True, True, 1)
em_ops.rowpass(0.2, X, mu, wl2, mu2, lpi.contiguous())
# round 2: fused first sweep, all-reduce buffer + merged M step, EM loop on the device (eager first call, WHILE graph second)
st2 = em_ops.lse_colstats(0.2, X, mu, wl2)
buf = em_ops.reduce_pack(st2, torch.round(st2[:, 0]), torch.zeros(5, device=dev))
em_ops.mstep_merged(buf, torch.round(st2[:, 0]), mu, w, True, True, 1, 5)
from diff_icp_b200.core.GMM import GaussianMixtureUnif        # noqa: E402
G = GaussianMixtureUnif(mu.clone(), sigma=0.2, spec={"device": dev, "dtype": torch.float32})
for _ in range(2):
    G.EM_optimization(X, max_iterations=4, tol=1e-6)
# round 2: one-launch closure on thread-block clusters (both CTA sizes, 8 and 16 CTAs per frame) and the device L-BFGS
from diff_icp_b200.tools.optim import LBFGS_optimization_lockstep      # noqa: E402
for D, Ms, Nxs, shape in ((2, [25, 7, 33], [1000, 333, 1290], "128,8"), (3, [9, 30], [257, 2100], "64,16"),
                          (2, [64, 5], [4000, 17], "128,16")):
    os.environ["DICP_CC_SHAPE"] = shape
    plan = shooting.BatchedClosurePlan(D, 3, "Euler", True, 0.3, 0.0, 10.0, dev, Ms, Nxs, use_graph=False)
    assert plan.one_launch and plan.device_lbfgs
    plan.set_geometry([uni(m, D) for m in Ms], [uni(n, D) for n in Nxs])
    plan.set_targets(uni(sum(Nxs), D), torch.full((sum(Nxs),), 5.0, device=dev))
    plan.active[:] = 1
    plan.evaluate()
    for _ in range(2):                       # eager rounds, then the WHILE graph
        LBFGS_optimization_lockstep([np.zeros((m, D), np.float32) for m in Ms], plan, nmax=1, tol=1e-4)
os.environ.pop("DICP_CC_SHAPE", None)
# point-set helpers
x = uni(700, 2)
decimate(x, 0.08)
min2_sqdist(x)
torch.cuda.synchronize()
print("sanitize_smoke: done")
