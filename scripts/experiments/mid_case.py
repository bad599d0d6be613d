"""Gradient of one batched-closure case under the mid-size adjoint stage vs the per-frame closure (and vs DICP_SMALL_MID=0)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from diff_icp_b200 import shooting
from diff_icp_b200.core.LDDMM import LDDMMModel
dev = torch.device("cuda:0")
D, version, scheme, nt, Ms, Nxs = 2, "logdet", "Ralston", 2, [1025, 200], [45000, 38000]
sig, lam = 0.25, 50.0
LM = LDDMMModel(sigma=sig, D=D, lambd=lam, version=version, scheme=scheme, nt=nt, spec={"device": dev, "dtype": torch.float32})
g = torch.Generator().manual_seed(3)
q0 = [torch.rand(m, D, generator=g).to(dev) for m in Ms]
x0 = [torch.rand(n, D, generator=g).to(dev) for n in Nxs]
y = [torch.rand(n, D, generator=g).to(dev) for n in Nxs]
inv = [(0.5 + torch.rand(n, generator=g)).to(dev) * 20 for n in Nxs]
p = [0.02 * torch.randn(m, D, generator=g) for m in Ms]
plan = shooting.BatchedClosurePlan(D, nt, scheme, LM.withlogdet, sig, LM.eta, lam, dev, Ms, Nxs, use_graph=False)
plan.set_geometry(q0, x0)
plan.set_targets(torch.cat(y), torch.cat(inv))
for k in range(2):
    plan.X[k, :Ms[k] * D] = p[k].reshape(-1).numpy()
plan.active[:] = [1, 1]
plan.evaluate()
out = {}
for k in range(2):
    go = plan.grads[k * plan.ostride:k * plan.ostride + Ms[k] * D].copy()
    sp = LM._spec_for(Ms[k], Nxs[k], dev)
    cp = shooting.ClosurePlan(sp, False, lam)
    cp.set_problem(q0[k], x0[k], y[k], inv[k])
    L, gr = cp.evaluate(p[k].to(dev))
    gr = gr.reshape(-1).cpu().numpy()
    print("MID", os.environ.get("DICP_SMALL_MID"), "frame", k, "loss", plan.losses[k], float(L), "grad rel err", np.abs(go - gr).max() / np.abs(gr).max())
    np.save(f"gpurun_out/mid_case_{os.environ.get('DICP_SMALL_MID','d')}_{k}.npy", go)
