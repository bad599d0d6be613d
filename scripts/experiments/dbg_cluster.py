"""Debug: one-launch cluster closure vs stage kernels, component by component."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from diff_icp_b200 import shooting
from diff_icp_b200.core.LDDMM import LDDMMModel
dev = torch.device("cuda:0")
def run(D, version, scheme, nt, Ms, Nxs):
    sig, lam = 0.25, 50.0
    LM = LDDMMModel(sigma=sig, D=D, lambd=lam, version=version, scheme=scheme, nt=nt, spec={"device": dev, "dtype": torch.float32})
    K = len(Ms)
    g = torch.Generator().manual_seed(3)
    q0 = [torch.rand(m, D, generator=g).to(dev) for m in Ms]
    x0 = [torch.rand(n, D, generator=g).to(dev) for n in Nxs]
    y = [torch.rand(n, D, generator=g).to(dev) for n in Nxs]
    inv = [(0.5 + torch.rand(n, generator=g)).to(dev) * 20 for n in Nxs]
    p = [0.02 * torch.randn(m, D, generator=g) for m in Ms]
    res = {}
    for ol in (True, False):
        shooting.BatchedClosurePlan.one_launch_closure = ol
        plan = shooting.BatchedClosurePlan(D, nt, scheme, LM.withlogdet, sig, LM.eta, lam, dev, Ms, Nxs, use_graph=False)
        print("one_launch", plan.one_launch)
        plan.set_geometry(q0, x0); plan.set_targets(torch.cat(y), torch.cat(inv))
        plan.active[:] = 1
        for k in range(K):
            plan.X[k, :Ms[k] * D] = p[k].reshape(-1).numpy()
        plan.evaluate()
        res[ol] = (plan._out_np.copy(), plan.traj.clone().cpu().numpy())
    a, b = res[True], res[False]
    for k in range(K):
        print("frame", k, "M", Ms[k], "Nx", Nxs[k])
        print("  scal one :", a[0][k, :8])
        print("  scal stage:", b[0][k, :8])
        ga, gb = a[0][k, 8:8 + Ms[k] * D], b[0][k, 8:8 + Ms[k] * D]
        print("  grad maxdiff", np.abs(ga - gb).max(), "max", np.abs(gb).max())
        MD = Ms[k] * D
        for t in (1, nt):
            xa = a[1][t, k, 2 * MD:2 * MD + Nxs[k] * D]; xb = b[1][t, k, 2 * MD:2 * MD + Nxs[k] * D]
            d = np.abs(xa - xb).reshape(-1, D).max(1)
            print("  x(t=%d) maxdiff" % t, d.max(), "argmax row", d.argmax(), "rows>1e-5:", np.nonzero(d > 1e-5)[0][:10])
        # data loss in fp64 from each path's own x(1)
        for nm, r in (("one", a), ("stage", b)):
            x1 = torch.tensor(r[1][nt, k, 2 * MD:2 * MD + Nxs[k] * D]).double().view(-1, D)
            dl = (inv[k].cpu().double()[:, None] * (x1 - y[k].cpu().double()) ** 2).sum().item()
            print("  fp64 data loss from", nm, "x1:", dl)
run(2, "hybrid", "Euler", 10, [25, 25, 25, 25, 25], [1000, 1300, 777, 128, 1701])
