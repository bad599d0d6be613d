"""A few end-to-end runs off the beaten path (ragged frames, 3-D grid supports, one frame, logdet with a decimated support,
outliers): each must run, keep the free energy from increasing and stay finite.   python scripts/robustness_runs.py"""
import math
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_icp_b200.core.GMM import GaussianMixtureUnif          # noqa: E402
from diff_icp_b200.core.LDDMM import LDDMMModel                 # noqa: E402
from diff_icp_b200.core.PSR import DiffPSR                      # noqa: E402

dev = torch.device("cuda:0")
spec = {"device": dev, "dtype": torch.float32}
warnings.simplefilter("ignore")


def cloud(n, D, g, shift=0.0):
    c = torch.rand(12, D, generator=g)
    x = c[torch.randint(0, 12, (n,), generator=g)] + 0.04 * torch.randn(n, D, generator=g) + shift
    return x.to(dev)


def run(tag, sizes, D, version, scheme, support, C=10, S=1, outliers=False, iters=2, rho=1.0):
    g = torch.Generator().manual_seed(len(tag))
    x = [[cloud(n + 13 * s, D, g, 0.02 * k) for s in range(S)] if S > 1 else cloud(n, D, g, 0.02 * k)
         for k, n in enumerate(sizes)]
    torch.manual_seed(0)
    G = GaussianMixtureUnif(torch.zeros(C, D), spec=spec, use_outliers=outliers)
    LM = LDDMMModel(sigma=0.25, D=D, lambd=200.0, version=version, scheme=scheme, nt=6, spec=spec)
    LM.use_cuda_graph = True
    P = DiffPSR(x, G, LM, dataspec=spec, compspec=spec)
    P.printstuff = False
    P.set_support_scheme(support, rho=rho)
    P.reinitialize_GMM()
    fes = [P.FE]
    for _ in range(iters):
        P.GMM_opt(max_iterations=5, tol=1e-3)
        fes.append(P.FE)
        P.Reg_opt(nmax=1, tol=1e-3)
        fes.append(P.FE)
    lock = P._batched_plan() is not None
    ok = all(math.isfinite(f) for f in fes) and all(b <= a + 1e-5 * abs(a) for a, b in zip(fes, fes[1:]))
    finite = all(bool(torch.isfinite(a).all()) for a in P.a0)
    print(f"{tag:34s} lockstep={lock!s:5s} support={int(P.q0[0].shape[0]):4d} FE {fes[0]:.6g} -> {fes[-1]:.6g}  "
          f"{'ok' if ok and finite else 'FAILED'}")
    return ok and finite


res = [
    run("ragged 2-D decim hybrid", [500, 20000, 3000, 129, 7777], 2, "hybrid", "Euler", "decim", rho=1.0),
    run("3-D grid hybrid Ralston", [4000, 5000, 6000], 3, "hybrid", "Ralston", "grid", rho=1.5),
    run("one frame 2-D grid classic", [8000], 2, "classic", "Euler", "grid", rho=math.sqrt(2)),
    run("logdet 2-D decim (ring, eta)", [3000, 3500], 2, "logdet", "Euler", "decim", rho=1.0),
    run("two structures + outliers", [2500, 2600, 2700], 2, "hybrid", "Euler", "grid", S=2, outliers=True, rho=math.sqrt(2)),
    run("3-D decim logdet Ralston", [6000, 6500], 3, "logdet", "Ralston", "decim", rho=1.2),
]
print("ALL OK" if all(res) else "SOME FAILED")
sys.exit(0 if all(res) else 1)
