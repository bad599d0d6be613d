"""Where does Reg_opt spend its time outside the lock-step rounds?  Wall-clock timers around the pieces (synchronised)."""
import math, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from groupwise_iteration import spiral_frames
from diff_icp_b200.core.GMM import GaussianMixtureUnif
from diff_icp_b200.core.LDDMM import LDDMMModel
from diff_icp_b200.core.PSR import DiffPSR
from diff_icp_b200 import shooting
from diff_icp_b200.tools import optim

dev = torch.device("cuda:0")
spec = {"device": dev, "dtype": torch.float32}
frames = spiral_frames(int(sys.argv[1]) if len(sys.argv) > 1 else 64, 10000)
torch.manual_seed(1234)
G = GaussianMixtureUnif(torch.zeros(50, 2), spec=spec)
LM = LDDMMModel(sigma=0.2, D=2, lambd=500.0, version="hybrid", scheme="Euler", nt=10, spec=spec)
LM.use_cuda_graph = True
P = DiffPSR([f.to(dev) for f in frames], G, LM, dataspec=spec, compspec=spec)
P.printstuff = False
P.lockstep_groups = 1
P.set_support_scheme("grid", rho=math.sqrt(2))
P.reinitialize_GMM()
T = {}
def timed(obj, name, tag):
    f = getattr(obj, name)
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); T[tag] = T.get(tag, 0.0) + time.perf_counter() - t0
        return r
    setattr(obj, name, w)
timed(optim, "LBFGS_optimization_lockstep", "lbfgs_lockstep")
timed(shooting.BatchedClosurePlan, "finalize", "finalize")
timed(shooting.BatchedClosurePlan, "set_targets", "set_targets")
timed(optim.DeviceLockstepLBFGS, "step", "  opt.step")
timed(DiffPSR, "_register_all", "register_all")
timed(DiffPSR, "update_FE", "update_FE")
import diff_icp_b200.core.PSR as psr
for it in range(5):
    P.GMM_opt(max_iterations=10, tol=1e-3)
    T.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    P.Reg_opt(tol=1e-3, nmax=1)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print(f"iter {it}: Reg_opt {1e3*tot:.2f} ms | " + " | ".join(f"{k} {1e3*v:.2f}" for k, v in T.items()), "| rounds", P._bplan[0].evaluations)
