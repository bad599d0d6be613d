"""cProfile (tottime) of one steady-state GMM_opt + Reg_opt of the 64 x 10k atlas: where is the host time outside the rounds?"""
import cProfile, io, math, os, pstats, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from groupwise_iteration import spiral_frames
from diff_icp_b200.core.GMM import GaussianMixtureUnif
from diff_icp_b200.core.LDDMM import LDDMMModel
from diff_icp_b200.core.PSR import DiffPSR
dev = torch.device("cuda:0")
spec = {"device": dev, "dtype": torch.float32}
frames = spiral_frames(64, 10000)
torch.manual_seed(1234)
G = GaussianMixtureUnif(torch.zeros(50, 2), spec=spec)
LM = LDDMMModel(sigma=0.2, D=2, lambd=500.0, version="hybrid", scheme="Euler", nt=10, spec=spec)
LM.use_cuda_graph = True
P = DiffPSR([f.to(dev) for f in frames], G, LM, dataspec=spec, compspec=spec)
P.printstuff = False
P.set_support_scheme("grid", rho=math.sqrt(2))
P.reinitialize_GMM()
for _ in range(3):
    P.GMM_opt(max_iterations=10, tol=1e-3); P.Reg_opt(tol=1e-3, nmax=1)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    P.GMM_opt(max_iterations=10, tol=1e-3); P.Reg_opt(tol=1e-3, nmax=1)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
