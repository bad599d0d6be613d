"""One-off measurements of the other BASELINE.json configurations at full size (recorded under profiles/):
  C2  api.ICP_two_set, 3-D 20k x 20k, dense support, API-default (logdet) model: 2 outer iterations
  C5  large-kernel stress: 3-D, 10^6 control points, classic, Ralston nt = 10, one forward + backward (loss + gradient)
"""
import json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
out = {}
if which == "c2":
    from diff_icp_b200.api.ICP_two_set import ICP_two_set
    xA, y, _ = bench.make_workload(1234)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    PSR, evol = ICP_two_set(xA.to(dev), y.to(dev), {"sigma": 0.1, "optimize_sigma": True, "outlier_weight": None},
                            {"type": "diffeomorphic", "lambda_LDDMM": 500.0, "sigma_LDDMM": 0.2},
                            numerical_options={"support_LDDMM": {"scheme": "dense"}},
                            optim_options={"max_iterations": 1}, plotstuff=False, printstuff=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    PSR.LMi.use_cuda_graph = True
    times = []
    for it in range(2):
        torch.cuda.synchronize(); a = time.perf_counter()
        PSR.GMM_opt(max_iterations=10, tol=1e-3)
        torch.cuda.synchronize(); b = time.perf_counter()
        PSR.Reg_opt(tol=1e-3, nmax=1)
        torch.cuda.synchronize(); c = time.perf_counter()
        times.append((b - a, c - b))
    out = {"config": "C2 ICP_two_set 20k x 20k 3-D dense, logdet (API default), Euler nt=10", "setup_plus_first_iteration_s": t1 - t0,
           "gmm_opt_s": [t[0] for t in times], "reg_opt_s": [t[1] for t in times], "FE": PSR.FE, "sigma_gmm": PSR.GMMi[0].sigma,
           "a0_nonzero_init": bool(evol["a0"][0][0].abs().max() > 0)}
else:
    from diff_icp_b200.core.LDDMM import LDDMMModel
    M = 1_000_000
    g = torch.Generator().manual_seed(7)
    q = torch.rand(M, 3, generator=g).to(dev)
    p = (1e-3 * torch.randn(M, 3, generator=g)).to(dev).requires_grad_(True)
    yy = (q + 0.01).detach()
    LM = LDDMMModel(sigma=0.05, D=3, lambd=100.0, spec={"device": dev, "dtype": torch.float32}, version="classic", scheme="Ralston", nt=10)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sh = LM.Shoot(q, p)
    L = LM.trajloss(sh) + ((sh[-1][0] - yy) ** 2).sum() / 2
    torch.cuda.synchronize(); t1 = time.perf_counter()
    L.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    pairs = 20.0 * M * M
    out = {"config": "C5 stress: 3-D, 1e6 control points, classic, Ralston nt=10 (20 RHS evaluations), loss = trajloss + |q(1)-y|^2/2",
           "forward_s": t1 - t0, "backward_s": t2 - t1, "physical_pairs_per_s_forward": pairs / (t1 - t0),
           "physical_pairs_per_s_backward": pairs / (t2 - t1), "reference_equivalent_pairs_per_s_forward": 2 * pairs / (t1 - t0),
           "loss": float(L), "grad_norm": float(p.grad.norm())}
print(json.dumps(out))
