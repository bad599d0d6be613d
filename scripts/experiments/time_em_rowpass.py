"""Device time of the full EM row pass (targets + free-energy sums) and of the fused first sweep at the three shapes of the bench
(atlas 640k x 50 2-D, configs[3] structure 1.07M x 20 3-D, few components 4M x 8 3-D): CUDA events around 20 back-to-back calls.
DICP_B200_LIB selects a library variant (tuning sweeps)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from diff_icp_b200 import em_ops
dev = torch.device("cuda:0")
res = []
for N, C, D, sig in ((640000, 50, 2, 0.05), (1066667, 20, 3, 0.1), (4000000, 8, 3, 0.2)):
    g = torch.Generator().manual_seed(7)
    X = torch.rand(N, D, generator=g).to(dev)
    mu = torch.rand(C, D, generator=g).to(dev)
    w = torch.zeros(C, device=dev)
    lgn = D * (math.log(sig) + 0.5 * math.log(2 * math.pi))
    wl2 = ((w - torch.logsumexp(w, 0) - lgn) * 1.4426950408889634).contiguous()
    lpi = (w - torch.logsumexp(w, 0)).contiguous()
    for name, fn in (("full", lambda: em_ops.rowpass(sig, X, mu, wl2, mu, lpi)), ("fused_first", lambda: em_ops.lse_colstats(sig, X, mu, wl2))):
        for _ in range(3):
            out = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            out = fn()
        e1.record(); torch.cuda.synchronize()
        res.append(f"{N}x{C} D{D} {name} {e0.elapsed_time(e1) / 20 * 1e3:.1f}us")
    chk = em_ops.rowpass(sig, X, mu, wl2, mu, lpi)
    res.append(f"chk {float(chk[2].sum()):.6g}")
print(os.environ.get("DICP_B200_LIB", "default"), " | ".join(res))
