"""A few calls of the EM sweeps at the atlas shape (640k points x 50 components, 2-D) and the few-component shape
(4M x 8, 3-D) for ncu:  ncu --set full -k regex:em_ python scripts/profile_em.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_icp_b200 import em_ops  # noqa: E402

dev = torch.device("cuda:0")
for N, C, D, sig in ((640000, 50, 2, 0.05), (4000000, 8, 3, 0.2)):
    g = torch.Generator().manual_seed(7)
    X = torch.rand(N, D, generator=g).to(dev)
    mu = torch.rand(C, D, generator=g).to(dev)
    w = torch.zeros(C, device=dev)
    lgn = D * (math.log(sig) + 0.5 * math.log(2 * math.pi))
    wl2 = ((w - torch.logsumexp(w, 0) - lgn) * 1.4426950408889634).contiguous()
    lpi = (w - torch.logsumexp(w, 0)).contiguous()
    for _ in range(2):
        st = em_ops.lse_colstats(sig, X, mu, wl2)
        out = em_ops.rowpass(sig, X, mu, wl2, mu, lpi)
    torch.cuda.synchronize()
    print(N, C, D, float(st[:, 1].sum()), float(out[2].sum()))
