"""Row-subset VJP of the fused right-hand side through the oracle (shared by the CPU self-check and the GPU tests)."""
import torch


def oracle_subset_grad(OR, q, p, a, u, g, R):
    """fp64 autograd gradient of L w.r.t. (q_R, p_R).  q, p, a, u: (M, D) fp64 CPU tensors; R: LongTensor of rows."""
    M = q.shape[0]
    qR = q[R].clone().requires_grad_(True)
    pR = p[R].clone().requires_grad_(True)
    qf = q.clone().index_put((R,), qR)
    pf = p.clone().index_put((R,), pR)
    rest = torch.ones(M, dtype=torch.bool)
    rest[R] = False

    def part(rq, rp, ra, ru, cq, cp):
        vq = OR.v(rq, cq, cp)
        Gq = OR.K.GenDKRed(rq, cq, cp, rp)
        if OR.eta != 0:
            Gq = Gq - OR.eta * OR.K.HessKRed(rq, cq, cp, rp) - OR.eta ** 2 * OR.K.GradLapKRed(rq, cq)
        out = (ra * vq).sum() - (ru * Gq).sum()
        if OR.withlogdet:                                   # dcost = mdivsum(q,q,p) (core/LDDMM.py:207-225), row-wise
            out = out + g * (rp * OR.K.GradKRed(rq, cq)).sum()
            if OR.eta != 0:
                out = out + g * OR.eta * OR.K.LapKRed(rq, cq).sum()
        return out

    L = part(qR, pR, a[R], u[R], qf, pf) + part(q[rest], p[rest], a[rest], u[rest], qR, pR)
    return torch.autograd.grad(L, [qR, pR])


def oracle_subset_grad_lowmem(OR, q, p, a, u, R, block=200_000):
    """Same as oracle_subset_grad for the classic model, the second part (10^6 rows x 32 columns) evaluated in row blocks
    so that autograd never holds more than block x 32 x D intermediates."""
    qR = q[R].clone().requires_grad_(True)
    pR = p[R].clone().requires_grad_(True)
    qf = q.clone().index_put((R,), qR)
    pf = p.clone().index_put((R,), pR)
    rest = torch.ones(q.shape[0], dtype=torch.bool)
    rest[R] = False
    vq = OR.v(qR, qf, pf)
    Gq = OR.K.GenDKRed(qR, qf, pf, pR)
    L = (a[R] * vq).sum() - (u[R] * Gq).sum()
    gq, gp = torch.autograd.grad(L, [qR, pR])
    idx = torch.nonzero(rest).flatten()
    old = OR.K.chunk
    OR.K.chunk = block
    for s in range(0, idx.numel(), block):
        rows = idx[s:s + block]
        L2 = (a[rows] * OR.v(q[rows], qR, pR)).sum() - (u[rows] * OR.K.GenDKRed(q[rows], qR, pR, p[rows])).sum()
        dq, dp = torch.autograd.grad(L2, [qR, pR])
        gq, gp = gq + dq, gp + dp
    OR.K.chunk = old
    return gq, gp
