"""GPU: the symmetric (q,q) adjoint engine (csrc/sym_engine.cuh: every unordered pair evaluated once, ring of column
accumulators) against the general engine (every ordered pair), itself parity-tested against the reference's gradients."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def adjoint(D, withlogdet, sigma, eta, q, p, a, u, gc, mode):
    """mode = the per-call engine argument of the C ABI (DICP_ENGINE_GENERAL / _SYMMETRIC / _SYMMETRIC_ALL)."""
    from diff_icp_b200 import ops
    M = q.shape[0]
    gq, gp = torch.zeros_like(q), torch.zeros_like(q)
    ws = ops.alloc_workspace(M, M, q.device)
    ops.rhs_adjoint(D, withlogdet, sigma, eta, q, p, None, a, u, None, gc, gq, gp, None, ws, engine=mode)
    torch.cuda.synchronize()
    return gq, gp


def forward(D, withlogdet, sigma, eta, q, p, mode):
    from diff_icp_b200 import ops
    M = q.shape[0]
    vq, dp, scal = torch.zeros_like(q), torch.zeros_like(q), torch.zeros(4, device=q.device)
    ws = ops.alloc_workspace(M, M, q.device)
    ops.rhs_forward(D, withlogdet, sigma, eta, q, p, None, vq, dp, None, scal, ws, engine=mode)
    torch.cuda.synchronize()
    return vq, dp, scal


@pytest.mark.parametrize("D", [2, 3])
@pytest.mark.parametrize("model", ["classic", "hybrid", "logdet"])
@pytest.mark.parametrize("M", [4096, 4097, 6999, 20000])
def test_symmetric_forward_matches_general_engine(D, model, M):
    if M == 20000 and D == 2:
        pytest.skip("large case once per model")
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + D)
    q = torch.rand(M, D, generator=g).to(dev)
    p = torch.randn(M, D, generator=g).to(dev)
    sigma = 0.2 if M > 10000 else 0.3
    eta = 0.02 if model == "logdet" else 0.0
    wld = model != "classic"
    ref = forward(D, wld, sigma, eta, q, p, 0)
    got = forward(D, wld, sigma, eta, q, p, 2)          # mode 2: the forward (q,q) pass on the symmetric engine too
    for r, s in zip(ref[:2], got[:2]):
        scale = float(r.abs().max())
        assert torch.isfinite(s).all()
        assert float((r - s).abs().max()) <= 2e-5 * scale, (float((r - s).abs().max()), scale)
    assert float((ref[2] - got[2]).abs().max()) <= 2e-5 * float(ref[2].abs().max()) + 1e-6       # dcost, A, B, C
    again = forward(D, wld, sigma, eta, q, p, 2)
    assert all(torch.equal(a, b) for a, b in zip(again, got))


@pytest.mark.parametrize("D", [2, 3])
@pytest.mark.parametrize("model", ["classic", "hybrid", "logdet"])
@pytest.mark.parametrize("M", [4096, 4097, 6999, 20000])
def test_symmetric_engine_matches_general_engine(D, model, M):
    if M == 20000 and D == 2:
        pytest.skip("large case once per model")
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + D)
    q = torch.rand(M, D, generator=g).to(dev)
    p, a, u = (torch.randn(M, D, generator=g).to(dev) for _ in range(3))
    gc = torch.tensor([0.7], device=dev)
    sigma = 0.2 if M > 10000 else 0.3
    eta = 0.02 if model == "logdet" else 0.0
    wld = model != "classic"
    ref = adjoint(D, wld, sigma, eta, q, p, a, u, gc, 0)
    got = adjoint(D, wld, sigma, eta, q, p, a, u, gc, 1)
    for r, s in zip(ref, got):
        scale = float(r.abs().max())
        assert torch.isfinite(s).all()
        assert float((r - s).abs().max()) <= 2e-5 * scale, (float((r - s).abs().max()), scale)
    # deterministic
    again = adjoint(D, wld, sigma, eta, q, p, a, u, gc, 1)
    assert torch.equal(again[0], got[0]) and torch.equal(again[1], got[1])


def test_small_and_huge_sizes_use_the_general_engine():
    from diff_icp_b200 import ops
    dev = torch.device("cuda:0")
    q = torch.rand(500, 3, device=dev)
    p, a, u = (torch.randn(500, 3, device=dev) for _ in range(3))
    r0 = adjoint(3, True, 0.3, 0.0, q, p, a, u, torch.ones(1, device=dev), 0)
    r1 = adjoint(3, True, 0.3, 0.0, q, p, a, u, torch.ones(1, device=dev), 1)
    assert torch.equal(r0[0], r1[0]) and torch.equal(r0[1], r1[1])          # below 4096 points: same engine, same bits


def adjoint_x(D, withlogdet, sigma, q, p, x, a, u, wx, gc, mode, eta=0.0):
    from diff_icp_b200 import ops
    M, Nx = q.shape[0], x.shape[0]
    gq, gp, gx = torch.zeros_like(q), torch.zeros_like(q), torch.zeros_like(x)
    ws = ops.alloc_workspace(max(M, Nx), max(M, Nx), q.device)
    ops.rhs_adjoint(D, withlogdet, sigma, eta, q, p, x, a, u, wx, gc, gq, gp, gx, ws, engine=mode)
    torch.cuda.synchronize()
    return gq, gp, gx


@pytest.mark.parametrize("D", [2, 3])
@pytest.mark.parametrize("withlogdet", [False, True, "logdet"])
@pytest.mark.parametrize("M,Nx", [(256, 2048), (257, 2049), (1584, 50000), (5000, 7001), (300, 100000)])
def test_fused_xq_adjoint_matches_two_pass_path(D, withlogdet, M, Nx):
    """(x,q) adjoint from ONE ring pass (rect_pair_kernel, Op AdjXQ) vs the two passes AdjXQx + AdjXQq of the general engine."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + Nx + D)
    q = torch.rand(M, D, generator=g).to(dev)
    x = torch.rand(Nx, D, generator=g).to(dev)
    p, a, u = (torch.randn(M, D, generator=g).to(dev) for _ in range(3))
    wx = torch.randn(Nx, D, generator=g).to(dev)
    gc = torch.tensor([-0.4], device=dev)
    sigma = 0.25
    eta = 0.03 if withlogdet == "logdet" else 0.0          # logdet model: the fused op is AdjXQE
    withlogdet = bool(withlogdet)
    ref = adjoint_x(D, withlogdet, sigma, q, p, x, a, u, wx, gc, 0, eta)
    got = adjoint_x(D, withlogdet, sigma, q, p, x, a, u, wx, gc, 1, eta)
    for r, s in zip(ref, got):
        scale = float(r.abs().max())
        assert torch.isfinite(s).all()
        assert float((r - s).abs().max()) <= 3e-5 * scale, (float((r - s).abs().max()), scale)
    again = adjoint_x(D, withlogdet, sigma, q, p, x, a, u, wx, gc, 1, eta)
    assert all(torch.equal(a_, b_) for a_, b_ in zip(again, got))


@pytest.mark.parametrize("D", [2, 3])
@pytest.mark.parametrize("withlogdet", [False, True, "logdet"])
@pytest.mark.parametrize("M,Nx", [(25, 10000), (7, 300), (33, 5111), (64, 2048), (1, 129)])
def test_small_support_ring_adjoint_matches_two_sided_kernel(D, withlogdet, M, Nx):
    """One-launch adjoint stage for small supports: ring form (every (x,q) pair once) vs the x-row / q-row form."""
    from diff_icp_b200 import ops
    dev = torch.device("cuda:0")
    lib = ops.load()
    eta = 0.03 if withlogdet == "logdet" else 0.0
    withlogdet = bool(withlogdet)
    g = torch.Generator().manual_seed(M * 7 + Nx + D)
    S = 2 * M * D + Nx * D + 1
    state = torch.rand(S, generator=g).to(dev)
    state[M * D:2 * M * D] = 0.3 * torch.randn(M * D, generator=g).to(dev)
    lam = torch.randn(S, generator=g).to(dev)
    base = torch.randn(S, generator=g).to(dev)
    add = torch.randn(S, generator=g).to(dev)
    outs = []
    for mode in (0, 1):
        prev = lib.dicp_sym_mode(mode)
        try:
            out, G = torch.zeros(S, device=dev), torch.zeros(S, device=dev)
            ws = ops.alloc_small_workspace(M, Nx, dev)
            for _ in range(2):                                   # twice: the ticket counters must reset themselves
                ops.small_adj_step(D, withlogdet, 0.3, eta, M, Nx, state, lam, base, None, add, 0.1, 0.0, out, G, ws)
            torch.cuda.synchronize()
            outs.append((out, G))
        finally:
            lib.dicp_sym_mode(prev)
    for a, b in zip(outs[0], outs[1]):
        scale = float(a.abs().max())
        assert torch.isfinite(b).all()
        assert float((a - b).abs().max()) <= 3e-5 * scale, (float((a - b).abs().max()), scale)


@pytest.mark.parametrize("model", ["classic", "hybrid", "logdet"])
@pytest.mark.parametrize("M", [66000, 70001])
def test_blocked_symmetric_adjoint_beyond_65536_points(model, M):
    """Above 65 536 points the (q,q) adjoint runs super-block by super-block (symmetric kernel inside a block, rectangular
    ring between blocks, outputs accumulated in a fixed order); 66 000 leaves a short last block for the general engine."""
    D = 3
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M)
    q = torch.rand(M, D, generator=g).to(dev)
    p, a, u = (torch.randn(M, D, generator=g).to(dev) for _ in range(3))
    gc = torch.tensor([0.3], device=dev)
    eta = 0.02 if model == "logdet" else 0.0
    wld = model != "classic"
    ref = adjoint(D, wld, 0.1, eta, q, p, a, u, gc, 0)
    got = adjoint(D, wld, 0.1, eta, q, p, a, u, gc, 1)
    for r, s in zip(ref, got):
        scale = float(r.abs().max())
        assert torch.isfinite(s).all()
        assert float((r - s).abs().max()) <= 3e-5 * scale, (float((r - s).abs().max()), scale)
    again = adjoint(D, wld, 0.1, eta, q, p, a, u, gc, 1)
    assert torch.equal(again[0], got[0]) and torch.equal(again[1], got[1])
