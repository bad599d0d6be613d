"""GPU: set-up kernels on point sets (csrc/pointset.cuh) -- index lists bit-exact against the reference's own decimate
(golden fixtures) and the oracle; nearest-neighbour scale and blurred-measure distance within fp32 tolerance."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to("cuda:0")


def test_decimate_indices_bit_exact_vs_reference(golden):
    from diff_icp_b200.tools.point_sets import decimate
    g = golden("pointsets")
    for tag in g["cases"]:
        tag = str(tag)
        kept, rej = decimate(cu(g[f"{tag}_x"]), float(g[f"{tag}_R"]))
        assert kept == g[f"{tag}_kept"].tolist(), tag
        assert rej == g[f"{tag}_rejected"].tolist(), tag


def test_decimate_larger_sets_vs_oracle_and_properties():
    from diff_icp_b200.tools.point_sets import decimate
    from oracle import pointsets
    g = torch.Generator().manual_seed(4)
    for D, n, R in ((2, 3000, 0.03), (3, 4000, 0.09), (2, 1, 0.1), (3, 130, 10.0)):
        x = torch.rand(n, D, generator=g)
        kept, rej = decimate(x.cuda(), R)
        ko, ro = pointsets.decimate(x, R)
        assert kept == ko and rej == ro, (D, n)
    # a size the dense oracle would not enjoy: covering + separation properties (kept points are pairwise farther than R
    # by construction of the greedy rule, every point is within R of a kept one)
    x = torch.rand(60000, 3, generator=g).cuda()
    R = 0.08
    kept, rej = decimate(x, R)
    assert len(kept) + len(rej) == 60000 and len(set(kept)) == len(kept)
    xk = x[kept]
    from diff_icp_b200.tools.kernel import GaussKernel
    K = GaussKernel(1.0, 3, spec={"device": x.device, "dtype": torch.float32})
    assert float(K.min_sqdist(x, xk).max()) <= R * R * (1 + 1e-5)
    d2 = ((xk[:, None] - xk[None]) ** 2).sum(-1) + 10 * torch.eye(len(kept), device=x.device)
    assert float(d2.min()) > R * R
    assert decimate(torch.empty(0, 2).cuda(), 0.1) == ([], [])


def test_min2_sqdist_and_intrinsic_scale():
    from diff_icp_b200.tools.point_sets import intrinsic_scale, min2_sqdist
    from oracle import pointsets
    g = torch.Generator().manual_seed(8)
    for D, n in ((2, 1500), (3, 2777), (3, 2)):
        x = torch.rand(n, D, generator=g)
        got = min2_sqdist(x.cuda()).cpu()
        assert torch.equal(got, pointsets.min2_sqdist(x))               # same operation order: bit-exact
        assert abs(intrinsic_scale(x.cuda()) - pointsets.intrinsic_scale(x.double())) <= 1e-6 * pointsets.intrinsic_scale(x.double())
    x = torch.rand(5, 2)
    x[3] = x[1]                                                          # duplicate point: second smallest is 0
    assert float(min2_sqdist(x.cuda())[1]) == 0.0


def test_point_set_distance_matches_reference(golden):
    from diff_icp_b200.tools.point_sets import point_set_distance
    g = golden("pointsets")
    X, Y = cu(g["psd_X"]), cu(g["psd_Y"])
    for tag, kw in (("auto", {}), ("fixed", {"sigma_X": 0.2, "sigma_Y": 0.15})):
        gold, ref32 = float(g[f"psd_{tag}_gold"]), float(g[f"psd_{tag}_ref32"])
        v = float(point_set_distance(X, Y, **kw))
        assert abs(v - gold) <= max(1e-5 * abs(gold), 2 * abs(ref32 - gold)) + 1e-7, (tag, v, gold, ref32)


def test_data_distance_matches_reference(golden):
    """RKHS distance of the comparator algorithm (core/PSR_standard.py:37-58) vs the reference's own function; the value is
    a difference of three O(1) sums that cancel to ~6e-3, hence the tolerance relative to the reference's fp32 error."""
    from diff_icp_b200.core.PSR_standard import data_distance
    from diff_icp_b200.tools.kernel import GaussKernel
    g = golden("pointsets")
    X, Y, w = cu(g["psd_X"]), cu(g["psd_Y"]), cu(g["dd_w"])
    K = GaussKernel(float(g["dd_sigma"]), 3, spec={"device": X.device, "dtype": torch.float32})
    for tag, args in (("plain", ()), ("weighted", (w,))):
        gold, ref32 = float(g[f"dd_{tag}_gold"]), float(g[f"dd_{tag}_ref32"])
        v = float(data_distance(K, X, Y, *args))
        assert abs(v - gold) <= max(2e-5 * abs(gold), 2 * abs(ref32 - gold)) + 2e-7, (tag, v, gold, ref32)
    # differentiable through the kernel-sum VJPs: d/dx against a finite difference along a random direction
    Xg = X.clone().requires_grad_(True)
    L = data_distance(K, Xg, Y)
    (gx,) = torch.autograd.grad(L, [Xg])
    d = torch.randn_like(X)
    eps = 1e-2
    fd = (float(data_distance(K, X + eps * d, Y)) - float(data_distance(K, X - eps * d, Y))) / (2 * eps)
    an = float((gx * d).sum())
    assert abs(fd - an) <= 5e-2 * abs(an) + 1e-6, (fd, an)


def test_cpu_tensors_are_refused():
    from diff_icp_b200.tools.point_sets import decimate, intrinsic_scale
    with pytest.raises(ValueError):
        decimate(torch.rand(10, 2), 0.1)
    with pytest.raises(ValueError):
        intrinsic_scale(torch.rand(10, 2))
