"""Oracle: set-up helpers on point sets (TEST INFRASTRUCTURE, never on the product path).

CPU restatements (dense torch / numpy, dtype-agnostic) of

  intrinsic_scale       /root/reference/diffICP/tools/point_sets.py:13-26   (KeOps Kmin(2): second smallest |x_i-x_j|^2)
  point_set_distance    /root/reference/diffICP/tools/point_sets.py:46-95
  decimate              /root/reference/diffICP/tools/point_sets.py:102-133 (greedy covering; index lists)
  data_distance         /root/reference/diffICP/core/PSR_standard.py:37-58    (RKHS distance of point clouds as measures)

Pinning: `decimate`, `point_set_distance` and `data_distance` are pinned against the reference's OWN functions, executed in the build
container from their source text by tests/golden/make_golden.py (the module itself cannot be imported: its line 8
hard-imports pykeops) -> tests/golden/pointsets.npz.  `intrinsic_scale` is one KeOps reduction that can be run nowhere:
parity unpinned for that function alone; it is restated from its definition (second smallest squared distance, the
smallest being the point itself) and cross-checked against the brute-force sort in the tests.
"""

from __future__ import annotations

import math

import numpy as np
import torch

from .kernels import GaussOracle


def min2_sqdist(x: torch.Tensor, chunk: int = 2048) -> torch.Tensor:
    """(N,) second smallest of {|x_i - x_j|^2, j = 0..N-1} (point_sets.py:23-24)."""
    out = []
    for a in range(0, x.shape[0], chunk):
        d2 = ((x[a:a + chunk, None, :] - x[None, :, :]) ** 2).sum(-1)
        out.append(d2.topk(2, dim=1, largest=False).values[:, 1])
    return torch.cat(out) if out else x.new_zeros(0)


def intrinsic_scale(x: torch.Tensor) -> float:
    """point_sets.py:13-26."""
    return float(min2_sqdist(x).mean().sqrt())


def decimate(x: torch.Tensor, R: float):
    """point_sets.py:102-133.  Returns (kept, rejected): kept in pick order, rejected ascending."""
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    near = (d2 <= R ** 2).numpy()
    n = x.shape[0]
    alive = np.ones(n, dtype=bool)
    kept = []
    while alive.any():
        ids = np.flatnonzero(alive)
        counts = near[np.ix_(ids, ids)].sum(axis=0)
        pick = int(ids[int(counts.argmax())])          # first maximum in ascending index order (:123-124)
        kept.append(pick)
        alive &= ~near[pick]                           # :126-127
    ks = set(kept)
    return kept, [i for i in range(n) if i not in ks]


def point_set_distance(X, Y, sigma_X=None, sigma_Y=None, w_X=None, w_Y=None) -> float:
    """point_sets.py:46-95: || f_X - f_Y ||_2^2 for the Gaussian-blurred point measures."""
    D = X.shape[1]
    sigma_X = intrinsic_scale(X) if sigma_X is None else sigma_X
    sigma_Y = intrinsic_scale(Y) if sigma_Y is None else sigma_Y
    # default weights in the DEFAULT dtype (fp32), exactly like the reference (:79-82), whatever the dtype of X and Y
    w_X = torch.ones(X.shape[0]) / X.shape[0] if w_X is None else w_X
    w_Y = torch.ones(Y.shape[0]) / Y.shape[0] if w_Y is None else w_Y
    w_X, w_Y = w_X.to(X.dtype), w_Y.to(Y.dtype)          # value-preserving promotion, as torch does in the reference
    sXX, sYY, sXY = math.sqrt(2) * sigma_X, math.sqrt(2) * sigma_Y, math.sqrt(sigma_X ** 2 + sigma_Y ** 2)
    c = lambda s: 1 / ((2 * math.pi) ** (D / 2) * s ** D)
    k = lambda s, a, b, w: GaussOracle(s, D).KRedScal(a, b, w).flatten()
    return float(c(sXX) * (k(sXX, X, X, w_X) * w_X).sum() + c(sYY) * (k(sYY, Y, Y, w_Y) * w_Y).sum()
                 - 2 * c(sXY) * (k(sXY, X, Y, w_Y) * w_X).sum())


def data_distance(sigma: float, x, y, w=None) -> float:
    """core/PSR_standard.py:37-58 with a Gaussian kernel of width sigma."""
    K = GaussOracle(sigma, x.shape[1])
    Nx, Ny = x.shape[0], y.shape[0]
    if w is None:
        return float(K.KBase(x, x).sum() / Nx ** 2 + K.KBase(y, y).sum() / Ny ** 2 - 2 * K.KBase(y, x).sum() / (Nx * Ny))
    return float(K.KBase(x, x).sum() / Nx ** 2 + (K.KRedScal(y, y, w).flatten() * w).sum()
                 - 2 * (K.KBase(y, x).flatten() * w).sum() / Nx)
