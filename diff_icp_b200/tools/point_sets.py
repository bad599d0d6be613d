"""Setup-time helpers on point sets (reference: tools/point_sets.py). Not on the hot path:
dense torch restatements, usable for the small sets they are called on (GMM centroids, decimated supports)."""

import math
import warnings

import numpy as np
import torch


def intrinsic_scale(x):
    """sqrt(mean_i min_{j != i} |x_i - x_j|^2): mean nearest-neighbour distance scale (reference: point_sets.py:13-26,
    KeOps Kmin(2))."""
    x = x.contiguous()
    n = x.shape[0]
    second = []
    for a in range(0, n, 4096):
        d2 = ((x[a:a + 4096, None, :] - x[None, :, :]) ** 2).sum(-1)
        second.append(d2.topk(2, dim=1, largest=False).values[:, 1])
    return float(torch.cat(second).mean().sqrt())


def decimate(x, R):
    """Greedy decimation with radius R (reference: point_sets.py:102-133): repeatedly keep the not-yet-covered point
    with most not-yet-covered neighbours. Returns (kept, rejected) index lists."""
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    near = (d2 <= R ** 2).cpu().numpy()
    n = x.shape[0]
    alive = np.ones(n, dtype=bool)
    kept = []
    while alive.any():
        ids = np.flatnonzero(alive)
        counts = near[np.ix_(ids, ids)].sum(axis=0)
        pick = ids[int(counts.argmax())]
        kept.append(int(pick))
        alive &= ~near[pick]
    kept_set = set(kept)
    return kept, [i for i in range(n) if i not in kept_set]


def point_set_distance(X, Y, sigma_X=None, sigma_Y=None, w_X=None, w_Y=None):
    """Squared L2 distance between the Gaussian-blurred point measures of X and Y (reference: point_sets.py:46-95)."""
    from .kernel import GaussKernel
    D = X.shape[1]
    sX_int, sY_int = intrinsic_scale(X), intrinsic_scale(Y)
    sigma_X = sX_int if sigma_X is None else sigma_X
    sigma_Y = sY_int if sigma_Y is None else sigma_Y
    if sigma_X < sX_int:
        warnings.warn("Required data distance scale `sigma_X` is smaller than 'intrinsic' scale for point set X. You should probably augment sigma_X.")
    if sigma_Y < sY_int:
        warnings.warn("Required data distance scale `sigma_Y` is smaller than 'intrinsic' scale for point set Y. You should probably augment sigma_Y.")
    if w_X is None:
        w_X = torch.ones(X.shape[0], dtype=X.dtype, device=X.device) / X.shape[0]
    if w_Y is None:
        w_Y = torch.ones(Y.shape[0], dtype=Y.dtype, device=Y.device) / Y.shape[0]
    sXX, sYY, sXY = math.sqrt(2) * sigma_X, math.sqrt(2) * sigma_Y, math.sqrt(sigma_X ** 2 + sigma_Y ** 2)
    norm = lambda s: 1 / ((2 * math.pi) ** (D / 2) * s ** D)
    kXX, kYY, kXY = (GaussKernel(s, D).KRedScal for s in (sXX, sYY, sXY))
    return norm(sXX) * (kXX(X, X, w_X).flatten() * w_X).sum() \
        + norm(sYY) * (kYY(Y, Y, w_Y).flatten() * w_Y).sum() \
        - 2 * norm(sXY) * (kXY(X, Y, w_Y).flatten() * w_X).sum()
