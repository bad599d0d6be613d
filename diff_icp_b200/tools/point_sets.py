"""Setup-time helpers on point sets (reference: tools/point_sets.py): nearest-neighbour scale,
greedy decimation and the blurred-measure distance, on the CUDA kernels of csrc/pointset.cuh and the kernel-sum engine.
CUDA tensors only, like every other compute entry point."""

import math
import warnings

import numpy as np
import torch


def _check(x):
    from .._lib import require_cuda
    require_cuda(x)
    if x.dim() != 2 or x.shape[1] not in (2, 3):
        raise ValueError("point sets must be (n,2) or (n,3)")
    return x.detach().contiguous()


def min2_sqdist(x):
    """(N,) second smallest squared distance from each point to the points of its own set (the smallest is the point
    itself): the `Kmin(2, dim=1)[:, 1]` reduction of the reference (tools/point_sets.py:23-24), as one CUDA kernel."""
    from .._lib import check, load, ptr, stream_ptr
    x = _check(x)
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().dicp_min2_sqdist(x.shape[1], ptr(x), x.shape[0], ptr(out), stream_ptr()), "dicp_min2_sqdist")
    return out


def intrinsic_scale(x):
    """sqrt(mean_i min_{j != i} |x_i - x_j|^2): mean nearest-neighbour distance scale (reference: point_sets.py:13-26)."""
    return float(min2_sqdist(x).mean().sqrt())


def decimate(x, R):
    """Greedy decimation with radius R (reference: point_sets.py:102-133): repeatedly keep the not-yet-covered point
    with most not-yet-covered neighbours (first index on ties) and cover its neighbours.  Returns (kept, rejected) index
    lists, kept in pick order like the reference.  Runs on the device, one launch per pick (csrc/pointset.cuh): no dense
    N x N matrix, no host loop over points."""
    import ctypes
    from .._lib import check, load, ptr, stream_ptr
    x = _check(x)
    n, D = x.shape
    if n == 0:
        return [], []
    lib = load()
    ws = torch.empty(int(lib.dicp_decimate_workspace_bytes(n)), dtype=torch.uint8, device=x.device)
    kept = torch.empty(n, dtype=torch.int32, device=x.device)
    status = (ctypes.c_int * 2)(0, 0)
    restart, batch = 1, 32
    with torch.cuda.device(x.device):
        while True:
            check(lib.dicp_decimate_steps(D, ptr(x), n, float(R), restart, batch, ptr(kept), ptr(ws), ws.numel(), stream_ptr()),
                  "dicp_decimate_steps")
            check(lib.dicp_decimate_status(ptr(ws), n, status, stream_ptr()), "dicp_decimate_status")
            restart = 0
            if status[1]:
                break
            batch = min(4 * batch, 1024)
    ids = kept[:status[0]].tolist()
    chosen = np.zeros(n, dtype=bool)
    chosen[ids] = True
    return ids, np.flatnonzero(~chosen).tolist()


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
