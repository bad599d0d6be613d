"""Helpers of the comparator algorithm (reference: ``diffICP/core/PSR_standard.py``).

Only ``data_distance`` is part of the B200 build: the RKHS distance between point clouds seen as signed measures, the data
term of the classical LDDMM point-set registration (``MultiPSR_std`` itself -- the comparator algorithm -- is outside the
hot path and is not built).  It runs on the kernel-sum engine (``KBase`` / ``KRedScal``), is differentiable through the
kernel-sum VJPs like on the reference's path, and accepts CUDA tensors only.
"""

from __future__ import annotations

from ..tools.kernel import GenKernel


def data_distance(Kernel: GenKernel, x, y, w=None):
    """|| mu_x - mu_y ||^2 in the RKHS of `Kernel`, mu_x = (1/Nx) sum_i delta_{x_i}, mu_y = (1/Ny) sum_j delta_{y_j} or
    sum_j w_j delta_{y_j} (reference: core/PSR_standard.py:37-58):
        L = sum_ij K(x_i,x_j)/Nx^2 + sum_ij c_i c_j K(y_i,y_j) - 2 sum_ij c_i K(y_i,x_j)/Nx,   c = 1/Ny or w."""
    KB = Kernel.KBase
    Nx, Ny = x.shape[0], y.shape[0]
    if w is None:
        return KB(x, x).sum() / Nx ** 2 + KB(y, y).sum() / Ny ** 2 - 2 * KB(y, x).sum() / (Nx * Ny)
    KRS = Kernel.KRedScal
    return KB(x, x).sum() / Nx ** 2 + (KRS(y, y, w).flatten() * w).sum() - 2 * (KB(y, x).flatten() * w).sum() / Nx
