"""First-order VJPs of the standalone kernel reductions, expressed as kernel sums.

Only first-order derivatives exist on the hot path (no create_graph anywhere in the reference,
core/LDDMM.py:157 is a comment), so the VJP outputs are not themselves differentiable.
"""

from __future__ import annotations

import torch

from .. import ops


def _k(kern, sel, x, y, b=None, c=None, d=None):
    return ops.ksum(sel, kern.sigma, x, y, b=b, c=c, d=d)[sel]


def vjp(kern, sel, g, x, y, b, c, d, need):
    """Returns (gx, gy, gb, gc, gd); entries not needed (or not applicable) are None."""
    nx, ny, nb, nc, nd = need
    gx = gy = gb = gc = gd = None
    if sel == ops.K_RED:
        # out_i = sum_j K b_j  (tools/kernel.py:138):  d/db_j = sum_i K g_i ; d/dx_i = sum_j gradK (g_i.b_j) ;
        # d/dy_j = - (same pair term) = sum_i gradK(y_j - x_i) (b_j.g_i)
        if nb:
            gb = _k(kern, ops.K_RED, y, x, b=g)
        if nx:
            gx = _k(kern, ops.K_GEND, x, y, b=b, c=g)
        if ny:
            gy = _k(kern, ops.K_GEND, y, x, b=g, c=b)
        return gx, gy, gb, gc, gd
    raise NotImplementedError(
        f"VJP of kernel reduction selector {sel} is not available yet; gradients of the LDDMM path go through the "
        f"fused shooting adjoint (diff_icp_b200.shooting), which does not need it")
