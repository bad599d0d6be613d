"""First-order VJPs of the standalone kernel reductions, themselves evaluated as kernel sums on
the device.  Only first-order derivatives exist on the reference's path (no create_graph anywhere, core/LDDMM.py:157 is
a comment), so the VJP outputs are not differentiable again.

The LDDMM optimisation path does NOT go through these: ``LDDMMModel.Shoot`` carries its own fused adjoint
(diff_icp_b200/shooting.py).  They serve direct differentiable use of ``GaussKernel`` (e.g. kernel data distances).
Provided: KBase, KRedScal, KRed, GradKRed (all inputs), LapKRed (x).  The remaining reductions are only ever
differentiated inside the fused adjoint kernels; asking for their standalone VJP raises NotImplementedError.
"""

from __future__ import annotations

import torch

from .. import ops


def _k(kern, sel, x, y, b=None, c=None, d=None):
    return ops.ksum(sel, kern.sigma, x, y, b=b, c=c, d=d)[sel]


def _grad_weighted(kern, rows, cols, wcol):
    """sum_col gradK(row - col) * w_col  -> (rows, D), via GenDKRed with c = e_1 and b = (w, 0, ..)."""
    D = rows.shape[1]
    b = torch.zeros(cols.shape[0], D, dtype=rows.dtype, device=rows.device)
    b[:, 0] = wcol
    c = torch.zeros(rows.shape[0], D, dtype=rows.dtype, device=rows.device)
    c[:, 0] = 1.0
    return _k(kern, ops.K_GEND, rows, cols, b=b, c=c)


def vjp(kern, sel, g, x, y, b, c, d, need):
    """Returns (gx, gy, gb, gc, gd); entries not needed (or not applicable) are None."""
    nx, ny, nb, nc, nd = need
    gx = gy = gb = gc = gd = None
    if sel == ops.K_RED:
        # out_i = sum_j K b_j (tools/kernel.py:138): d/db_j = sum_i K g_i; d/dx_i = sum_j gradK (g_i.b_j); d/dy_j = -(pair term)
        if nb:
            gb = _k(kern, ops.K_RED, y, x, b=g)
        if nx:
            gx = _k(kern, ops.K_GEND, x, y, b=b, c=g)
        if ny:
            gy = _k(kern, ops.K_GEND, y, x, b=g, c=b)
    elif sel == ops.K_BASE:
        # out_i = sum_j K: d/dx_i = g_i sum_j gradK(x_i-y_j); d/dy_j = sum_i g_i gradK(y_j-x_i)
        if nx:
            gx = g[:, None] * _k(kern, ops.K_GRAD, x, y)
        if ny:
            gy = _grad_weighted(kern, y, x, g)
    elif sel == ops.K_REDSCAL:
        # out_i = sum_j K d_j (tools/kernel.py:135)
        if nd:
            gd = _k(kern, ops.K_REDSCAL, y, x, d=g)
        if nx:
            gx = g[:, None] * _grad_weighted(kern, x, y, d)
        if ny:
            gy = d[:, None] * _grad_weighted(kern, y, x, g)
    elif sel == ops.K_GRAD:
        # out_i = sum_j gradK(z): d/dx_i = sum_j HessK(z) g_i; d/dy_j = -sum_i HessK(z) g_i   (HessK is even in z)
        zx = torch.zeros_like(y)
        if nx:
            gx = _k(kern, ops.K_HESS, x, y, b=zx, c=g)
        if ny:
            gy = _k(kern, ops.K_HESS, y, x, b=g, c=torch.zeros_like(y))
    elif sel == ops.K_LAP and not ny:
        if nx:
            gx = g[:, None] * _k(kern, ops.K_GRADLAP, x, y)
    else:
        raise NotImplementedError(
            f"standalone VJP of kernel reduction selector {sel} is not provided; gradients of the LDDMM path go through "
            f"the fused shooting adjoint (diff_icp_b200.shooting)")
    return gx, gy, gb, gc, gd
