"""Gaussian kernel reductions on the B200: drop-in for the reference's ``diffICP/tools/kernel.py``.

Same class names, constructor signature, attribute names and per-reduction call signatures as the reference
(``GaussKernel(sigma, D, computversion="keops", spec=defspec)``, tools/kernel.py:254-336), so that
``LDDMMModel`` / ``DiffPSR`` code written against the reference runs unchanged.  The ten reductions that the
reference builds symbolically with KeOps (tools/kernel.py:125-168) are served by ONE hand-written tiled
sm_100a kernel family (csrc/ops_ksum.cuh through the C ABI ``dicp_ksum``).

``computversion``: the reference's seam rebinds the ten reduction attributes to their "keops" or "torch"
implementation (tools/kernel.py:91-110).  Here every accepted value -- "b200" (new), "keops", "torch" -- binds the
same CUDA kernels: there is no KeOps, no dense-torch and no CPU path in this package.

Shapes follow the reference's torch twins: scalar-valued reductions return ``(M,)`` (KeOps returned ``(M,1)``;
every caller reduces with ``.sum()`` / ``.flatten()`` / ``.any()``).
"""

from __future__ import annotations

import warnings

import numpy as np
import torch

from .. import ops
from .spec import defspec, getspec

_ACCEPTED = ("b200", "keops", "torch")


def SVDpow(M, alpha, rcond=None):
    """SVD-based (pseudo-)power of a hermitian matrix (reference: tools/kernel.py:31-44). Setup-time helper."""
    U, S, Vh = torch.linalg.svd(M)
    keep = S > rcond * S[0] if rcond is not None else torch.ones_like(S, dtype=torch.bool)
    return U[:, keep] @ torch.diag(S[keep] ** alpha) @ Vh[keep, :]


class _Ksum(torch.autograd.Function):
    """One kernel-sum sweep with first-order VJPs, themselves kernel sums."""

    @staticmethod
    def forward(ctx, kern, sel, x, y, b, c, d):
        out = ops.ksum(sel, kern.sigma, x, y, b=b, c=c, d=d)[sel]
        ctx.kern, ctx.sel = kern, sel
        ctx.save_for_backward(x, y, b, c, d)
        return out

    @staticmethod
    def backward(ctx, g):
        from . import kernel_vjp
        x, y, b, c, d = ctx.saved_tensors
        gx, gy, gb, gc, gd = kernel_vjp.vjp(ctx.kern, ctx.sel, g.contiguous(), x, y, b, c, d, ctx.needs_input_grad[2:])
        return None, None, gx, gy, gb, gc, gd


class GenKernel:
    """Base class holding the reduction entry points (reference: tools/kernel.py:58-242)."""

    def __init__(self, D, computversion="keops"):
        self.computversion = None
        self.set_computversion(computversion)

    def set_computversion(self, version):
        if version not in _ACCEPTED:
            raise ValueError(f"unkown computversion : {version}. Choices are 'b200' (or the reference's 'keops' / 'torch', "
                             f"which bind the same CUDA kernels here)")
        self.computversion = version

    # pickling: same hook as the reference (tools/kernel.py:334-336)
    def __setstate__(self, state):
        self.__dict__.update(state)
        self.spec = defspec

    def _run(self, sel, x, y, b=None, c=None, d=None):
        getspec(*(t for t in (x, y, b, c, d) if t is not None))
        return _Ksum.apply(self, sel, x, y, b, c, d)

    # ---- the ten reductions (USAGE lines as in tools/kernel.py:130-168) ---------------------------------
    def KBase(self, x, y):
        """(M,)   X(i) = sum_j K(x_i-y_j)"""
        return self._run(ops.K_BASE, x, y)

    def KRedScal(self, x, y, d):
        """(M,)   X(i) = sum_j K(x_i-y_j) d_j"""
        return self._run(ops.K_REDSCAL, x, y, d=d.reshape(-1))

    def KRed(self, x, y, b):
        """(M,D)  X(i,:) = sum_j K(x_i-y_j) b_j"""
        return self._run(ops.K_RED, x, y, b=b)

    def GradKRed(self, x, y):
        """(M,D)  X(i,:) = sum_j (grad K)(x_i-y_j)"""
        return self._run(ops.K_GRAD, x, y)

    def GradKRed_rev(self, x, y, d):
        """(N,)   Y(j) = sum_i (grad K)(x_i-y_j) . d_i   (reduction over i)"""
        # rows = y, columns = x:  (grad K)(x_i - y_j).d_i = s K (y_j - x_i).d_i
        return self._run(ops.K_DOT, y, x, b=d)

    def DDKRed(self, x, y, b):
        """(M,D)  X(i,d) = sum_j (d_d K)(x_i-y_j) b_j^d"""
        return self._run(ops.K_DD, x, y, b=b)

    def GenDKRed(self, x, y, b, c):
        """(M,D)  X(i,:) = sum_j (grad K)(x_i-y_j) (c_i . b_j)"""
        return self._run(ops.K_GEND, x, y, b=b, c=c)

    def HessKRed(self, x, y, b, c):
        """(M,D)  X(i,:) = sum_j (Hess K)(x_i-y_j) (c_i - b_j)"""
        return self._run(ops.K_HESS, x, y, b=b, c=c)

    def LapKRed(self, x, y):
        """(M,)   X(i) = sum_j (Laplacian K)(x_i-y_j)"""
        return self._run(ops.K_LAP, x, y)

    def GradLapKRed(self, x, y):
        """(M,D)  X(i,:) = sum_j (grad Laplacian K)(x_i-y_j)"""
        return self._run(ops.K_GRADLAP, x, y)


class GaussKernel(GenKernel):
    """K(z) = exp(-|z|^2 / (2 sigma^2)) and its reductions (reference: tools/kernel.py:254-336)."""

    def __init__(self, sigma, D, computversion="keops", spec=defspec):
        self.sigma = sigma
        self.D = D
        self.spec = spec
        super().__init__(D, computversion)

    # ---- dense helpers kept for the setup-time solves (reference: tools/kernel.py:259-267); O(M*N) memory ----
    def K_torch(self, x, y):
        return (-((x[:, None, :] - y[None, :, :]) ** 2).sum(-1) / (2 * self.sigma ** 2)).exp()

    # ---- coverage (reference: tools/kernel.py:324-329; the KeOps branch :326 is the specification) ------------
    def min_sqdist(self, X, Y):
        return ops.ksum(ops.K_MINSQ, self.sigma, X, Y)[ops.K_MINSQ]

    def check_coverage(self, X, Y, Rthreshold):
        """bool (M,): True where X_i is farther than Rthreshold*sigma from every Y_j."""
        return self.min_sqdist(X, Y) > (Rthreshold * self.sigma) ** 2

    # ---- linear solves with K(x,x): setup-time, NOT on the hot path ---------------------
    DENSE_SOLVE_MAX = 4000          # above this the reference's dense O(M^3) CPU lstsq is replaced by the matrix-free solve
    PINV_MAX_RANK = 8192            # largest retained subspace of the matrix-free truncated pseudo-inverse

    def _Kmatmat(self, x, V):
        """K(x,x) @ V for V (M, r): the KRed kernel as mat-vec, D columns per sweep."""
        M, D = x.shape
        r = V.shape[1]
        out = torch.empty(M, r, dtype=V.dtype, device=V.device)
        for j in range(0, r, D):
            w = min(D, r - j)
            blk = V[:, j:j + D]
            if w < D:
                blk = torch.cat((blk, torch.zeros(M, D - w, dtype=V.dtype, device=V.device)), 1)
            out[:, j:j + w] = ops.ksum(ops.K_RED, self.sigma, x, x, b=blk.contiguous())[ops.K_RED][:, :w]
        return out

    def top_eigenpairs(self, x, rcond, r0=96, power=2, seed=0):
        """Eigenpairs (lam (r,) descending, U (M,r)) of the symmetric PSD matrix K(x,x) that span every eigenvalue above
        rcond * lam_max: randomised subspace iteration with the tiled kernel sum as mat-vec (no M x M matrix), the rank
        doubled until the smallest Ritz value sits two decades under the cut-off.  A Gaussian kernel matrix has a
        super-exponentially decaying spectrum (a few hundred eigenvalues above 1e-3 * lam_max for sigma = 0.2 in the unit
        cube, whatever M), which is what makes this cheap: O(M^2 r / D) pair evaluations instead of the O(M^3) dense SVD."""
        M = x.shape[0]
        gen = torch.Generator(device=x.device).manual_seed(seed)
        r = min(M, r0)
        Q = torch.empty(M, 0, dtype=x.dtype, device=x.device)
        while True:
            Om = torch.randn(M, r - Q.shape[1], generator=gen, dtype=x.dtype, device=x.device)
            Q = torch.linalg.qr(torch.cat((Q, self._Kmatmat(x, Om)), 1)).Q      # previous subspace kept, new directions added
            for _ in range(power):
                Q = torch.linalg.qr(self._Kmatmat(x, Q)).Q
            B = Q.double().t() @ self._Kmatmat(x, Q).double()                    # Rayleigh-Ritz in fp64 (r x r)
            lam, W = torch.linalg.eigh(0.5 * (B + B.t()))
            lam, W = lam.flip(0), W.flip(1)
            if r >= M or r >= self.PINV_MAX_RANK or float(lam[-1]) < 1e-2 * rcond * float(lam[0]):
                return lam, (Q.double() @ W)
            r = min(M, 2 * r)

    def KpinvSolve(self, x, v, rcond=None):
        """Least-squares b with sum_j K(x_i-x_j) b_j ~ v_i (reference: tools/kernel.py:227-232, numpy lstsq on the
        dense M x M matrix; rcond = singular values below rcond * s_max are dropped, i.e. a truncated pseudo-inverse).
        * v == 0  =>  b = 0 exactly (minimum-norm solution): the case that runs before every optimisation
          (DiffPSR.initialize_a0 / update_a0 with eta = 0 and zero initial speeds).
        * M <= DENSE_SOLVE_MAX: dense numpy lstsq, the reference's own call.
        * larger M (where the reference's O(M^3) CPU solve is infeasible): the SAME truncated
          pseudo-inverse, matrix-free: K is symmetric PSD, so its singular triplets are its eigenpairs;
          b = sum_{lam_k > rcond lam_1} u_k (u_k . v) / lam_k with the eigenpairs from `top_eigenpairs`.
          rcond = None (machine-precision cut-off: the retained rank is ~M) goes to conjugate gradients instead
          (`KridgeSolve_keops` with a ridge at eps * M * lam_max); its fitted speeds K b agree with the reference's,
          the momenta themselves are not determined at that cut-off (the reference's own fp32 and fp64 runs differ
          by orders of magnitude there, tests/golden/v2p.npz)."""
        if not bool(v.any()):
            return torch.zeros_like(v)
        if x.shape[0] <= self.DENSE_SOLVE_MAX:
            K_xx = self.K_torch(x, x)
            sol = np.linalg.lstsq(K_xx.detach().cpu().numpy(), v.detach().cpu().numpy(), rcond=rcond)[0]
            return torch.from_numpy(sol).to(**getspec(x, v))
        if rcond is None:
            lam, _ = self.top_eigenpairs(x, 0.5, r0=8)
            alpha = float(torch.finfo(x.dtype).eps) * x.shape[0] * float(lam[0])
            return self.KridgeSolve_keops(x, v, alpha=alpha)
        lam, U = self.top_eigenpairs(x, rcond)
        keep = lam > rcond * lam[0]
        U, lam = U[:, keep], lam[keep]
        return (U @ ((U.t() @ v.double()) / lam[:, None])).to(**getspec(x, v)).contiguous()

    def KridgeSolve_torch(self, x, v, alpha=1e-4):
        K_xx = self.K_torch(x, x)
        return torch.linalg.solve(K_xx + alpha * torch.eye(K_xx.shape[0], **getspec(x)), v)

    def KridgeSolve_keops(self, x, v, alpha=1e-4, tol=1e-6, maxiter=1000):
        """Matrix-free conjugate gradient on (K_xx + alpha I) b = v with the KRed kernel as mat-vec
        (what LazyTensor.solve does in the reference, tools/kernel.py:240-242)."""
        b = torch.zeros_like(v)
        r = v.clone()
        p = r.clone()
        rs = (r * r).sum()
        rs0 = float(rs)
        if rs0 == 0.0:
            return b
        for _ in range(maxiter):
            Ap = ops.ksum(ops.K_RED, self.sigma, x, x, b=p)[ops.K_RED] + alpha * p
            a = rs / (p * Ap).sum()
            b = b + a * p
            r = r - a * Ap
            rs_new = (r * r).sum()
            if float(rs_new) <= tol * tol * rs0:
                break
            p = r + (rs_new / rs) * p
            rs = rs_new
        return b
