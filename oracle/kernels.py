"""Oracle: Gaussian-kernel reductions (TEST INFRASTRUCTURE, never on the product path).

Restates, on the CPU with dense torch ops evaluated in row chunks, the ten
reductions of the reference's ``GaussKernel`` plus ``check_coverage``:

  reference formulas            /root/reference/diffICP/tools/kernel.py
    K, grad K, Laplacian K      :248-252 (header), :259-267
    KBase .. GenDKRed, LapKRed  :177-207
    HessKRed, GradLapKRed       :282-292
    check_coverage              :324-329 (KeOps branch :326; the torch branch is broken)

Conventions: x is (M,D) "i" points, y is (N,D) "j" points, z = x_i - y_j,
K = exp(-|z|^2 / (2 sigma^2)), s = 1/sigma^2.  Everything is dtype-agnostic so the
same code gives the fp32 CPU baseline and the fp64 gold values.  Row chunking is
bit-identical to the unchunked evaluation for these axis-1 reductions
(identity 4) and keeps the (M,N,D) temporaries bounded.
"""

from __future__ import annotations

import torch


class GaussOracle:
    """Dense CPU evaluation of the reference's Gaussian kernel reductions."""

    def __init__(self, sigma: float, D: int, chunk: int = 2048):
        self.sigma = float(sigma)
        self.D = int(D)
        self.chunk = int(chunk)

    # -- pairwise building blocks -------------------------------------------------
    def _pairs(self, x, y):
        """z (m,N,D), r2 (m,N), K (m,N) for a row block x (m,D)."""
        z = x.unsqueeze(1) - y.unsqueeze(0)
        r2 = (z * z).sum(-1)
        K = torch.exp(-r2 / (2.0 * self.sigma ** 2))
        return z, r2, K

    def _rows(self, x, fn, *rowargs):
        """Apply fn(block of x, blocks of row-indexed args) chunk by chunk and stack."""
        M = x.shape[0]
        if M <= self.chunk:
            return fn(x, *rowargs)
        out = []
        for a in range(0, M, self.chunk):
            b = min(M, a + self.chunk)
            out.append(fn(x[a:b], *[r[a:b] for r in rowargs]))
        return torch.cat(out, 0)

    # -- i-indexed reductions (sum over j) ----------------------------------------
    def KBase(self, x, y):                       # kernel.py:178-179
        return self._rows(x, lambda xb: self._pairs(xb, y)[2].sum(1))

    def KRedScal(self, x, y, d):                 # kernel.py:182-183
        return self._rows(x, lambda xb: self._pairs(xb, y)[2] @ d)

    def KRed(self, x, y, b):                     # kernel.py:186-187
        return self._rows(x, lambda xb: self._pairs(xb, y)[2] @ b)

    def GradKRed(self, x, y):                    # kernel.py:190-191, gradK = -z K / sigma^2
        s = 1.0 / self.sigma ** 2

        def f(xb):
            z, _, K = self._pairs(xb, y)
            return -s * torch.einsum("mn,mnd->md", K, z)
        return self._rows(x, f)

    def DDKRed(self, x, y, b):                   # kernel.py:198-199
        s = 1.0 / self.sigma ** 2

        def f(xb):
            z, _, K = self._pairs(xb, y)
            return -s * torch.einsum("mn,mnd,nd->md", K, z, b)
        return self._rows(x, f)

    def GenDKRed(self, x, y, b, c):              # kernel.py:202-203
        s = 1.0 / self.sigma ** 2

        def f(xb, cb):
            z, _, K = self._pairs(xb, y)
            w = cb @ b.t()                        # (m,N) = c_i . b_j
            return -s * torch.einsum("mn,mnd->md", K * w, z)
        return self._rows(x, f, c)

    def LapKRed(self, x, y):                     # kernel.py:206-207 with :265-267
        s = 1.0 / self.sigma ** 2

        def f(xb):
            _, r2, K = self._pairs(xb, y)
            return (K * (s * s * r2 - self.D * s)).sum(1)
        return self._rows(x, f)

    def HessKRed(self, x, y, b, c):              # kernel.py:284-286, e = c_i - b_j
        s = 1.0 / self.sigma ** 2

        def f(xb, cb):
            z, _, K = self._pairs(xb, y)
            e = cb.unsqueeze(1) - b.unsqueeze(0)
            ze = (z * e).sum(-1)
            return torch.einsum("mn,mnd->md", K, s * s * ze.unsqueeze(-1) * z - s * e)
        return self._rows(x, f, c)

    def GradLapKRed(self, x, y):                 # kernel.py:289-292
        s = 1.0 / self.sigma ** 2

        def f(xb):
            z, r2, K = self._pairs(xb, y)
            a = s ** 3 * r2 - (self.D + 2) * s * s
            return -torch.einsum("mn,mnd->md", K * a, z)
        return self._rows(x, f)

    # -- j-indexed reduction (sum over i) -------------------------------------------
    def GradKRed_rev(self, x, y, d):             # kernel.py:194-195 -> (N,)
        s = 1.0 / self.sigma ** 2
        out = torch.zeros(y.shape[0], dtype=x.dtype)
        for a in range(0, x.shape[0], self.chunk):
            z, _, K = self._pairs(x[a:a + self.chunk], y)
            out = out + (-s) * torch.einsum("mn,mnd,md->n", K, z, d[a:a + self.chunk])
        return out

    # -- coverage test -------------------------------------------------------------
    def check_coverage(self, X, Y, Rthreshold):  # kernel.py:324-326
        thr = (Rthreshold * self.sigma) ** 2

        def f(xb):
            return self._pairs(xb, Y)[1].min(dim=1).values > thr
        return self._rows(X, f)

    # -- min_j squared distance (helper for tests of the coverage kernel) -----------
    def min_sqdist(self, X, Y):
        return self._rows(X, lambda xb: self._pairs(xb, Y)[1].min(dim=1).values)
