"""Oracle: isotropic-uniform GMM, EM step (TEST INFRASTRUCTURE, never on the product path).

Restates /root/reference/diffICP/core/GMM.py:

  E step, torch twin                    :263-283
  M step + values, torch twin           :286-325   (variant="torch")
  E / M / values, KeOps formulation     :402-496   (variant="keops"; cannot be executed anywhere --
                                                    pykeops is absent and unpinned -- so it is restated
                                                    from the published formulas; it coincides with the
                                                    torch twin when skip_M=True, which is how it is pinned)
  EM_optimization                       :330-357
  log_responsibilities                  :221-232
  log_ratio_to_proba                    :205-217
  intrinsic_scale                       /root/reference/diffICP/tools/point_sets.py:13-26

The two variants differ only in WHICH parameters are "current" when sigma and the
gaussian normalisation are evaluated:
  torch : sigma' from squared distances to the OLD mu;  loggaussnorm from the OLD sigma
  keops : sigma' from squared distances to the NEW mu;  loggaussnorm from the NEW sigma
"""

from __future__ import annotations

import math

import torch


def intrinsic_scale(x):
    """sqrt(mean_i of the 2nd smallest_j |x_i - x_j|^2)  (point_sets.py:22-26)."""
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    second = d2.topk(2, dim=1, largest=False).values[:, 1]
    return float(second.mean().sqrt())


def log_ratio_to_proba(eta):
    """(log p, log q) of a Bernoulli with log-odds eta  (GMM.py:205-217)."""
    if not torch.is_tensor(eta):
        eta = torch.tensor(float(eta), dtype=torch.float64)
    Z = torch.nn.functional.softplus(eta)
    return eta - Z, -Z


class GMMOracle:
    def __init__(self, mu, sigma, w=None, outliers=None, to_optimize=None, ensure_continuum=False):
        self.mu = mu.clone()
        self.C, self.D = mu.shape
        self.sigma = float(sigma)
        self.w = torch.zeros(self.C, dtype=mu.dtype) if w is None else w.clone()
        self.outliers = None if outliers is None else dict(outliers)   # {"vol0":..., "eta0":...}
        self.to_optimize = {"sigma": True, "mu": True, "w": True, "eta0": True}
        if to_optimize:
            self.to_optimize.update(to_optimize)
        self.ensure_continuum = ensure_continuum

    def _lgn(self, sigma):
        return self.D * (math.log(sigma) + 0.5 * math.log(2 * math.pi))

    def log_responsibilities(self, X):
        d2 = ((X[:, None, :] - self.mu[None, :, :]) ** 2).sum(-1)
        return torch.log_softmax(self.w[None, :] - d2 / (2 * self.sigma ** 2), dim=1)

    def em_step(self, X, skip_M=False, variant="keops"):
        N = X.shape[0]
        opt = self.to_optimize
        d2 = ((X[:, None, :] - self.mu[None, :, :]) ** 2).sum(-1)
        lgn_old = self._lgn(self.sigma)
        t = self.w[None, :] - torch.logsumexp(self.w, 0) - d2 / (2 * self.sigma ** 2) - lgn_old
        T = torch.logsumexp(t, dim=1)
        lg = t - T[:, None]
        g = lg.exp()

        if self.outliers is not None:
            if self.outliers.get("vol0") is None:
                self.outliers["vol0"] = float((X.max(0).values - X.min(0).values).prod())
            logJ0 = -math.log(self.outliers["vol0"])
            eta_n = self.outliers["eta0"] + logJ0 - T
            lg0, lgT = log_ratio_to_proba(eta_n)

        if not skip_M:
            if opt["mu"]:
                self.mu = torch.softmax(lg, dim=0).t() @ X
            if self.outliers is not None and opt["eta0"]:
                self.outliers["eta0"] = float(torch.logsumexp(lg0, 0) - torch.logsumexp(lgT, 0))
            if opt["w"]:
                self.w = torch.logsumexp(lg, dim=0)
            if opt["sigma"]:
                if variant == "keops":
                    d2s = ((X[:, None, :] - self.mu[None, :, :]) ** 2).sum(-1)
                else:
                    d2s = d2
                self.sigma = float(((g * d2s).sum() / (self.D * N)).sqrt())
                if self.ensure_continuum:
                    self.sigma = max(self.sigma, intrinsic_scale(self.mu))

        Y = g @ self.mu
        lpi = self.w - torch.logsumexp(self.w, 0)
        lgn = self._lgn(self.sigma) if variant == "keops" else lgn_old
        mu2 = (self.mu ** 2).sum(-1)
        y2 = (Y ** 2).sum(-1)
        cfe_n = (g * ((mu2[None, :] - y2[:, None]) / (2 * self.sigma ** 2) + lg - lpi[None, :])).sum(1) + lgn
        sq = ((X - Y) ** 2).sum(-1)
        if self.outliers is None:
            Cfe = cfe_n.sum()
            FE = Cfe + float(sq.sum()) / (2 * self.sigma ** 2)
        else:
            g0, gT = lg0.exp(), lgT.exp()
            lpi0, lpiT = log_ratio_to_proba(self.outliers["eta0"])
            lpi0, lpiT = float(lpi0), float(lpiT)
            Cfe = float((gT * (cfe_n + lgT - lpiT) + g0 * (-logJ0 + lg0 - lpi0)).sum())
            FE = Cfe + float((gT * sq).sum()) / (2 * self.sigma ** 2)
        return Y, Cfe, FE

    def em_optimization(self, X, max_iterations=100, tol=1e-5, variant="keops"):
        if X.shape[0] == 0:
            return torch.empty_like(X), torch.tensor(0.0), torch.tensor(0.0), 0
        last = None
        for i in range(max_iterations):
            Y, Cfe, FE = self.em_step(X, variant=variant)
            if last is not None and tol is not None and abs(FE - last) < tol * abs(last):
                return Y, Cfe, FE, i + 1
            last = FE
        return Y, Cfe, FE, i + 1
