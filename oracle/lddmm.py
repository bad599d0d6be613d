"""Oracle: LDDMM Hamiltonian system, integrators, shooting (TEST INFRASTRUCTURE).

Restates on the CPU (torch, autograd-enabled, dtype-agnostic)

  LDDMMModel.v / mdivsum / Hamiltonian / dtrajcost / ODE   /root/reference/diffICP/core/LDDMM.py:100-227
  LDDMMModel.Shoot / trajloss                              core/LDDMM.py:286-334
  EulerIntegrator / RalstonIntegrator                      /root/reference/diffICP/tools/integrators.py:20-51

Gradients are obtained by torch autograd through the unrolled loop, exactly as
the reference does (tools/optim.py:34-47: ``L.backward()`` in the LBFGS closure),
so this file is also the oracle for the hand-written discrete adjoint.

Model variants (core/LDDMM.py:43-56):
  classic : gradcomponent=False, withlogdet=False   (eta = 0, dcost = 0)
  hybrid  : gradcomponent=False, withlogdet=True    (eta = 0, dcost = -sum div v)
  logdet  : gradcomponent=True,  withlogdet=True    (eta = 1/lambda)
"""

from __future__ import annotations

import torch

from .kernels import GaussOracle


def euler(f, state, nt, T=1.0):
    """Explicit Euler on a tuple state; returns the list of nt+1 states (integrators.py:20-31)."""
    h = T / nt
    cur = tuple(s.clone() for s in state)
    traj = [cur]
    for _ in range(nt):
        k = f(*cur)
        cur = tuple(s + h * ds for s, ds in zip(cur, k))
        traj.append(cur)
    return traj


def ralston(f, state, nt, T=1.0):
    """Ralston RK2 (2/3 midpoint, weights 1/4 and 3/4); integrators.py:36-51."""
    h = T / nt
    cur = tuple(s.clone() for s in state)
    traj = [cur]
    for _ in range(nt):
        k1 = f(*cur)
        mid = tuple(s + (2.0 * h / 3.0) * ds for s, ds in zip(cur, k1))
        k2 = f(*mid)
        cur = tuple(s + (0.25 * h) * (a + 3.0 * b) for s, a, b in zip(cur, k1, k2))
        traj.append(cur)
    return traj


class LDDMMOracle:
    def __init__(self, sigma=1.0, D=2, lambd=2.0, version="logdet", scheme="Ralston", nt=10,
                 chunk=2048, try_trajcost_optim=False):
        flags = {"classic": (False, False), "hybrid": (False, True), "logdet": (True, True)}
        self.gradcomponent, self.withlogdet = flags[version]
        self.K = GaussOracle(sigma, D, chunk)
        self.D, self.lam, self.nt = D, float(lambd), int(nt)
        self.eta = 1.0 / self.lam if self.gradcomponent else 0.0
        self.scheme = scheme
        self.integrator = {"Euler": euler, "Ralston": ralston}[scheme]
        self.try_trajcost_optim = try_trajcost_optim

    # v(x) = sum_j K(x-q_j) p_j - eta * sum_j gradK(x-q_j)            core/LDDMM.py:100-116
    def v(self, x, q, p):
        if x.numel() == 0:
            return torch.empty_like(x)
        out = self.K.KRed(x, q, p)
        if self.gradcomponent:
            out = out - self.eta * self.K.GradKRed(x, q)
        return out

    # -sum_k div v(x_k)                                               core/LDDMM.py:120-138
    def mdivsum(self, x, q, p):
        if x.numel() == 0:
            return torch.zeros(1, dtype=q.dtype)
        out = (p * self.K.GradKRed(q, x)).sum()
        if self.gradcomponent:
            out = out + self.eta * self.K.LapKRed(q, x).sum()
        return out

    # H(q,p)                                                          core/LDDMM.py:142-159
    def hamiltonian(self, q, p):
        H = 0.5 * (p * self.K.KRed(q, q, p)).sum()
        if self.gradcomponent:
            H = H - self.eta * (p * self.K.GradKRed(q, q)).sum() \
                - 0.5 * self.eta ** 2 * self.K.LapKRed(q, q).sum()
        return H

    # lambda*H + mdivsum(q,q,p) shortcut                              core/LDDMM.py:163-172
    def dtrajcost(self, q, p):
        return 0.5 * self.lam * (p * self.K.KRed(q, q, p)).sum() + 0.5 * self.eta * self.K.LapKRed(q, q).sum()

    # right-hand side                                                 core/LDDMM.py:176-227
    def ode(self, q, p, cost, x=None):
        vq = self.v(q, q, p)
        Gq = self.K.GenDKRed(q, q, p, p)
        if self.eta != 0:
            Gq = Gq - self.eta * self.K.HessKRed(q, q, p, p) - self.eta ** 2 * self.K.GradLapKRed(q, q)
        zero = torch.zeros(1, dtype=q.dtype)
        if x is None:
            if self.withlogdet:
                if self.gradcomponent and self.try_trajcost_optim:
                    dcost = self.dtrajcost(q, p)
                else:
                    dcost = self.mdivsum(q, q, p)
            else:
                dcost = zero
            return vq, -Gq, dcost
        dcost = self.mdivsum(x, q, p) if self.withlogdet else zero
        return vq, -Gq, dcost, self.v(x, q, p)

    # shooting                                                        core/LDDMM.py:286-299
    def shoot(self, q0, p0, x0=None):
        cost0 = torch.zeros(1, dtype=q0.dtype)
        st = (q0, p0, cost0) if x0 is None else (q0, p0, cost0, x0)
        return self.integrator(self.ode, st, self.nt)

    # trajectory energy                                               core/LDDMM.py:318-334
    def trajloss(self, shoot):
        end = shoot[-1]
        if len(end) == 3 and self.withlogdet and self.gradcomponent and self.try_trajcost_optim:
            return end[2]
        q0, p0 = shoot[0][:2]
        return self.lam * self.hamiltonian(q0, p0) + end[2]

    # total registration loss for given quadratic targets            core/LDDMM.py:363-371 + core/PSR.py:498-516
    def loss(self, q0, p0, x0, y, inv2sig2):
        """inv2sig2: scalar or (n,) tensor of 1/(2 sigma_s^2) per data point."""
        sh = self.shoot(q0, p0, x0)
        moved = sh[-1][0] if x0 is None else sh[-1][3]
        w = inv2sig2 if not torch.is_tensor(inv2sig2) or inv2sig2.dim() == 0 else inv2sig2[:, None]
        return self.trajloss(sh) + (((moved - y) ** 2) * w).sum(), sh
