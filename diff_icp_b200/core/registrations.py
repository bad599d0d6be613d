"""Registration objects: apply a fitted deformation to external points (reference: core/registrations.py:21-87)."""

import warnings

import torch

from .LDDMM import LDDMMModel


class Registration:
    """Informal common interface (apply / backward / shoot)."""

    def apply(self, X: torch.Tensor):
        raise NotImplementedError

    def backward(self, Y: torch.Tensor):
        raise NotImplementedError

    def shoot(self, X: torch.Tensor, backward=False):
        raise NotImplementedError


class LDDMMRegistration(Registration):

    def __init__(self, LMi: LDDMMModel, q0: torch.Tensor, a0: torch.Tensor):
        self.LMi = LMi
        self.q0 = q0
        self.a0 = a0

    def shoot(self, X: torch.Tensor, backward=False, previous_forwardshoot=None):
        """Shoot the external points X along the geodesic (q0,a0); with backward=True along the inverse flow, obtained
        by re-shooting from the arrival state with reversed momenta (q1, -a1) (reference: core/registrations.py:56-70;
        an exact inverse only when eta = 0)."""
        if not backward:
            if previous_forwardshoot is not None:
                warnings.warn("variable 'previous_forwardshoot' is useless when backward=False [default]", RuntimeWarning)
            return self.LMi.Shoot(self.q0, self.a0, X)
        if previous_forwardshoot is None:
            previous_forwardshoot = self.shoot(None)
        q1, a1 = previous_forwardshoot[-1][0], previous_forwardshoot[-1][1]
        return self.LMi.Shoot(q1, -a1, X)

    def apply(self, X: torch.Tensor):
        return self.shoot(X)[-1][3]

    def backward(self, Y: torch.Tensor, previous_forwardshoot=None):
        return self.shoot(Y, backward=True, previous_forwardshoot=previous_forwardshoot)[-1][3]
