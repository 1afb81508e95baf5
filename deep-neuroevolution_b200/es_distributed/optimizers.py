"""``es_distributed.optimizers`` with the reference surface (optimizers.py:4-50): ``SGD`` / ``Adam`` objects whose
``update(globalg)`` returns ``(ratio, theta)`` -- but theta, m and v live in HBM and the step is one fused
kernel (dne_adam_step / dne_sgd_step).

The reference's master calls ``optimizer.update(-g + l2coeff*theta)`` (es.py:298).  The fused kernel forms that
direction itself, so the engine path calls ``update_from_gradient(g, l2coeff)``; ``update(globalg)`` is kept for
callers that already hold the full direction (it is then run with l2coeff = 0 on ``-globalg``).
"""
from __future__ import annotations

import numpy as np
import torch

from dne.engine import ESUpdate


class Optimizer(object):
    kind = None

    def __init__(self, theta, ctx=None, **args):
        if ctx is None:
            from .es import default_context
            ctx = default_context()
        self._upd = ESUpdate(ctx, theta, self.kind, **args)
        self.dim = self._upd.P

    @property
    def t(self):
        return self._upd.t

    @property
    def theta(self) -> np.ndarray:
        return self._upd.theta.cpu().numpy()

    @property
    def device_theta(self) -> torch.Tensor:
        return self._upd.theta

    def update_from_gradient(self, g: torch.Tensor, l2coeff: float):
        """es.py:298 fused: theta <- theta + step(-g + l2coeff*theta).  Returns (device ratio scalar, device theta)."""
        ratio = self._upd.step(l2coeff, g)
        return ratio, self._upd.theta

    def update(self, globalg):
        """optimizers.py:10-17."""
        g = torch.as_tensor(np.asarray(globalg, dtype=np.float32)).to(self._upd.device) \
            if not isinstance(globalg, torch.Tensor) else globalg
        ratio = self._upd.step(0.0, (-g).contiguous())
        return float(ratio.cpu()), self.theta


class SGD(Optimizer):
    kind = "sgd"

    def __init__(self, theta, stepsize, momentum=0.9, ctx=None):
        Optimizer.__init__(self, theta, ctx=ctx, stepsize=stepsize, momentum=momentum)
        self.stepsize, self.momentum = stepsize, momentum


class Adam(Optimizer):
    kind = "adam"

    def __init__(self, theta, stepsize, beta1=0.9, beta2=0.999, epsilon=1e-08, ctx=None):
        Optimizer.__init__(self, theta, ctx=ctx, stepsize=stepsize, beta1=beta1, beta2=beta2, epsilon=epsilon)
        self.stepsize, self.beta1, self.beta2, self.epsilon = stepsize, beta1, beta2, epsilon
