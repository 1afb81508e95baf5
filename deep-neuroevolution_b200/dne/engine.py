"""Host-side engine objects over the C ABI: batched perturb+forward over env slots, and the generation update.

PyTorch tensors are the device-memory container; all arithmetic on the named path happens in libdne.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _ffi as F
from .nets import NetSpec
from .noise import SharedNoiseTable


def _dev_bytes(n: int, device) -> torch.Tensor:
    return torch.empty(max(int(n), 256), dtype=torch.uint8, device=device)


class SlotForward:
    """perturb + forward + action-select for ``n_slots`` environment slots (one member per slot).

    Slot s runs weights ``theta[theta_idx[s]] + scale[s] * noise[noise_idx[s] : +P]``:
      ES   (es.py:412-419)  scale = +sigma / -sigma on slots (2p, 2p+1) sharing one index (``paired=True``)
      eval (es.py:388-391)  scale = 0
      GA   (ga.py:256-264; models/base.py:148-156)  theta_idx selects the cached parent, scale = mutation power
    """

    def __init__(self, ctx: F.Context, net: NetSpec, n_slots: int, n_ref: int = 128):
        self.ctx, self.net, self.n_slots, self.n_ref = ctx, net, int(n_slots), int(n_ref)
        dev = torch.device("cuda", ctx.device)
        self.device = dev
        self.noise_idx = torch.zeros(n_slots, dtype=torch.int64, device=dev)
        self.scale = torch.zeros(n_slots, dtype=torch.float32, device=dev)
        self.theta_idx: Optional[torch.Tensor] = None
        self.active: Optional[torch.Tensor] = None
        self.actions = torch.zeros(n_slots, dtype=torch.int32, device=dev)
        self.logits = torch.zeros(n_slots, net.n_out, dtype=torch.float32, device=dev)
        nb = C.c_size_t()
        F.check(F.lib().dne_forward_ws_bytes(C.byref(net.desc), n_slots, C.byref(nb)))
        self.ws = _dev_bytes(nb.value, dev)
        F.check(F.lib().dne_theta_forget(ctx.handle, F.ptr(self.ws)))     # a recycled address must not inherit a prepared entry
        self.vbn = None
        self.vbn_ws = None
        if net.vbn_len:
            self.vbn = torch.zeros(n_slots, net.vbn_len, dtype=torch.float32, device=dev)

    # -- slot table -------------------------------------------------------------------------------------
    def set_slots(self, noise_idx, scale, active=None, theta_idx=None):
        """Upload the slot table (host numpy arrays or device tensors)."""
        self.noise_idx.copy_(torch.as_tensor(noise_idx, dtype=torch.int64), non_blocking=True)
        self.scale.copy_(torch.as_tensor(scale, dtype=torch.float32), non_blocking=True)
        if active is None:
            self.active = None
        else:
            if self.active is None:
                self.active = torch.zeros(self.n_slots, dtype=torch.uint8, device=self.device)
            self.active.copy_(torch.as_tensor(active, dtype=torch.uint8), non_blocking=True)
        if theta_idx is None:
            self.theta_idx = None
        else:
            if self.theta_idx is None:
                self.theta_idx = torch.zeros(self.n_slots, dtype=torch.int32, device=self.device)
            self.theta_idx.copy_(torch.as_tensor(theta_idx, dtype=torch.int32), non_blocking=True)

    # -- per tick -----------------------------------------------------------------------------------------
    def forward(self, theta: torch.Tensor, obs: torch.Tensor, *, paired: bool, ob_mean=None, ob_std=None,
                n_slots: Optional[int] = None) -> torch.Tensor:
        """One env tick for all slots.  Returns the device tensor of actions (int32 [n_slots] for conv policies,
        float32 [n_slots, n_out] for the MLP).  Asynchronous on the current stream."""
        n = self.n_slots if n_slots is None else int(n_slots)
        net, L = self.net, F.lib()
        assert theta.dim() in (1, 2) and theta.shape[-1] == net.num_params
        if net.ob_kind == F.OB_ATARI_U8:
            if self.theta_idx is None:
                self.prepare(theta, n)
            F.check(L.dne_perturb_forward_conv(
                self.ctx.handle, C.byref(net.desc), F.ptr(theta, torch.float32), F.ptr(self.noise_idx),
                F.ptr(self.scale), F.ptr(self.theta_idx), F.ptr(self.active), n, int(paired),
                F.ptr(obs, torch.uint8), F.ptr(self.vbn), F.ptr(self.actions), F.ptr(self.logits),
                F.ptr(self.ws), self.ws.numel(), F.stream_ptr()))
            return self.actions
        F.check(L.dne_perturb_forward_mlp(
            self.ctx.handle, C.byref(net.desc), F.ptr(theta, torch.float32), F.ptr(self.noise_idx),
            F.ptr(self.scale), F.ptr(self.theta_idx), F.ptr(self.active), n, int(paired),
            F.ptr(obs, torch.float32), F.ptr(ob_mean), F.ptr(ob_std), F.ptr(self.logits),
            F.ptr(self.ws), self.ws.numel(), F.stream_ptr()))
        return self.logits

    def prepare(self, theta: torch.Tensor, n_slots: Optional[int] = None):
        """Once per theta: dne_theta_prepare (tensor-core operand layout of the fc weights in this table's workspace).
        Keyed by (pointer, torch in-place version, engine epoch, slot count); ESUpdate.step and the other engine methods
        that rewrite theta through the C ABI bump the epoch of the context."""
        n = self.n_slots if n_slots is None else int(n_slots)
        key = (theta.data_ptr(), theta._version, getattr(self.ctx, "theta_epoch", 0), n)
        if key == getattr(self, "_prep_key", None):
            return
        F.check(F.lib().dne_theta_prepare(self.ctx.handle, C.byref(self.net.desc), F.ptr(theta, torch.float32), n,
                                          F.ptr(self.ws), self.ws.numel(), F.stream_ptr()))
        self._prep_key = key

    # -- per episode (ESAtariPolicy) -------------------------------------------------------------------------
    def vbn_reference_pass(self, theta: torch.Tensor, ref_batch: torch.Tensor, active: Optional[torch.Tensor] = None):
        """policies.py:399 -- refresh the virtual-batch-norm statistics of the slots flagged in ``active``
        (all if None) from the shared reference batch [n_ref,84,84,4] uint8."""
        net, L = self.net, F.lib()
        assert net.vbn_len, "net has no batch norm"
        n_ref = int(ref_batch.shape[0])
        if self.vbn_ws is None or self._vbn_ws_ref != n_ref:
            nb = C.c_size_t()
            F.check(L.dne_vbn_ws_bytes(C.byref(net.desc), self.n_slots, n_ref, C.byref(nb)))
            self.vbn_ws = _dev_bytes(nb.value, self.device)
            self._vbn_ws_ref = n_ref
        F.check(L.dne_vbn_reference_pass(
            self.ctx.handle, C.byref(net.desc), F.ptr(theta, torch.float32), F.ptr(self.noise_idx),
            F.ptr(self.scale), F.ptr(self.theta_idx), F.ptr(active), self.n_slots, F.ptr(ref_batch, torch.uint8),
            n_ref, F.ptr(self.vbn), F.ptr(self.vbn_ws), self.vbn_ws.numel(), F.stream_ptr()))


class ESUpdate:
    """The master's update block (es.py:273-301) on the device: ranks -> gradient -> optimizer step.
    theta, Adam m/v never leave HBM."""

    def __init__(self, ctx: F.Context, theta0, optimizer: str = "adam", **opt_args):
        dev = torch.device("cuda", ctx.device)
        self.ctx, self.device = ctx, dev
        self.theta = torch.as_tensor(np.asarray(theta0, dtype=np.float32)).to(dev).contiguous() \
            if not isinstance(theta0, torch.Tensor) else theta0.to(dev, torch.float32).contiguous().clone()
        self.P = int(self.theta.numel())
        self.kind = optimizer
        self.args = dict(opt_args)
        self.t = 0
        self.m = torch.zeros_like(self.theta) if optimizer == "adam" else None
        self.v = torch.zeros_like(self.theta)
        self.g = torch.zeros_like(self.theta)
        self.ratio = torch.zeros(1, dtype=torch.float32, device=dev)

    def centered_ranks(self, returns_n2: torch.Tensor):
        """compute_centered_ranks (es.py:81-85).  Returns (centered f32 [n,2], ranks int32 [2n])."""
        x = returns_n2.to(self.device, torch.float32).contiguous()
        out = torch.empty_like(x)
        ranks = torch.empty(x.numel(), dtype=torch.int32, device=self.device)
        F.check(F.lib().dne_centered_rank(F.ptr(x), x.numel(), F.ptr(out), F.ptr(ranks), F.stream_ptr()))
        return out, ranks

    def gradient(self, proc_n2: torch.Tensor, noise_idx: torch.Tensor, denom: float, accumulate: bool = False):
        """batched_weighted_sum / size (es.py:291-296) over this rank's noise indices; ``denom`` is the GLOBAL
        returns_n2.size so per-rank partial gradients simply add up."""
        n = int(noise_idx.numel())
        assert proc_n2.shape == (n, 2)
        proc_n2, noise_idx = proc_n2.contiguous(), noise_idx.contiguous()   # keep references: F.ptr() borrows
        F.check(F.lib().dne_es_grad(self.ctx.handle, F.ptr(proc_n2, torch.float32),
                                    F.ptr(noise_idx, torch.int64), n, self.P, float(denom),
                                    F.ptr(self.g), int(accumulate), F.stream_ptr()))
        return self.g

    def step(self, l2coeff: float, g: Optional[torch.Tensor] = None) -> torch.Tensor:
        """optimizer.update(-g + l2coeff*theta) (es.py:298).  Returns the device scalar update ratio."""
        g = self.g if g is None else g
        self.t += 1
        self.ctx.theta_epoch = getattr(self.ctx, "theta_epoch", 0) + 1        # theta is rewritten through the C ABI
        L = F.lib()
        if self.kind == "adam":
            a = self.args
            F.check(L.dne_adam_step(self.ctx.handle, F.ptr(self.theta), F.ptr(self.m), F.ptr(self.v), F.ptr(g), self.P,
                                    float(l2coeff), float(a["stepsize"]), float(a.get("beta1", 0.9)),
                                    float(a.get("beta2", 0.999)), float(a.get("epsilon", 1e-8)), self.t,
                                    F.ptr(self.ratio), F.stream_ptr()))
        elif self.kind == "sgd":
            a = self.args
            F.check(L.dne_sgd_step(self.ctx.handle, F.ptr(self.theta), F.ptr(self.v), F.ptr(g), self.P, float(l2coeff),
                                   float(a["stepsize"]), float(a.get("momentum", 0.9)), F.ptr(self.ratio),
                                   F.stream_ptr()))
        else:
            raise NotImplementedError(self.kind)
        return self.ratio


def make_context(device: int, noise: SharedNoiseTable) -> F.Context:
    ctx = F.Context(device)
    ctx.bind_noise(noise.device_tensor, noise.count)
    return ctx
