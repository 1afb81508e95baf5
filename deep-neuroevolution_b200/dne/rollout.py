"""Episode scheduler: runs a list of rollout units (antithetic pairs / GA offspring / eval episodes) to completion
on a fixed table of environment slots, refilling slots as episodes end, with the host environment step of one
half of the slots overlapped with the device forward of the other half.

Reference behaviour reproduced:
  * worker inner loop            es_distributed/es.py:411-426  (theta+v rollout, theta-v rollout, sums / signs / lengths)
  * Policy.rollout               es_distributed/policies.py:378-429 (reset -> ref-batch pass -> act/step until done
                                 or timestep_limit; length counts env steps)
  * slot refill                  gpu_implementation/neuroevolution/concurrent_worker.py:72-125 (running mask, per-slot
                                 cumulative reward/length, finished slots returned to the pool)
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _ffi as F
from .engine import SlotForward
from .envs import BatchEnv
from .nets import NetSpec


@dataclass
class Unit:
    """G episodes sharing one noise index (G = 2: the +/- pair of es.py:412-421; G = 1: one GA offspring)."""
    noise_idx: int
    scales: Sequence[float]
    theta_idx: int = 0
    noiseless: bool = False    # evaluation episodes (es.py:388-391): no action noise, never sampled for ob statistics


@dataclass
class RolloutResult:
    returns: np.ndarray        # float32 [n_units, G]   es.py:425
    signreturns: np.ndarray    # float32 [n_units, G]   es.py:423
    lengths: np.ndarray        # int32   [n_units, G]   es.py:426
    bcs: Optional[list] = None  # per unit, per member: behaviour characterisation (policies.py:418,429)
    steps: int = 0             # env steps executed (== lengths.sum())
    ticks: int = 0             # forward launches
    ob_sum: Optional[np.ndarray] = None     # float64 [ob_dim]: sum of the observations of the sampled episodes (es.py:358-359)
    ob_sumsq: Optional[np.ndarray] = None
    ob_count: int = 0                       # number of observations in the sums


class _Half:
    def __init__(self, ctx, net, lo, hi, n_ref):
        self.lo, self.hi = lo, hi
        n = hi - lo
        self.sf = SlotForward(ctx, net, n, n_ref=n_ref)
        dev = self.sf.device
        self.stream = torch.cuda.Stream(device=dev)
        self.event = torch.cuda.Event()
        if net.ob_kind == F.OB_ATARI_U8:
            self.obs_dev = torch.zeros(n, 84, 84, 4, dtype=torch.uint8, device=dev)
            self.act_host = torch.zeros(n, dtype=torch.int32).pin_memory()
        else:
            self.obs_dev = torch.zeros(n, net.ob_dim, dtype=torch.float32, device=dev)
            self.act_host = torch.zeros(n, net.n_out, dtype=torch.float32).pin_memory()
        self.noise_idx = np.zeros(n, dtype=np.int64)
        self.scale = np.zeros(n, dtype=np.float32)
        self.theta_idx = np.zeros(n, dtype=np.int32)
        self.active = np.zeros(n, dtype=np.uint8)
        self.unit = np.full(n, -1, dtype=np.int64)       # unit id occupying the slot
        self.member = np.zeros(n, dtype=np.int64)
        self.ret = np.zeros(n, dtype=np.float64)
        self.sret = np.zeros(n, dtype=np.float64)
        self.length = np.zeros(n, dtype=np.int64)
        self.dirty = True
        self.launched = False
        self.fresh = np.zeros(n, dtype=np.uint8)          # slots that start an episode at the next launch
        self.noiseless = np.zeros(n, dtype=bool)          # evaluation episodes: no action noise
        self.save = np.zeros(n, dtype=np.uint8)           # episode sampled for the observation statistics (es.py:356-357)
        self.save_m = 0                                   # number of slots in save_list
        self.save_list = None                             # device int32 [n]: active & sampled slots (built lazily)
        self.ob_sum = self.ob_sumsq = None                # device float64 [ob_dim] (per half: the halves run on two streams)


class RolloutRunner:
    def __init__(self, ctx: F.Context, net: NetSpec, env: BatchEnv, n_slots: int, group: int = 2, pipeline: int = 2,
                 ref_batch: Optional[torch.Tensor] = None):
        assert n_slots % (group * pipeline) == 0, "n_slots must be a multiple of group*pipeline"
        assert env.n_slots >= n_slots
        self.ctx, self.net, self.env, self.n_slots, self.G = ctx, net, env, n_slots, group
        per = n_slots // pipeline
        n_ref = int(ref_batch.shape[0]) if ref_batch is not None else 128
        self.halves = [_Half(ctx, net, i * per, (i + 1) * per, n_ref) for i in range(pipeline)]
        self.ref_batch = ref_batch
        self.use_theta_idx = False
        self.action_fn = None          # optional host map from the network's output rows to environment actions

    # ---------------------------------------------------------------------------------------------------
    def run(self, theta: torch.Tensor, units: List[Unit], timestep_limit: Optional[int] = None, *, ob_mean=None,
            ob_std=None, collect_bc: Optional[str] = None, ac_noise_std: float = 0.0,
            random_stream: Optional[np.random.RandomState] = None, save_obs_prob: float = 0.0) -> RolloutResult:
        """Evaluate every unit once.  ``collect_bc``: None | 'trace' (RAM after every step, ES Atari,
        policies.py:410,418) | 'final' (RAM / position at episode end, policies.py:510,292-299)."""
        G, env = self.G, self.env
        n_units = len(units)
        limit = env.max_episode_steps if timestep_limit is None else \
            (timestep_limit if env.max_episode_steps is None else min(timestep_limit, env.max_episode_steps))
        assert limit is not None and limit >= 1
        res = RolloutResult(np.zeros((n_units, G), np.float32), np.zeros((n_units, G), np.float32),
                            np.zeros((n_units, G), np.int32), [[None] * G for _ in range(n_units)] if collect_bc else None)
        self.use_theta_idx = theta.dim() == 2 and theta.shape[0] > 1
        pending = deque(range(n_units))
        remaining = [G] * n_units
        want_obstat = save_obs_prob != 0.0 and self.net.ob_kind == F.OB_VECTOR
        obstat_stream = random_stream if random_stream is not None else np.random.RandomState(0)
        for h in self.halves:
            h.save[:] = 0
            h.save_m = 0
            if want_obstat:
                if h.ob_sum is None:
                    h.ob_sum = torch.zeros(self.net.ob_dim, dtype=torch.float64, device=h.sf.device)
                    h.ob_sumsq = torch.zeros_like(h.ob_sum)
                    h.save_list = torch.zeros(h.hi - h.lo, dtype=torch.int32, device=h.sf.device)
                    h.save_host = torch.zeros(h.hi - h.lo, dtype=torch.int32).pin_memory()
                h.ob_sum.zero_()
                h.ob_sumsq.zero_()
        if collect_bc == "trace":                          # per-slot RAM trace buffers, filled with vectorised writes
            for h in self.halves:
                if getattr(h, "bc_buf", None) is None or h.bc_buf.shape[1] < limit:
                    h.bc_buf = np.zeros((h.hi - h.lo, limit, 128), dtype=np.uint8)
        cur = torch.cuda.current_stream()
        for h in self.halves:
            h.unit[:] = -1
            h.active[:] = 0
            h.launched = False
            h.dirty = True
            h.stream.wait_stream(cur)

        def refill(h: _Half):
            n = h.hi - h.lo
            for u0 in range(0, n, G):
                if not pending:
                    break
                if h.unit[u0] >= 0:
                    continue
                uid = pending.popleft()
                unit = units[uid]
                for g in range(G):
                    s = u0 + g
                    h.unit[s], h.member[s] = uid, g
                    h.noise_idx[s], h.scale[s], h.theta_idx[s] = unit.noise_idx, unit.scales[g], unit.theta_idx
                    h.active[s], h.fresh[s] = 1, 1
                    h.ret[s] = h.sret[s] = 0.0
                    h.length[s] = 0
                    h.noiseless[s] = unit.noiseless
                    # es.py:356-357: each (non-evaluation) episode is sampled with probability calc_obstat_prob
                    h.save[s] = 1 if (want_obstat and not unit.noiseless and obstat_stream.rand() < save_obs_prob) else 0
                env.reset(h.lo + np.arange(u0, u0 + G))
                h.dirty = True

        def launch(h: _Half):
            # torch.cuda.set_stream pair instead of the `with torch.cuda.stream(...)` context manager: the manager's device-index
            # validation costs ~28 us per half-tick (cProfile), a quarter of the host time of a 64-slot half-table
            torch.cuda.set_stream(h.stream)
            try:
                if h.dirty:
                    h.sf.set_slots(h.noise_idx, h.scale, active=h.active,
                                   theta_idx=h.theta_idx if self.use_theta_idx else None)
                    if self.net.needs_ref_batch and h.fresh.any():
                        mask = torch.as_tensor(h.fresh).to(h.sf.device, non_blocking=True)
                        h.sf.vbn_reference_pass(theta, self.ref_batch, active=mask)     # policies.py:399
                    h.fresh[:] = 0
                    if want_obstat:                                  # slot list of the sampled episodes still running
                        loc = np.nonzero(np.logical_and(h.active, h.save))[0].astype(np.int32)
                        h.save_m = len(loc)
                        if h.save_m:
                            h.save_host[:h.save_m] = torch.from_numpy(loc)
                            h.save_list[:h.save_m].copy_(h.save_host[:h.save_m], non_blocking=True)
                    h.dirty = False
                if hasattr(env, "device_obs"):                                           # raw frames -> device preprocess (dne/raw_env.py)
                    h.obs_dev = env.device_obs(h.lo, h.hi)
                else:
                    h.obs_dev.copy_(env.obs_block(h.lo, h.hi), non_blocking=True)       # pinned -> HBM
                if want_obstat and h.save_m:                         # es.py:358-359 on the device, unnormalised observations
                    F.check(F.lib().dne_ob_stat_accumulate(F.ptr(h.obs_dev, torch.float32), self.net.ob_dim,
                                                           F.ptr(h.save_list), h.save_m, F.ptr(h.ob_sum),
                                                           F.ptr(h.ob_sumsq), F.stream_ptr()))
                    res.ob_count += h.save_m
                out = h.sf.forward(theta, h.obs_dev, paired=(G == 2), ob_mean=ob_mean, ob_std=ob_std)
                h.act_host.copy_(out, non_blocking=True)
                h.event.record(h.stream)
            finally:
                torch.cuda.set_stream(cur)
            h.launched = True
            res.ticks += 1

        def finish(h: _Half):
            h.event.synchronize()
            h.launched = False
            loc = np.nonzero(h.active)[0]
            acts = h.act_host.numpy()[loc]
            if self.action_fn is not None:                             # discretised MuJoCo heads: scores -> bin values
                acts = self.action_fn(acts)
            if ac_noise_std != 0.0 and random_stream is not None and acts.dtype != np.int32:
                noisy = ~h.noiseless[loc]                             # evaluation episodes act without noise (es.py:388-391)
                acts = acts + (random_stream.randn(*acts.shape).astype(np.float32) * np.float32(ac_noise_std)) * \
                    noisy[:, None].astype(np.float32)                 # policies.py:204-205
            rew, done = env.step(h.lo + loc, acts)
            h.ret[loc] += rew
            h.sret[loc] += np.sign(rew)
            h.length[loc] += 1
            res.steps += len(loc)
            if collect_bc == "trace":
                h.bc_buf[loc, h.length[loc] - 1] = env.get_ram(h.lo + loc)        # policies.py:410,418
            fin = loc[np.logical_or(done, h.length[loc] >= limit)]
            if len(fin):
                for s in fin:
                    uid, g = int(h.unit[s]), int(h.member[s])
                    res.returns[uid, g] = np.float32(h.ret[s])
                    res.signreturns[uid, g] = np.float32(h.sret[s])
                    res.lengths[uid, g] = h.length[s]
                    if collect_bc == "trace":
                        res.bcs[uid][g] = h.bc_buf[s, :h.length[s]].copy()
                    elif collect_bc == "final":
                        res.bcs[uid][g] = env.get_ram(np.array([h.lo + s]))[0]
                    h.active[s] = 0
                    remaining[uid] -= 1
                    if remaining[uid] == 0:
                        base = (s // G) * G
                        h.unit[base:base + G] = -1
                h.dirty = True

        for h in self.halves:
            refill(h)
            if h.active.any():
                launch(h)
        while any(h.launched for h in self.halves):
            for h in self.halves:
                if not h.launched:
                    continue
                finish(h)
                refill(h)
                if h.active.any():
                    launch(h)
            if hasattr(env, "advance"):
                env.advance()
        for h in self.halves:
            cur.wait_stream(h.stream)
        assert not pending and all(r == 0 for r in remaining)
        if want_obstat:
            res.ob_sum = sum(h.ob_sum for h in self.halves).cpu().numpy()
            res.ob_sumsq = sum(h.ob_sumsq for h in self.halves).cpu().numpy()
        return res
