"""Batched environment interface.  Emulator stepping stays on the HOST (north_star: "ALE env-step stays on the
host behind pinned cudaMemcpyAsync"); the device only ever sees uint8 frame stacks (or float vectors) and
returns actions.

Reference counterparts: gym env + ``wrap_deepmind`` (es_distributed/atari_wrappers.py:204-222) stepped one at a
time by each worker (policies.py:398-409); on the reference GPU path the TF ops ``EnvironmentReset/Observation/
Step`` over a batch of ALE instances (gpu_implementation/gym_tensorflow/tf_env.cpp:115-316).

ALE / gym / MuJoCo are not vendored by the reference and are absent from this image, so the environment shipped
here is the synthetic Frostbite-shaped stub the measurement plan names: i.i.d. uint8 84x84x4
observations from a fixed pool, rewards 10*Bernoulli(0.05), fixed or ragged episode lengths.  A real emulator
plugs in by subclassing ``BatchEnv``.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape if shape is not None else np.shape(low)).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.low.shape).copy()
        self.shape = self.low.shape


class BatchEnv:
    """``n_slots`` independent environments.  Observations live in a pinned host tensor ``obs`` [n_slots, ...]
    so the engine can cudaMemcpyAsync them without staging."""
    observation_space = None
    action_space = None
    n_slots = 0
    max_episode_steps: Optional[int] = None     # env.spec...max_episode_steps of the reference (policies.py:383)

    def reset(self, slots: np.ndarray) -> None:
        raise NotImplementedError

    def step(self, slots: np.ndarray, actions: np.ndarray):
        """Step the listed slots.  Returns (rewards float32 [k], done bool [k]); new observations are written
        into ``self.obs[slots]``."""
        raise NotImplementedError

    def obs_block(self, lo: int, hi: int) -> torch.Tensor:
        """Pinned host view of the observations of slots [lo, hi) for the next forward."""
        return self.obs[lo:hi]

    def get_ram(self, slots: np.ndarray) -> np.ndarray:
        """Behaviour characterisation source (policies.py:410): uint8 [k,128]."""
        raise NotImplementedError

    def random_actions(self, k: int, rs: np.random.RandomState) -> np.ndarray:
        return rs.randint(0, self.action_space.n, size=k)


class SyntheticAtariEnv(BatchEnv):
    """Frostbite-shaped stub: 18 actions, 84x84x4 uint8 observations drawn i.i.d. uniform
    from ``torch.Generator(seed)`` (a pool of frames, rotated every tick), reward 10*Bernoulli(0.05), episode
    length fixed (``episode_len``) or ragged ``U{lo..hi}`` per episode (seeded)."""

    def __init__(self, n_slots: int, num_actions: int = 18, episode_len=1000, seed: int = 0, pool_blocks: int = 4,
                 pin: bool = True):
        self.n_slots = int(n_slots)
        self.observation_space = Box(0, 255, (84, 84, 4), dtype=np.uint8)
        self.action_space = Discrete(num_actions)
        g = torch.Generator().manual_seed(seed)
        self.pool_blocks = int(pool_blocks)
        pool = torch.randint(0, 256, (self.pool_blocks, self.n_slots, 84, 84, 4), dtype=torch.uint8, generator=g)
        self.pool = pool.pin_memory() if (pin and torch.cuda.is_available()) else pool
        self.obs = self.pool[0]
        self._tick = 0
        self.rs = np.random.RandomState(seed)
        self.episode_len_spec = episode_len
        self.max_episode_steps = episode_len if isinstance(episode_len, int) else int(episode_len[1])
        self.ep_len = np.zeros(self.n_slots, dtype=np.int64)
        self.t = np.zeros(self.n_slots, dtype=np.int64)
        self.ram = self.rs.randint(0, 256, size=(self.n_slots, 128)).astype(np.uint8)

    def _draw_len(self, k):
        if isinstance(self.episode_len_spec, int):
            return np.full(k, self.episode_len_spec, dtype=np.int64)
        lo, hi = self.episode_len_spec
        return self.rs.randint(lo, hi + 1, size=k).astype(np.int64)

    def reset(self, slots):
        slots = np.asarray(slots, dtype=np.int64)
        self.ep_len[slots] = self._draw_len(len(slots))
        self.t[slots] = 0

    def step(self, slots, actions):
        slots = np.asarray(slots, dtype=np.int64)
        assert len(actions) == len(slots)
        self.t[slots] += 1
        rew = (self.rs.random_sample(len(slots)) < 0.05).astype(np.float32) * np.float32(10.0)
        done = self.t[slots] >= self.ep_len[slots]
        # behaviour characterisation stand-in: RAM drifts with the action taken
        self.ram[slots, self.t[slots] % 128] = (np.asarray(actions).astype(np.int64) * 13 + self.t[slots]) & 255
        return rew, done

    def advance(self):
        """Rotate the observation pool: the next forward sees a fresh block of frames (obs do not depend on the
        actions in the stub, but the engine still waits for the actions before calling ``step``)."""
        self._tick += 1
        self.obs = self.pool[self._tick % self.pool_blocks]

    def get_ram(self, slots):
        return self.ram[np.asarray(slots, dtype=np.int64)].copy()


class DeterministicAtariEnv(BatchEnv):
    """Atari-shaped test environment whose episodes are a pure function of the ACTIONS: the observation after t steps is
    frame ``t % R`` of one fixed seeded sequence, the reward is 10 when ``(7*action + t) % 11 == 0``, the episode length is
    fixed, the RAM drifts with the actions.  An episode's return / length / behaviour characterisation therefore depend
    only on the policy weights, not on which slot, wave or rank ran it: world-size-1 and world-size-N runs of a driver
    must agree exactly (tests/test_gpu_multi.py)."""

    def __init__(self, n_slots: int, num_actions: int = 18, episode_len: int = 6, seed: int = 0, frames: int = 8,
                 pin: bool = True):
        self.n_slots = int(n_slots)
        self.observation_space = Box(0, 255, (84, 84, 4), dtype=np.uint8)
        self.action_space = Discrete(num_actions)
        g = torch.Generator().manual_seed(seed)
        self.frames = torch.randint(0, 256, (frames, 84, 84, 4), dtype=torch.uint8, generator=g)
        obs = torch.zeros(self.n_slots, 84, 84, 4, dtype=torch.uint8)
        self.obs = obs.pin_memory() if (pin and torch.cuda.is_available()) else obs
        self.max_episode_steps = int(episode_len)
        self.t = np.zeros(self.n_slots, dtype=np.int64)
        self.ram = np.zeros((self.n_slots, 128), dtype=np.uint8)

    def reset(self, slots):
        slots = np.asarray(slots, dtype=np.int64)
        self.t[slots] = 0
        self.ram[slots] = 0
        self.obs[torch.from_numpy(slots)] = self.frames[0]

    def step(self, slots, actions):
        slots = np.asarray(slots, dtype=np.int64)
        a = np.asarray(actions).astype(np.int64)
        t = self.t[slots]
        rew = (((7 * a + t) % 11) == 0).astype(np.float32) * np.float32(10.0)
        self.ram[slots, t % 128] = (a * 13 + t + 1) & 255
        self.t[slots] = t + 1
        self.obs[torch.from_numpy(slots)] = self.frames[torch.from_numpy((t + 1) % len(self.frames))]
        return rew, self.t[slots] >= self.max_episode_steps

    def get_ram(self, slots):
        return self.ram[np.asarray(slots, dtype=np.int64)].copy()


class SyntheticVectorEnv(BatchEnv):
    """Humanoid-shaped stub: float32 observations ~ N(0,1) of dimension ``ob_dim``,
    continuous actions of dimension ``ac_dim``, reward = -|a|^2*1e-3 + 1 (alive bonus), fixed length."""

    def __init__(self, n_slots: int, ob_dim: int = 376, ac_dim: int = 17, episode_len: int = 1000, seed: int = 0,
                 pool_blocks: int = 4, pin: bool = True):
        self.n_slots = int(n_slots)
        self.observation_space = Box(-np.inf, np.inf, (ob_dim,))
        self.action_space = Box(-0.4, 0.4, (ac_dim,))
        g = torch.Generator().manual_seed(seed)
        pool = torch.randn(pool_blocks, self.n_slots, ob_dim, generator=g)
        self.pool = pool.pin_memory() if (pin and torch.cuda.is_available()) else pool
        self.pool_blocks = pool_blocks
        self.obs = self.pool[0]
        self._tick = 0
        self.max_episode_steps = int(episode_len)
        self.t = np.zeros(self.n_slots, dtype=np.int64)
        self.pos = np.zeros((self.n_slots, 2), dtype=np.float64)

    def reset(self, slots):
        slots = np.asarray(slots, dtype=np.int64)
        self.t[slots] = 0
        self.pos[slots] = 0

    def step(self, slots, actions):
        slots = np.asarray(slots, dtype=np.int64)
        a = np.asarray(actions, dtype=np.float32).reshape(len(slots), -1)
        self.t[slots] += 1
        self.pos[slots] += a[:, :2]
        rew = (1.0 - 1e-3 * np.square(a).sum(axis=1)).astype(np.float32)
        return rew, self.t[slots] >= self.max_episode_steps

    def advance(self):
        self._tick += 1
        self.obs = self.pool[self._tick % self.pool_blocks]

    def get_ram(self, slots):          # final (x, y) position BC (policies.py:292-299)
        return self.pos[np.asarray(slots, dtype=np.int64)].copy()

    def random_actions(self, k, rs):
        return rs.uniform(-0.4, 0.4, size=(k, self.action_space.shape[0])).astype(np.float32)


def make_env(env_id: str, n_slots: int, seed: int = 0, episode_len=None, allow_synthetic: bool = False, **kw) -> BatchEnv:
    """``gym.make(exp['env_id'])`` (es.py:131) for a whole slot table.

    ALE / gym / MuJoCo are not vendored by the reference and are absent from this image, so the only backends here are
    the synthetic stubs.  They are returned for the explicit ids ``SyntheticAtari*`` / ``SyntheticVector*``; for a REAL id
    (``FrostbiteNoFrameskip-v4``, ``Humanoid-v1`` ...) they are returned only when the caller opts in
    (``exp['allow_synthetic_env'] = true`` or ``DNE_ALLOW_SYNTHETIC_ENV=1``), with a loud warning -- a run that silently
    optimised random frames while logging and snapshotting like a real one would be worse than an error.  A real emulator
    backend registers itself in ``ENV_BACKENDS`` (id prefix -> factory)."""
    import logging
    import os
    for prefix, factory in ENV_BACKENDS.items():
        if env_id.startswith(prefix):
            return factory(env_id, n_slots, seed=seed, episode_len=episode_len, **kw)
    atari = env_id.endswith("NoFrameskip-v4") or env_id.startswith("SyntheticAtari")
    vector = env_id.startswith("Humanoid") or env_id.startswith("SyntheticVector")
    if not (atari or vector):
        raise KeyError(f"no environment backend for {env_id!r} in this build (gym/ALE/MuJoCo are not vendored)")
    if not env_id.startswith("Synthetic"):
        if not (allow_synthetic or os.environ.get("DNE_ALLOW_SYNTHETIC_ENV") == "1"):
            raise KeyError(f"no real environment backend for {env_id!r} in this build (gym/ALE/MuJoCo are not vendored); "
                           "register one in dne.envs.ENV_BACKENDS, or opt in to the synthetic stand-in with "
                           "exp['allow_synthetic_env'] = true / DNE_ALLOW_SYNTHETIC_ENV=1")
        logging.getLogger(__name__).warning(
            "SYNTHETIC ENVIRONMENT standing in for %r: observations are random frames / vectors and rewards are Bernoulli "
            "noise -- throughput measurements only, NOT training on the real task", env_id)
    if atari:
        env = SyntheticAtariEnv(n_slots, episode_len=episode_len if episode_len is not None else 1000, seed=seed, **kw)
    else:
        env = SyntheticVectorEnv(n_slots, episode_len=episode_len if episode_len is not None else 1000, seed=seed, **kw)
    env.synthetic = True
    return env


ENV_BACKENDS = {}        # id prefix -> factory(env_id, n_slots, seed=, episode_len=, **kw) -> BatchEnv (real emulators plug in here)
