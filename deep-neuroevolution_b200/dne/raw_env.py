"""Raw-frame Atari environments behind the batched interface: emulators stepped on a host thread pool, the whole
observation pipeline of the reference on the device.

Reference pipeline per agent step (es_distributed/atari_wrappers.py ``wrap_deepmind``, :204-222):
  NoopResetEnv (:8-31, up to 30 no-ops)  ->  MaxAndSkipEnv (:86-107: repeat the action 4 times, sum the rewards, observation =
  per-pixel max of the last two raw frames)  ->  FireResetEnv (:33-48, games whose action 1 is FIRE -- Frostbite is one: on
  reset one agent step of FIRE and one of action 2)  ->  WarpFrame (:129-142: gray + PIL BILINEAR 210x160 -> 84x84 uint8)  ->
  FrameStack(4) (:167-180)  ->  ScaledFloatFrame (:182-186, /255: folded into the conv1 epilogue on the device).
The host-side logic (reset / skip / done handling) is pinned to the reference's own wrapper classes, executed on these
emulators under an old-gym stand-in (tests/golden/make_golden_wrappers.py, tests/test_host.py).
On the reference GPU path the emulators run on TF's CPU thread pool (gym_tensorflow/tf_env.cpp:231-316,
atari/tf_atari.cpp:24-128) and max / gray / resize / stack are TF ops (tf_atari.py:88-92, wrappers/stack_frames.py:33-43).

Here: ``RawFrameAtariEnv`` keeps one emulator per slot, steps the requested slots on a ``ThreadPoolExecutor`` (a real ALE
binding releases the GIL in ``act``), writes the LAST TWO raw frames of every agent step into a pinned host buffer
[n_slots, 2, 210, 160, 3] (RGB) and exposes ``device_obs(lo, hi)``: one pinned H2D copy of the raw pairs (2 x 100 KB per
slot), then ``dne_warp_atari_rgb`` (max + gray + Pillow-exact resize) and ``dne_preprocess_atari`` (frame stack, reset
convention of atari_wrappers.py:167-172) on the caller's stream -- the 84x84x4 uint8 stacks never exist on the host.
The rollout scheduler (dne/rollout.py) uses ``device_obs`` when an environment provides it.

ALE itself is not vendored by the reference and is absent from this image: ``SyntheticEmulator`` (deterministic frames
from (seed, t, action)) stands in for tests; ``ALEEmulator`` adapts ``ale_py`` when it is importable."""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Optional

import numpy as np
import torch

from . import _ffi as F
from .envs import BatchEnv, Box, Discrete

RAW_H, RAW_W = 210, 160
# ALE's 18 actions in index order (gym's ACTION_MEANING table): what FireResetEnv / NoopResetEnv look at
ALE_ACTION_MEANINGS = ["NOOP", "FIRE", "UP", "RIGHT", "LEFT", "DOWN", "UPRIGHT", "UPLEFT", "DOWNRIGHT", "DOWNLEFT", "UPFIRE", "RIGHTFIRE",
                       "LEFTFIRE", "DOWNFIRE", "UPRIGHTFIRE", "UPLEFTFIRE", "DOWNRIGHTFIRE", "DOWNLEFTFIRE"]


class Emulator:
    """One game instance.  ``reset() -> frame``, ``act(a) -> (reward, game_over, frame)`` with frame uint8 [210,160,3].
    ``action_meanings``: optional list of names (gym's get_action_meanings); [1] == 'FIRE' switches the fire-reset on."""
    num_actions = 18
    action_meanings = None

    def reset(self) -> np.ndarray:
        raise NotImplementedError

    def act(self, action: int):
        raise NotImplementedError


class SyntheticEmulator(Emulator):
    """Deterministic stand-in: the frame after t emulator steps is a seeded pattern shifted by t and tinted by the last
    action; reward 10 when (7*a + t) % 97 == 0; the game ends after ``frames`` emulator steps."""

    def __init__(self, seed: int, frames: int = 400, num_actions: int = 18):
        self.num_actions, self.frames, self.seed = num_actions, frames, seed
        rs = np.random.RandomState(seed)
        self.base = rs.randint(0, 256, size=(RAW_H, RAW_W, 3)).astype(np.uint8)
        self.t = 0

    def _frame(self, a):
        f = np.roll(self.base, (self.t % RAW_H, (3 * self.t) % RAW_W), axis=(0, 1)).copy()
        f[:, :, a % 3] = (f[:, :, a % 3].astype(np.int32) + 16 * a) & 255
        return f

    def reset(self):
        self.t = 0
        return self._frame(0)

    def act(self, action):
        self.t += 1
        rew = 10.0 if (7 * int(action) + self.t) % 97 == 0 else 0.0
        return rew, self.t >= self.frames, self._frame(int(action))


class ALEEmulator(Emulator):          # pragma: no cover - needs ale_py + a ROM, absent from this image
    def __init__(self, rom_path: str, seed: int = 0):
        import ale_py
        self.ale = ale_py.ALEInterface()
        self.ale.setInt("random_seed", seed)
        self.ale.setFloat("repeat_action_probability", 0.0)
        self.ale.loadROM(rom_path)
        self.actions = self.ale.getMinimalActionSet()
        self.num_actions = len(self.actions)
        self.action_meanings = [ALE_ACTION_MEANINGS[int(a)] for a in self.actions]

    def reset(self):
        self.ale.reset_game()
        return self.ale.getScreenRGB()

    def act(self, action):
        r = self.ale.act(self.actions[int(action)])
        return float(r), bool(self.ale.game_over()), self.ale.getScreenRGB()


class RawFrameAtariEnv(BatchEnv):
    def __init__(self, emulators: List[Emulator], frame_skip: int = 4, noop_max: int = 30, max_episode_steps: Optional[int] = None,
                 seed: int = 0, threads: Optional[int] = None, device=None, fire_reset: Optional[bool] = None,
                 noops: Optional[int] = None):
        """``fire_reset``: FireResetEnv of the reference (None = when the emulators' action 1 is 'FIRE', atari_wrappers.py:217);
        ``noops``: fixed number of reset no-ops (wrap_deepmind's ``noops`` override, :211-212) instead of a draw in [1, noop_max]."""
        self.emus = list(emulators)
        self.n_slots = len(self.emus)
        self.observation_space = Box(0, 255, (84, 84, 4), dtype=np.uint8)
        self.action_space = Discrete(self.emus[0].num_actions)
        self.skip, self.noop_max = int(frame_skip), int(noop_max)
        self.max_episode_steps = max_episode_steps
        self.rs = np.random.RandomState(seed)
        meanings = getattr(self.emus[0], "action_meanings", None)
        self.fire_reset = bool(meanings and len(meanings) >= 3 and meanings[1] == "FIRE") if fire_reset is None else bool(fire_reset)
        self.noops = noops
        if device is None:                                               # (host-only use -- the CPU tests -- needs no CUDA device)
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        pin = torch.cuda.is_available()
        raw = torch.zeros(self.n_slots, 2, RAW_H, RAW_W, 3, dtype=torch.uint8)
        self.raw = raw.pin_memory() if pin else raw                      # last two raw frames of the current agent step
        self._raw_np = self.raw.numpy()
        self.fresh = np.zeros(self.n_slots, dtype=np.uint8)              # episode just (re)started: stack = first frame x 4
        self.pool = ThreadPoolExecutor(max_workers=threads or min(32, self.n_slots))
        self._rs_lock = threading.Lock()
        self._dev = {}                                                   # (lo, hi) -> device buffers
        self.ram = np.zeros((self.n_slots, 128), dtype=np.uint8)

    # -- host side: emulator stepping on the thread pool -----------------------------------------------------------
    def _noop_reset(self, s, noops):
        f = self.emus[s].reset()
        for _ in range(noops):                                           # atari_wrappers.py:25-30 (action 0 = NOOP, raw env steps)
            _, over, f = self.emus[s].act(0)
            if over:
                f = self.emus[s].reset()
        self._raw_np[s, 0] = self._raw_np[s, 1] = f                      # MaxAndSkipEnv._reset (:109-114): buffer = [first frame]

    def _reset_one(self, s, noops):
        """NoopResetEnv._reset + MaxAndSkipEnv._reset (+ FireResetEnv._reset, atari_wrappers.py:40-48)."""
        self._noop_reset(s, noops)
        if self.fire_reset:
            for a in (1, 2):                                             # one agent step of FIRE, one of action 2
                _, over = self._step_one(s, a)
                if over:                                                 # "if done: self.env.reset()": the inner stack again, with
                    with self._rs_lock:                                  # a fresh no-op draw.  (After the action-2 step the
                        k = int(self._draw_noops(1)[0])                  # reference still returns that step's frame; here the
                    self._noop_reset(s, k)                               # fresh game's first frame -- unreachable with real games.)

    def _step_one(self, s, a):
        total, over = 0.0, False
        prev = cur = self._raw_np[s, 1]
        for _ in range(self.skip):                                       # atari_wrappers.py:95-107
            r, over, f = self.emus[s].act(a)
            prev, cur = cur, f
            total += r
            if over:
                break
        self._raw_np[s, 0], self._raw_np[s, 1] = prev, cur
        return total, over

    def _draw_noops(self, n):
        if self.noops is not None:
            return np.full(n, int(self.noops), dtype=np.int64)
        return self.rs.randint(1, self.noop_max + 1, size=n) if self.noop_max > 0 else np.zeros(n, dtype=np.int64)

    def reset(self, slots):
        slots = np.asarray(slots, dtype=np.int64)
        list(self.pool.map(self._reset_one, slots.tolist(), self._draw_noops(len(slots)).tolist()))
        self.fresh[slots] = 1

    def step(self, slots, actions):
        slots = np.asarray(slots, dtype=np.int64)
        res = list(self.pool.map(self._step_one, slots.tolist(), np.asarray(actions).astype(np.int64).tolist()))
        rew = np.array([r for r, _ in res], dtype=np.float32)
        done = np.array([d for _, d in res], dtype=bool)
        return rew, done

    def get_ram(self, slots):
        return self.ram[np.asarray(slots, dtype=np.int64)].copy()

    def obs_block(self, lo, hi):
        raise RuntimeError("RawFrameAtariEnv keeps its frame stacks on the device: use device_obs(lo, hi)")

    # -- device side: raw pairs -> max + gray + 84x84 warp -> frame stack ---------------------------------------------
    def device_obs(self, lo: int, hi: int) -> torch.Tensor:
        """uint8 [hi-lo, 84, 84, 4] frame stacks of slots [lo, hi) after this agent step, on the current CUDA stream."""
        n = hi - lo
        b = self._dev.get((lo, hi))
        if b is None:
            b = dict(raw=torch.empty(n, 2, RAW_H, RAW_W, 3, dtype=torch.uint8, device=self.device),
                     frame=torch.empty(n, 84, 84, dtype=torch.uint8, device=self.device),
                     stack=torch.zeros(n, 84, 84, 4, dtype=torch.uint8, device=self.device),
                     fresh=torch.zeros(n, dtype=torch.uint8, device=self.device),
                     fresh_host=(torch.zeros(n, dtype=torch.uint8).pin_memory() if torch.cuda.is_available() else torch.zeros(n, dtype=torch.uint8)))
            self._dev[(lo, hi)] = b
        b["raw"].copy_(self.raw[lo:hi], non_blocking=True)               # pinned -> HBM: 2 x 100 KB per slot
        b["fresh_host"].copy_(torch.from_numpy(self.fresh[lo:hi]))
        b["fresh"].copy_(b["fresh_host"], non_blocking=True)
        self.fresh[lo:hi] = 0
        L = F.lib()
        F.check(L.dne_warp_atari_rgb(F.ptr(b["raw"]), F.ptr(b["frame"]), n, F.stream_ptr()))        # atari_wrappers.py:105,138-142
        F.check(L.dne_preprocess_atari(None, F.ptr(b["frame"]), F.ptr(b["stack"]), F.ptr(b["fresh"]), n, 0, F.stream_ptr()))  # :167-180
        return b["stack"]


def make_synthetic_raw_env(n_slots: int, seed: int = 0, frames: int = 400, **kw) -> RawFrameAtariEnv:
    return RawFrameAtariEnv([SyntheticEmulator(seed * 100003 + s, frames=frames) for s in range(n_slots)], seed=seed, **kw)
