"""Generate tests/golden/ref_wrappers.npz by EXECUTING the reference's Atari wrapper stack.

Run in the build container only (imports /root/reference; nothing under tests/ reads that path at test time):
    python tests/golden/make_golden_wrappers.py

es_distributed/atari_wrappers.py `wrap_deepmind` (NoopResetEnv -> MaxAndSkipEnv -> FireResetEnv -> WarpFrame -> FrameStack ->
ScaledFloatFrame, atari_wrappers.py:204-222) is plain numpy + Pillow on top of gym's OLD wrapper API (`_step` / `_reset` /
`_observation` dispatch), and gym is absent from the image.  A ~40-line stand-in for that API lets the reference's own
classes run on a deterministic fake emulator (tests/golden/wrappers_common.py): every reset and agent step records the sha1
of the uint8 frame stack, the reward, the done flag and the emulator's raw step counter (which pins the number of no-ops, the
fire steps and the break-on-done of the skip loop)."""
import hashlib
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import wrappers_common as WC   # noqa: E402


def make_gym():
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Box:
        def __init__(self, low, high, shape=None):
            self.low, self.high, self.shape = low, high, tuple(shape) if shape is not None else np.shape(low)

    class Discrete:
        def __init__(self, n):
            self.n = n

    class Env:
        def step(self, a):
            return self._step(a)

        def reset(self):
            return self._reset()

        @property
        def unwrapped(self):
            return self

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env
            self.observation_space, self.action_space = env.observation_space, env.action_space

        def _step(self, a):
            return self.env.step(a)

        def _reset(self):
            return self.env.reset()

        @property
        def unwrapped(self):
            return self.env.unwrapped

        @property
        def spec(self):
            return self.env.spec

    class ObservationWrapper(Wrapper):
        def _reset(self):
            return self._observation(self.env.reset())

        def _step(self, a):
            obs, r, d, info = self.env.step(a)
            return self._observation(obs), r, d, info

    spaces.Box, spaces.Discrete = Box, Discrete
    gym.Env, gym.Wrapper, gym.ObservationWrapper, gym.RewardWrapper, gym.spaces = Env, Wrapper, ObservationWrapper, Wrapper, spaces
    return gym, spaces


def main():
    gym, spaces = make_gym()
    sys.modules["gym"], sys.modules["gym.spaces"] = gym, spaces
    pkg = types.ModuleType("refes")
    pkg.__path__ = ["/root/reference/es_distributed"]
    sys.modules["refes"] = pkg
    AW = importlib.import_module("refes.atari_wrappers")

    class FakeAtari(gym.Env):
        def __init__(self):
            self.emu = WC.RedOnlyEmulator(WC.EMU_SEED, frames=WC.EMU_FRAMES)
            self.np_random = np.random.RandomState(WC.ENV_SEED)
            self.observation_space = spaces.Box(0, 255, (210, 160, 3))
            self.action_space = spaces.Discrete(18)
            self.spec = types.SimpleNamespace(id="FrostbiteNoFrameskip-v4")

        def get_action_meanings(self):
            return list(WC.RedOnlyEmulator.action_meanings)

        def _step(self, a):
            r, over, f = self.emu.act(int(a))
            return f, r, over, {}

        def _reset(self):
            return self.emu.reset()

    raw = FakeAtari()
    env = AW.wrap_deepmind(raw)
    kinds, sha, rew, done, tcount, first = [], [], [], [], [], None

    def rec(kind, obs, r, d):
        nonlocal first
        u8 = np.rint(np.asarray(obs, dtype=np.float32) * 255.0).astype(np.uint8)
        assert u8.shape == (84, 84, 4) and np.array_equal(u8.astype(np.float32) / 255.0, obs)
        if first is None:
            first = u8.copy()
        kinds.append(kind); sha.append(hashlib.sha1(np.ascontiguousarray(u8).tobytes()).hexdigest())
        rew.append(float(r)); done.append(bool(d)); tcount.append(int(raw.emu.t))

    rec(0, env.reset(), 0.0, False)
    for a in WC.ACTIONS:
        obs, r, d, _ = env.step(a)
        rec(1, obs, r, d)
        if d:
            rec(0, env.reset(), 0.0, False)
    out = dict(kind=np.array(kinds, np.int8), sha1=np.array(sha), reward=np.array(rew, np.float64), done=np.array(done),
               emu_t=np.array(tcount, np.int64), first_stack=first)
    np.savez_compressed(os.path.join(HERE, "ref_wrappers.npz"), **out)
    print("events", len(kinds), "resets", int((np.array(kinds) == 0).sum()), "dones", int(np.sum(done)), "emu_t", tcount[:8], "...")
    print("wrote", os.path.join(HERE, "ref_wrappers.npz"))


if __name__ == "__main__":
    main()
