"""Shared by tests/golden/make_golden_wrappers.py (build container) and tests/test_host.py: the emulator both sides step.

Frames carry only the RED channel: the reference's gray conversion is np.dot(obs float32, [0.299, 0.587, 0.114]) -- a BLAS sgemv
whose rounding depends on the build -- and with G = B = 0 it is fl(0.299f * R) in any summation order, with or without FMA, so
the whole wrapper pipeline is reproducible bit for bit (the Pillow resize is pinned exactly elsewhere)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-neuroevolution_b200")]
from dne.raw_env import ALE_ACTION_MEANINGS, SyntheticEmulator   # noqa: E402

ACTIONS = [3, 1, 0, 7, 7, 2, 5, 11, 0, 4, 9, 1, 6, 2, 2, 13, 8, 0, 3, 17, 5, 5, 1, 10, 0, 12, 4, 16, 2, 7] * 2
EMU_SEED, EMU_FRAMES, ENV_SEED = 5, 131, 77


class RedOnlyEmulator(SyntheticEmulator):
    action_meanings = list(ALE_ACTION_MEANINGS)           # [1] == 'FIRE': the fire-reset applies, as for Frostbite

    def _frame(self, a):
        f = super()._frame(a)
        f[:, :, 1:] = 0
        return f
