"""Generate tests/golden/ref_numpy.npz by calling the REFERENCE's own numpy functions.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Imports es_distributed.es / es_distributed.optimizers from /root/reference with a stub ``redis`` module
(es.py:8 -> dist.py:8 imports redis; nothing else is needed for the numpy-only functions) and records
their outputs on seeded inputs.  The fixtures pin oracle/oracle.py (tests/test_oracle.py) and, through it,
the CUDA path.

numpy note: the reference targets numpy 1.12; this container has numpy 2.x, under which
``Adam._compute_step`` / ``SGD._compute_step`` silently promote to float64 (NEP 50).  The optimizer
outputs recorded here are therefore float64 and the float32 oracle/kernels are compared to them with a
relative tolerance (1e-6), not bit-exactly.  Everything else recorded here is dtype-stable.
"""
import os
import sys
import types

import numpy as np

sys.modules["redis"] = types.ModuleType("redis")
sys.path.insert(0, "/root/reference")
import es_distributed.es as ref_es            # noqa: E402
import es_distributed.optimizers as ref_opt   # noqa: E402

out = {}
rs = np.random.RandomState(20260922)

# --- noise table prefix (es.py:51-67) ------------------------------------------------------------
COUNT = 400_000
noise = np.random.RandomState(123).randn(COUNT).astype(np.float32)       # exactly es.py:60 on a prefix
out["noise_count"] = np.int64(COUNT)
out["noise_head"] = noise[:64].copy()
out["noise_tail"] = noise[-64:].copy()
out["noise_sum64"] = np.float64(noise.astype(np.float64).sum())

# sample_index stream (es.py:66-67) with a seeded RandomState and the real table length
class _T:  # minimal stand-in exposing len(self.noise), calling the unbound reference method
    noise = np.empty(250_000_000, dtype=np.int8)
stream = np.random.RandomState(7)
out["sample_index_P4052658"] = np.array(
    [ref_es.SharedNoiseTable.sample_index(_T, stream, 4052658) for _ in range(16)], dtype=np.int64)

# --- ranks (es.py:70-85), tie-free inputs ---------------------------------------------------------
for n in (1, 8, 500, 5000):
    x = (rs.permutation(2 * n).astype(np.float32) * np.float32(0.731) - np.float32(n)).reshape(n, 2)  # distinct values
    assert len(np.unique(x)) == x.size
    out[f"rank_in_{n}"] = x
    out[f"rank_ranks_{n}"] = ref_es.compute_ranks(x.ravel()).astype(np.int64)
    out[f"rank_centered_{n}"] = ref_es.compute_centered_ranks(x)

# --- batched_weighted_sum + normalise (es.py:115-122, 291-296) -----------------------------------
P = 3001
n = 1203                        # > 2 slabs of 500
idx = rs.randint(0, COUNT - P + 1, size=n).astype(np.int64)
returns = (rs.permutation(2 * n).astype(np.float32) * np.float32(10.0)).reshape(n, 2)   # tie-free
proc = ref_es.compute_centered_ranks(returns)
g, count = ref_es.batched_weighted_sum(proc[:, 0] - proc[:, 1], (noise[i:i + P] for i in idx), batch_size=500)
g = g / returns.size
assert count == n and g.dtype == np.float32
out["grad_P"], out["grad_idx"], out["grad_returns"], out["grad_g"] = np.int64(P), idx, returns, g

# --- optimizers (optimizers.py) -------------------------------------------------------------------
theta0 = rs.randn(P).astype(np.float32)
grads = [rs.randn(P).astype(np.float32) * 0.1 for _ in range(3)]
out["opt_theta0"] = theta0
out["opt_grads"] = np.stack(grads)
adam = ref_opt.Adam(theta0.copy(), stepsize=0.01)
sgd = ref_opt.SGD(theta0.copy(), stepsize=0.01, momentum=0.9)
a_ratio, a_theta, s_ratio, s_theta = [], [], [], []
for gk in grads:
    r, t = adam.update(-gk + 0.005 * adam.theta.astype(np.float32))       # es.py:298
    a_ratio.append(r); a_theta.append(np.asarray(t, dtype=np.float64))
    r, t = sgd.update(-gk + 0.005 * sgd.theta.astype(np.float32))
    s_ratio.append(r); s_theta.append(np.asarray(t, dtype=np.float64))
out["adam_ratio"], out["adam_theta"] = np.array(a_ratio, dtype=np.float64), np.stack(a_theta)
out["sgd_ratio"], out["sgd_theta"] = np.array(s_ratio, dtype=np.float64), np.stack(s_theta)

# --- RunningStat (es.py:26-48) --------------------------------------------------------------------
st = ref_es.RunningStat((5,), eps=1e-2)
obs = rs.randn(37, 5).astype(np.float32) * 3 + 1
st.increment(obs.sum(axis=0), np.square(obs).sum(axis=0), len(obs))
out["rstat_obs"], out["rstat_mean"], out["rstat_std"] = obs, st.mean, st.std

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_numpy.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: np.asarray(v).shape for k, v in out.items()})
