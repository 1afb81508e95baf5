"""CPU oracle for the ES/GA rollout-and-update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this module: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may.  It is the checker, never the thing measured or shipped.

It is a numpy / torch-CPU(fp32) restatement of the reference algorithm, each function citing the
``/root/reference`` file:line it follows.  The reference is Python, so the oracle is Python.

Pin status
----------
* numpy-only reference functions (ranks, weighted sum, RunningStat, noise table, SGD/Adam) are pinned
  against outputs of the reference itself, imported in the build container by
  ``tests/golden/make_golden.py`` (fixtures: ``tests/golden/ref_numpy.npz``).
* TensorFlow-dependent pieces (conv/dense forward, virtual batch norm, GA ``reinitialize``) cannot be
  imported here (no tensorflow / h5py / gym in the image) and the reference ships no
  tests or golden vectors: for those **parity is unpinned** -- they follow the cited
  source lines and the documented TF semantics (SAME padding, HWIO kernels, NHWC, biased batch
  variance), and are cross-checked against an independent naive-loop numpy implementation in
  ``tests/test_oracle.py``.

numpy-2 hazard: the reference targets numpy 1.12 where ``float64 scalar * float32
array`` stays float32.  Under numpy 2 the same source promotes to float64.  The restatement below pins
the numpy-1.12 float32 semantics explicitly (scalars rounded to float32 first).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32

# --------------------------------------------------------------------------------------------------
# a-1  Shared noise table                                   es_distributed/es.py:51-67
# --------------------------------------------------------------------------------------------------
NOISE_SEED = 123            # es.py:54
NOISE_COUNT = 250_000_000   # es.py:55


def noise_table(count: int = NOISE_COUNT, seed: int = NOISE_SEED) -> np.ndarray:
    """es.py:60 -- ``RandomState(seed).randn(count)`` cast float64 -> float32.

    The legacy MT19937 + polar Box-Muller stream is frozen by numpy, so a *prefix* of the full table is
    obtained by asking for a smaller (even) count (gauss() produces values in pairs)."""
    return np.random.RandomState(seed).randn(count).astype(f32)


def sample_index(stream: np.random.RandomState, table_len: int, dim: int) -> int:
    """es.py:66-67."""
    return int(stream.randint(0, table_len - dim + 1))


# --------------------------------------------------------------------------------------------------
# a-8  ranks                                                 es_distributed/es.py:70-85
# --------------------------------------------------------------------------------------------------

def compute_ranks(x: np.ndarray) -> np.ndarray:
    """es.py:70-78 with the canonical *stable* tie rule: ascending by (value, flat
    index).  For tie-free input this is bit-identical to the reference's ``x.argsort()``."""
    assert x.ndim == 1
    ranks = np.empty(len(x), dtype=np.int64)
    ranks[np.argsort(x, kind="stable")] = np.arange(len(x))
    return ranks


def compute_centered_ranks(x: np.ndarray) -> np.ndarray:
    """es.py:81-85: float32 ranks, ``/= (size-1)`` then ``-= .5`` in float32."""
    y = compute_ranks(x.ravel()).reshape(x.shape).astype(f32)
    y /= f32(x.size - 1)
    y -= f32(0.5)
    return y


# --------------------------------------------------------------------------------------------------
# a-9  ES gradient                                           es_distributed/es.py:115-122, 291-296
# --------------------------------------------------------------------------------------------------

def batched_weighted_sum(weights: np.ndarray, noise: np.ndarray, idx: Sequence[int], dim: int,
                         batch_size: int = 500) -> np.ndarray:
    """es.py:115-122: float32 ``np.dot(w[b], V[b, dim])`` in slabs of ``batch_size`` rows, running
    float32 add."""
    total = np.zeros(dim, dtype=f32)
    for s in range(0, len(idx), batch_size):
        V = np.stack([noise[i:i + dim] for i in idx[s:s + batch_size]]).astype(f32)
        total = (total + np.dot(np.asarray(weights[s:s + batch_size], dtype=f32), V)).astype(f32)
    return total


def es_gradient(proc_returns_n2: np.ndarray, noise: np.ndarray, idx: Sequence[int], dim: int,
                dtype=np.float64) -> np.ndarray:
    """es.py:291-296: ``g = sum_i (y[i,0]-y[i,1]) * noise[idx_i : idx_i+dim] / returns_n2.size``.

    ``dtype=float64`` is the referee used to arbitrate the 1e-5 relative tolerance (the reference's own
    float32 summation order is BLAS-dependent); ``dtype=float32`` follows es.py:115-122 literally."""
    w = (proc_returns_n2[:, 0] - proc_returns_n2[:, 1]).astype(f32)   # float32 subtraction (es.py:292)
    if dtype == np.float32:
        g = batched_weighted_sum(w, noise, idx, dim)
        g /= f32(proc_returns_n2.size)
        return g
    g = np.zeros(dim, dtype=np.float64)
    for wi, i in zip(w, idx):
        g += np.float64(wi) * noise[i:i + dim].astype(np.float64)
    return g / proc_returns_n2.size


# --------------------------------------------------------------------------------------------------
# a-10 optimizers                                            es_distributed/optimizers.py:4-50
# --------------------------------------------------------------------------------------------------

class Adam:
    """optimizers.py:35-50, numpy-1.12 float32 semantics pinned (module docstring)."""

    def __init__(self, theta, stepsize, beta1=0.9, beta2=0.999, epsilon=1e-08):
        self.theta = np.asarray(theta, dtype=f32).copy()
        self.stepsize, self.beta1, self.beta2, self.epsilon = stepsize, beta1, beta2, epsilon
        self.m = np.zeros_like(self.theta)
        self.v = np.zeros_like(self.theta)
        self.t = 0

    def step_scale(self) -> float:
        # optimizers.py:46, evaluated in float64 python scalars then rounded to f32 when it meets the array
        return self.stepsize * math.sqrt(1 - self.beta2 ** self.t) / (1 - self.beta1 ** self.t)

    def update(self, globalg):
        self.t += 1                                                            # optimizers.py:11
        g = np.asarray(globalg, dtype=f32)
        a = f32(self.step_scale())
        self.m = f32(self.beta1) * self.m + f32(1 - self.beta1) * g           # :47
        self.v = f32(self.beta2) * self.v + f32(1 - self.beta2) * (g * g)     # :48
        step = (-a) * self.m / (np.sqrt(self.v) + f32(self.epsilon))          # :49
        ratio = np.linalg.norm(step) / np.linalg.norm(self.theta)             # :14
        self.theta = (self.theta + step).astype(f32)                           # :15
        return ratio, self.theta


class SGD:
    """optimizers.py:23-32 (momentum EMA form ``v = m*v + (1-m)*g``)."""

    def __init__(self, theta, stepsize, momentum=0.9):
        self.theta = np.asarray(theta, dtype=f32).copy()
        self.v = np.zeros_like(self.theta)
        self.stepsize, self.momentum = stepsize, momentum
        self.t = 0

    def update(self, globalg):
        self.t += 1
        g = np.asarray(globalg, dtype=f32)
        self.v = f32(self.momentum) * self.v + f32(1. - self.momentum) * g    # :30
        step = f32(-self.stepsize) * self.v                                    # :31
        ratio = np.linalg.norm(step) / np.linalg.norm(self.theta)
        self.theta = (self.theta + step).astype(f32)
        return ratio, self.theta


def es_update_direction(g: np.ndarray, theta: np.ndarray, l2coeff: float) -> np.ndarray:
    """es.py:298 -- ``-g + l2coeff * theta`` in float32."""
    return (-g.astype(f32) + f32(l2coeff) * theta.astype(f32)).astype(f32)


class RunningStat:
    """es.py:26-48."""

    def __init__(self, shape, eps):
        self.sum = np.zeros(shape, dtype=f32)
        self.sumsq = np.full(shape, eps, dtype=f32)
        self.count = eps

    def increment(self, s, ssq, c):
        self.sum += s
        self.sumsq += ssq
        self.count += c

    @property
    def mean(self):
        return self.sum / self.count

    @property
    def std(self):
        return np.sqrt(np.maximum(self.sumsq / self.count - np.square(self.mean), 1e-2))


# --------------------------------------------------------------------------------------------------
# a-3  flat parameter layouts                                tf_util.py:224-246, models/base.py:165-192
# --------------------------------------------------------------------------------------------------

@dataclass
class Var:
    name: str
    shape: Tuple[int, ...]
    kind: str                 # 'w' | 'b' | 'beta' | 'gamma'
    std: float = 1.0          # normc / scale_by std for 'w'
    offset: int = 0

    @property
    def size(self) -> int:
        return int(np.prod(self.shape))


@dataclass
class Layer:
    kind: str                 # 'conv' | 'dense'
    cin: int
    cout: int
    ksize: int = 1
    stride: int = 1
    hin: int = 1              # conv: input spatial size (square)
    act: str = "relu"         # 'relu' | 'tanh' | 'none'
    bias: bool = True
    bn: str = "none"          # 'none' | 'tf' (contrib batch_norm, gamma/beta) | 'vbn_gpu'
    std: float = 1.0
    vars: List[Var] = field(default_factory=list)

    @property
    def hout(self) -> int:
        return -(-self.hin // self.stride)            # TF SAME: ceil(in/stride)

    @property
    def pad_before(self) -> int:
        total = max((self.hout - 1) * self.stride + self.ksize - self.hin, 0)
        return total // 2                              # TF SAME: extra pixel goes after


@dataclass
class Net:
    name: str
    layers: List[Layer]
    ob_shape: Tuple[int, ...]
    num_params: int = 0
    ob_norm: bool = False     # MujocoPolicy clip((o-mean)/std, -5, 5)

    def variables(self) -> List[Var]:
        return [v for l in self.layers for v in l.vars]


def _finish(net: Net) -> Net:
    off = 0
    for l in net.layers:
        for v in l.vars:
            v.offset = off
            off += v.size
    net.num_params = off
    return net


def make_net(name: str, num_actions: int = 18, ob_dim: int = 376, hidden=(256, 256), ac_dim: int = 17) -> Net:
    """Flat layouts in variable-creation order.

    LargeModel      gpu_implementation/neuroevolution/models/dqn.py:39-47 (+ base.py:54-99): w then b per layer
    Model           dqn.py:25-36
    GAAtariPolicy   es_distributed/policies.py:449-459 (tf_util.py:133-162): name/w, name/b
    ESAtariPolicy   policies.py:319-330: weights, biases, BatchNorm/beta, BatchNorm/gamma per BN'd layer
    MujocoPolicy    policies.py:155-162,195-196 ('continuous:' head, out layer normc 0.01)
    """
    def conv(cin, cout, k, s, hin, **kw):
        l = Layer("conv", cin, cout, k, s, hin, **kw)
        l.vars.append(Var("w", (k, k, cin, cout), "w", l.std))
        if l.bias:
            l.vars.append(Var("b", (cout,), "b"))
        if l.bn == "tf":
            l.vars += [Var("beta", (cout,), "beta"), Var("gamma", (cout,), "gamma")]
        return l

    def dense(cin, cout, **kw):
        l = Layer("dense", cin, cout, **kw)
        l.vars.append(Var("w", (cin, cout), "w", l.std))
        if l.bias:
            l.vars.append(Var("b", (cout,), "b"))
        if l.bn == "tf":
            l.vars += [Var("beta", (cout,), "beta"), Var("gamma", (cout,), "gamma")]
        return l

    A = num_actions
    if name == "LargeModel":
        layers = [conv(4, 32, 8, 4, 84), conv(32, 64, 4, 2, 21), conv(64, 64, 3, 1, 11),
                  dense(11 * 11 * 64, 512), dense(512, A, act="none", std=0.1)]
        return _finish(Net(name, layers, (84, 84, 4)))
    if name in ("Model", "GAAtariPolicy"):
        layers = [conv(4, 16, 8, 4, 84), conv(16, 32, 4, 2, 21),
                  dense(11 * 11 * 32, 256), dense(256, A, act="none", std=0.1)]
        return _finish(Net(name, layers, (84, 84, 4)))
    if name == "ESAtariPolicy":
        layers = [conv(4, 16, 8, 4, 84, bn="tf"), conv(16, 32, 4, 2, 21, bn="tf"),
                  dense(11 * 11 * 32, 256, bn="tf"), dense(256, A, act="none")]
        return _finish(Net(name, layers, (84, 84, 4)))
    if name == "ModelVirtualBN":
        # gpu_implementation/neuroevolution/models/batchnorm.py:50-123: conv / dense WITHOUT bias, then
        # (x - mean) * 1/sqrt(var + 1e-3) + b with a per-channel bias b created inside the BatchNorm scope (no gamma):
        # creation order w, b per layer -- the same flat layout as Model, but 'b' acts AFTER the normalisation
        # the output layer is created with the DEFAULT std = 1.0 (batchnorm.py:106, unlike Model / LargeModel's 0.1): found by
        # executing the reference class (tests/golden/make_golden_models.py)
        layers = [conv(4, 16, 8, 4, 84, bn="vbn_gpu"), conv(16, 32, 4, 2, 21, bn="vbn_gpu"),
                  dense(11 * 11 * 32, 256, bn="vbn_gpu"), dense(256, A, act="none", std=1.0)]
        return _finish(Net(name, layers, (84, 84, 4)))
    if name == "MujocoPolicy":
        dims = [ob_dim] + list(hidden)
        layers = [dense(dims[i], dims[i + 1], act="tanh") for i in range(len(hidden))]
        layers.append(dense(dims[-1], ac_dim, act="none", std=0.01))
        return _finish(Net(name, layers, (ob_dim,), ob_norm=True))
    raise KeyError(name)


def unflatten(net: Net, theta: np.ndarray):
    """tf_util.py:224-240 (SetFromFlat): C-order reshape of consecutive slices."""
    assert theta.shape == (net.num_params,)
    return [{v.kind: theta[v.offset:v.offset + v.size].reshape(v.shape) for v in l.vars} for l in net.layers]


# --------------------------------------------------------------------------------------------------
# a-2  perturb                                               es_distributed/es.py:412-419
# --------------------------------------------------------------------------------------------------

def perturb(theta: np.ndarray, noise: np.ndarray, idx: int, sigma: float, sign: int) -> np.ndarray:
    """es.py:413-419: ``v = noise_stdev * noise[idx:idx+P]`` (float32), then ``theta + v`` / ``theta - v``."""
    v = f32(sigma) * noise[idx:idx + theta.size]
    return (theta + v).astype(f32) if sign > 0 else (theta - v).astype(f32)


# --------------------------------------------------------------------------------------------------
# a-4 / a-5 / a-13  forward                                  policies.py:319-330,449-459,150-162; dqn.py:39-47
# --------------------------------------------------------------------------------------------------
BN_EPS = 1e-3   # policies.py:322


def _conv_same(x_nhwc, w_hwio, stride):
    """tf.nn.conv2d(x, w, [1,s,s,1], 'SAME') on NHWC/HWIO (tf_util.py:139; models/base.py:59-72)."""
    import torch
    import torch.nn.functional as F
    n, h, _, _ = x_nhwc.shape
    k = w_hwio.shape[0]
    hout = -(-h // stride)
    total = max((hout - 1) * stride + k - h, 0)
    pb, pa = total // 2, total - total // 2
    xt = torch.from_numpy(np.ascontiguousarray(x_nhwc)).permute(0, 3, 1, 2)
    wt = torch.from_numpy(np.ascontiguousarray(w_hwio)).permute(3, 2, 0, 1)
    y = F.conv2d(F.pad(xt, (pb, pa, pb, pa)), wt, stride=stride)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def _act(x, kind):
    if kind == "relu":
        return np.maximum(x, f32(0))
    if kind == "tanh":
        return np.tanh(x).astype(f32)
    return x


def forward(net: Net, theta: np.ndarray, obs: np.ndarray, *, vbn_stats=None, is_ref: bool = False,
            ob_mean=None, ob_std=None, return_all: bool = False):
    """Batched forward with ONE weight vector.

    obs: Atari uint8 [B,84,84,4] (scaled /255, atari_wrappers.py:186) or float32 [B,ob_dim].
    ESAtariPolicy: ``is_ref=True`` runs the virtual-batch-norm reference pass (policies.py:322-328,399:
    ``batch_norm(scale=True, is_training=True, decay=0., epsilon=1e-3)``) and returns the per-layer
    (mean, biased var) it stored; otherwise ``vbn_stats`` from a previous reference pass is used.
    Returns (logits [B,A], stats) ; action = argmax(logits) (policies.py:330,459).
    """
    import torch
    torch.set_num_threads(1)
    params = unflatten(net, theta.astype(f32))
    if obs.dtype == np.uint8:
        x = obs.astype(f32) / f32(255.0)
    else:
        x = obs.astype(f32)
    if net.ob_norm:  # policies.py:151
        x = np.clip((x - ob_mean.astype(f32)) / ob_std.astype(f32), f32(-5.0), f32(5.0)).astype(f32)
    stats_out, acts = [], []
    bn_i = 0
    for l, p in zip(net.layers, params):
        if l.kind == "conv":
            y = _conv_same(x, p["w"], l.stride)
        else:
            if x.ndim > 2:
                x = x.reshape(x.shape[0], -1)       # flatten (h,w,c)  tf_util.py:284-285
            y = (torch.from_numpy(np.ascontiguousarray(x)) @ torch.from_numpy(np.ascontiguousarray(p["w"]))).numpy()
        if l.bias and l.bn != "vbn_gpu":
            y = y + p["b"].reshape((1,) * (y.ndim - 1) + (-1,))
        if l.bn == "vbn_gpu":                     # batchnorm.py:64-93: moments over the batch / spatial axes, inv std stored
            axes = tuple(range(y.ndim - 1))
            if is_ref:
                mean = y.mean(axis=axes, dtype=np.float64).astype(f32)
                var = np.square(y.astype(np.float64) - mean.astype(np.float64)).mean(axis=axes).astype(f32)
                stats_out.append((mean, var))
            else:
                mean, var = vbn_stats[bn_i]
            bn_i += 1
            inv = (f32(1.0) / np.sqrt(var + f32(BN_EPS))).astype(f32)
            y = ((y - mean) * inv + p["b"]).astype(f32)
        if l.bn == "tf":
            axes = tuple(range(y.ndim - 1))
            if is_ref:
                mean = y.mean(axis=axes, dtype=np.float64).astype(f32)
                var = np.square(y.astype(np.float64) - mean.astype(np.float64)).mean(axis=axes).astype(f32)
                stats_out.append((mean, var))
            else:
                mean, var = vbn_stats[bn_i]
            bn_i += 1
            inv = (f32(1.0) / np.sqrt(var + f32(BN_EPS))).astype(f32)
            y = ((y - mean) * inv * p["gamma"] + p["beta"]).astype(f32)
        x = _act(y.astype(f32), l.act)
        acts.append(x)
    if return_all:
        return x, stats_out, acts
    return x, stats_out


def act(net: Net, theta, obs, **kw) -> np.ndarray:
    logits, _ = forward(net, theta, obs, **kw)
    return np.argmax(logits, axis=1)      # first max on ties (numpy / TF convention)


def forward_naive_conv(x_nhwc: np.ndarray, w_hwio: np.ndarray, stride: int) -> np.ndarray:
    """Independent naive-loop SAME conv (float64 accumulate) used only to cross-check ``_conv_same``."""
    n, h, w_, cin = x_nhwc.shape
    k, _, _, cout = w_hwio.shape
    hout = -(-h // stride)
    total = max((hout - 1) * stride + k - h, 0)
    pb = total // 2
    out = np.zeros((n, hout, hout, cout), dtype=np.float64)
    for oy in range(hout):
        for ox in range(hout):
            for ky in range(k):
                iy = oy * stride - pb + ky
                if iy < 0 or iy >= h:
                    continue
                for kx in range(k):
                    ix = ox * stride - pb + kx
                    if ix < 0 or ix >= w_:
                        continue
                    out[:, oy, ox, :] += x_nhwc[:, iy, ix, :].astype(np.float64) @ w_hwio[ky, kx].astype(np.float64)
    return out


# --------------------------------------------------------------------------------------------------
# a-11  GA genome -> weights, truncation selection           ga.py:135-158,250-264; models/base.py:123-156
# --------------------------------------------------------------------------------------------------

def ga_scale_by(net: Net) -> np.ndarray:
    """models/base.py:165-178 + dqn.py:26-28: per-variable ``std/sqrt(prod(shape[:-1]))``, 0 for biases."""
    s = np.zeros(net.num_params, dtype=f32)
    for v in net.variables():
        if v.kind == "w":
            s[v.offset:v.offset + v.size] = f32(v.std / np.sqrt(np.prod(v.shape[:-1])))
    return s


def ga_materialize_gpu(net: Net, noise: np.ndarray, seeds) -> np.ndarray:
    """models/base.py:140-146,155-156: ``theta = noise[idx0]*scale_by; theta += power_k*noise[idx_k]``.
    seeds = (idx0, (idx1, power1), ...)."""
    P = net.num_params
    theta = (noise[seeds[0]:seeds[0] + P].copy() * ga_scale_by(net)).astype(f32)
    for idx, power in seeds[1:]:
        theta = (theta + f32(power) * noise[idx:idx + P]).astype(f32)
    return theta


def ga_reinitialize(net: Net, theta: np.ndarray) -> np.ndarray:
    """policies.py:42-44 + tf_util.py:122-130,137,143,152,158: every weight matrix reshaped to
    [-1, n_out] and each output column rescaled to L2 norm ``std``; biases set to zero."""
    out = theta.astype(f32).copy()
    for v in net.variables():
        sl = slice(v.offset, v.offset + v.size)
        if v.kind == "w":
            m = out[sl].reshape(-1, v.shape[-1])
            m *= f32(v.std) / np.sqrt(np.square(m).sum(axis=0, keepdims=True))
            out[sl] = m.reshape(-1)
        else:
            out[sl] = 0
    return out


def ga_materialize_cpu(net: Net, noise: np.ndarray, seeds: Sequence[int], sigma: float) -> np.ndarray:
    """ga.py:256-264: ``v = reinitialize(noise[seed0]); v += noise_stdev*noise[seed]`` for later seeds."""
    P = net.num_params
    v = ga_reinitialize(net, noise[seeds[0]:seeds[0] + P])
    for s in seeds[1:]:
        v += f32(sigma) * noise[s:s + P]
    return v


def ga_truncate(fitness: np.ndarray, T: int) -> np.ndarray:
    """Top-T indices, descending fitness, ties by arrival order (canonical stable rule)
    == Python ``sorted(key=fitness, reverse=True)`` as at gpu_implementation/ga.py:180, and the
    ``ga.py:145-147`` argpartition result for tie-free input."""
    order = np.argsort(-fitness.astype(np.float64), kind="stable")
    return order[:T].astype(np.int32)


def deep_ga_validation_population(population_sorted, elite, validation_threshold: int):
    """gpu_implementation/ga.py:185-188: the top ``validation_threshold`` individuals of the fitness-sorted population,
    with last generation's elite put first (dropping the last of them)."""
    val = list(population_sorted[:validation_threshold])
    if elite is not None:
        val = [elite] + val[:-1]
    return val


def deep_ga_elite(validation_population, validation_returns):
    """gpu_implementation/ga.py:191-198: mean validation return per candidate; elite = argmax (first maximum)."""
    means = [float(np.mean(r)) for r in validation_returns]
    return validation_population[int(np.argmax(means))], means


def deep_ga_parents(population_sorted, elite, selection_threshold: int):
    """gpu_implementation/ga.py:260-271: the top ``selection_threshold`` individuals; if the elite is not among them it
    takes the first place and the last of them is dropped."""
    top = list(population_sorted[:selection_threshold])
    if elite in top:
        return top
    return [elite] + top[:selection_threshold - 1]


# --------------------------------------------------------------------------------------------------
# a-12  novelty                                              es_distributed/nses.py:12-32,226-228
# --------------------------------------------------------------------------------------------------

def euclidean_distance(x: np.ndarray, y: np.ndarray) -> float:
    """nses.py:12-20 (shorter BC sequence is padded with its last row)."""
    n, m = len(x), len(y)
    if n > m:
        a = np.linalg.norm(y - x[:m])
        b = np.linalg.norm(y[-1] - x[m:])
    else:
        a = np.linalg.norm(x - y[:n])
        b = np.linalg.norm(x[-1] - y[n:])
    return np.sqrt(a ** 2 + b ** 2)


def compute_novelty_vs_archive(archive, novelty_vector, k: int) -> float:
    """nses.py:22-32 (``np.float`` -> float64)."""
    nov = novelty_vector.astype(np.float64)
    distances = np.array([euclidean_distance(p.astype(np.float64), nov) for p in archive])
    top_k = np.sort(distances, kind="stable")[:k]
    return top_k.mean()


def nsr_blend(returns_n2, novelty_n2):
    """nses.py:221-228 with return_proc_mode=centered_sign_rank: (rank(reward)+rank(novelty))/2."""
    return ((compute_centered_ranks(returns_n2) + compute_centered_ranks(novelty_n2)) / f32(2.0)).astype(f32)


# --------------------------------------------------------------------------------------------------
# a-6  observation preprocess                                atari_wrappers.py:105,167-186; stack_frames.py:33-43
# --------------------------------------------------------------------------------------------------

def max_and_stack(frames_prev: np.ndarray, frames_cur: np.ndarray, stack: np.ndarray, reset_mask: np.ndarray,
                  mode: str = "cpu") -> np.ndarray:
    """Max over the last two 84x84 uint8 frames (atari_wrappers.py:105 / tf_atari.py:90), then frame stack
    k=4 on the channel axis.  mode 'cpu' (atari_wrappers.py:167-180): reset fills all 4 channels with the
    first frame; mode 'gpu' (stack_frames.py:33-43): reset = three zero frames + the first frame.
    Otherwise: shift channels left by one and append."""
    new = np.maximum(frames_prev, frames_cur)                       # [S,84,84] uint8
    out = stack.copy()
    for s in range(stack.shape[0]):
        if reset_mask[s]:
            if mode == "cpu":
                out[s] = new[s][:, :, None]
            else:
                out[s] = 0
                out[s, :, :, 3] = new[s]
        else:
            out[s, :, :, :3] = stack[s, :, :, 1:]
            out[s, :, :, 3] = new[s]
    return out


# --------------------------------------------------------------------------------------------------
# a-6  210x160 -> 84x84 warp, both reference flavours
# --------------------------------------------------------------------------------------------------

def gray_rgb(obs_rgb_u8: np.ndarray) -> np.ndarray:
    """atari_wrappers.py:139: ``np.dot(obs.astype('float32'), [0.299, 0.587, 0.114] float32)``.  numpy hands this to the
    BLAS sgemv of its build, whose summation order / FMA use is build specific (here np.dot, np.einsum and ``@`` give
    three different float32 results on the same input), so the reference itself is only defined to 1 ulp.  Canonical
    restatement: ((r*0.299 + g*0.587) + b*0.114), every operation rounded to float32 -- within 1 ulp of the recorded numpy
    output (tests/test_oracle.py)."""
    o = obs_rgb_u8.astype(f32)
    w = np.array([0.299, 0.587, 0.114], dtype=f32)
    return ((o[..., 0] * w[0] + o[..., 1] * w[1]).astype(f32) + o[..., 2] * w[2]).astype(f32)


def pillow_bilinear_coeffs(in_size: int, out_size: int):
    """Pillow ``precompute_coeffs`` (src/libImaging/Resample.c) for the BILINEAR (triangle, support 1) filter: when
    shrinking, the support is scaled by in/out (an area-weighted triangle, NOT 2-tap bilinear).  Returns
    (xmin[out], count[out], k[out, ksize]) in float64, exactly the double arithmetic of the C code."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    xmins, counts, kk = np.zeros(out_size, np.int32), np.zeros(out_size, np.int32), np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            a = -a if a < 0.0 else a
            w = 1.0 - a if a < 1.0 else 0.0
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        xmins[xx], counts[xx] = xmin, xmax
    return xmins, counts, kk


def resize_pillow_bilinear(gray: np.ndarray, res: int = 84) -> np.ndarray:
    """``Image.fromarray(frame).resize((res, res), BILINEAR)`` on a mode-'F' image followed by ``np.array(..., dtype=np.uint8)``
    (atari_wrappers.py:140-141).  Pillow's 32-bit-float path: horizontal pass first, then vertical, each output =
    float32(sum_x double(pixel) * k[x]) accumulated in double (ImagingResampleHorizontal_32bpc / Vertical_32bpc); the uint8
    cast truncates.  Pinned bit-exactly against Pillow itself (tests/test_oracle.py)."""
    gray = np.asarray(gray, dtype=f32)
    H, Wd = gray.shape
    xm, xc, xk = pillow_bilinear_coeffs(Wd, res)
    tmp = np.zeros((H, res), dtype=f32)
    for xx in range(res):
        acc = np.zeros(H, dtype=np.float64)
        for x in range(xc[xx]):
            acc = acc + gray[:, xm[xx] + x].astype(np.float64) * xk[xx, x]
        tmp[:, xx] = acc.astype(f32)
    ym, yc, yk = pillow_bilinear_coeffs(H, res)
    out = np.zeros((res, res), dtype=f32)
    for yy in range(res):
        acc = np.zeros(res, dtype=np.float64)
        for y in range(yc[yy]):
            acc = acc + tmp[ym[yy] + y].astype(np.float64) * yk[yy, y]
        out[yy] = acc.astype(f32)
    return out.astype(np.uint8)                                   # values are in [0, 255]: truncation toward zero


def warp_frame_cpu(obs_rgb_u8: np.ndarray, res: int = 84) -> np.ndarray:
    """atari_wrappers.py:138-142 ``WarpFrame._observation``: gray (canonical float32 formula) -> Pillow BILINEAR -> uint8."""
    return resize_pillow_bilinear(gray_rgb(obs_rgb_u8), res)


def warp_frame_gpu(pal_idx_2: np.ndarray, gray_palette: np.ndarray, res: int = 84) -> np.ndarray:
    """gpu_implementation/gym_tensorflow/atari/tf_atari.py:88-92: palette index -> gray float32 LUT (``:149``), max over
    the two raw frames, ``tf.image.resize_bilinear(..., align_corners=True)`` -> float32 [84, 84] in [0, 1].
    TF kernel arithmetic (resize_bilinear_op.cc): scale = (in-1)/(out-1) float32; in = i*scale; lower = int(in);
    upper = min(lower+1, in_size-1); lerp = in - lower; top = tl + (tr-tl)*xl; bottom = bl + (br-bl)*xl;
    out = top + (bottom-top)*yl, all float32.  TensorFlow is absent here: parity unpinned, by formula."""
    g = gray_palette.astype(f32).reshape(-1)[pal_idx_2.astype(np.int64)]        # [2, 210, 160]
    img = np.maximum(g[0], g[1])
    H, Wd = img.shape

    def weights(n_in, n_out):
        scale = f32((n_in - 1) / f32(n_out - 1)) if n_out > 1 else f32(0)
        pos = (np.arange(n_out, dtype=f32) * scale).astype(f32)
        lo = pos.astype(np.int64)
        hi = np.minimum(lo + 1, n_in - 1)
        return lo, hi, (pos - lo.astype(f32)).astype(f32)
    ylo, yhi, yl = weights(H, res)
    xlo, xhi, xl = weights(Wd, res)
    tl, tr = img[ylo][:, xlo], img[ylo][:, xhi]
    bl, br = img[yhi][:, xlo], img[yhi][:, xhi]
    top = (tl + ((tr - tl).astype(f32) * xl[None, :]).astype(f32)).astype(f32)
    bot = (bl + ((br - bl).astype(f32) * xl[None, :]).astype(f32)).astype(f32)
    return (top + ((bot - top).astype(f32) * yl[:, None]).astype(f32)).astype(f32)


def ntsc_gray_palette() -> np.ndarray:
    """tf_atari.py:100-149: NTSC palette (128 colours at even indices) -> RGB/255 -> gray float32 [256].  The table
    itself is data of the emulator front end; the synthetic pipeline here uses a deterministic stand-in with the same
    structure (even entries populated, odd entries zero) because only the LUT-gather semantics are on the path."""
    rs = np.random.RandomState(1977)
    rgb = np.zeros((256, 3), dtype=np.uint8)
    rgb[0::2] = rs.randint(0, 256, size=(128, 3))
    return (rgb.astype(f32) * f32(1.0 / 255.0)) @ np.array([0.299, 0.587, 0.114], dtype=f32)


# --------------------------------------------------------------------------------------------------
# a-7  rollout accounting                                    es.py:423-426 ; policies.py:378-429
# --------------------------------------------------------------------------------------------------

def episode_accounting(rewards: np.ndarray):
    """es.py:423-426: (sum of rewards f32, sum of sign(rewards) f32, length)."""
    r = np.asarray(rewards, dtype=f32)
    return f32(r.sum()), f32(np.sign(r).sum()), int(len(r))


# --------------------------------------------------------------------------------------------------
# whole-generation ES update                                 es.py:273-301
# --------------------------------------------------------------------------------------------------

def es_generation_update(theta, optimizer, noise, noise_inds_n, returns_n2, l2coeff, *, dtype=np.float64,
                         signreturns_n2=None, return_proc_mode="centered_rank"):
    """es.py:281-301.  Returns (g, update_ratio, new_theta)."""
    if return_proc_mode == "centered_rank":
        proc = compute_centered_ranks(returns_n2)
    elif return_proc_mode == "sign":
        proc = signreturns_n2
    elif return_proc_mode == "centered_sign_rank":
        proc = compute_centered_ranks(signreturns_n2)
    else:
        raise NotImplementedError(return_proc_mode)
    g = es_gradient(proc, noise, noise_inds_n, theta.size, dtype=dtype).astype(f32)
    ratio, new_theta = optimizer.update(es_update_direction(g, theta, l2coeff))
    return g, ratio, new_theta
