"""CPU restatement of the reference WORKER inner loop and MASTER update, used only as the timed CPU baseline
(`bench.py --impl reference`, and the `cpu_baseline` leg) -- TEST/BENCH INFRASTRUCTURE, never the product path.

The real reference workers cannot run here or on the GPU box (TensorFlow 0.12, gym, ALE, redis are absent:
SURVEY.md 8c), so this follows BASELINE.md's plan:
  worker   es_distributed/es.py:411-426 + policies.py:399-409: per antithetic pair, theta +/- sigma*noise[idx] is
           written into the layer tensors (SetFromFlat, tf_util.py:224-240), then one forward + argmax per env step
           at batch 1 with ONE intra-op thread (es.py:91), the environment stubbed (synthetic 84x84x4 uint8
           observation per step, reward 10*Bernoulli(0.05));
  master   es.py:273-301: compute_centered_ranks, batched_weighted_sum (500-row float32 slabs), Adam step.
Workers are forked one per host core and share the noise table by fork, exactly like es_distributed/main.py:79-85.
The per-step forward uses prepared torch-CPU tensors (weights converted once per member, as TF assigns variables
once per set_trainable_flat) so the baseline is not handicapped by Python conversions; it is checked against
oracle.forward in tests/test_oracle.py.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from typing import List

import numpy as np

from . import oracle as O

_G = {}   # fork-shared state (noise table, theta, net)


def prepare(net: O.Net, theta: np.ndarray):
    """SetFromFlat: slice the flat vector into layer tensors (torch, conv kernels OIHW)."""
    import torch
    prep = []
    for l, p in zip(net.layers, O.unflatten(net, theta)):
        d = {}
        if l.kind == "conv":
            d["w"] = torch.from_numpy(np.ascontiguousarray(p["w"])).permute(3, 2, 0, 1).contiguous()
        else:
            d["w"] = torch.from_numpy(np.ascontiguousarray(p["w"]))
        for k in ("b", "beta", "gamma"):
            if k in p:
                d[k] = torch.from_numpy(np.ascontiguousarray(p[k]))
        prep.append(d)
    return prep


def forward_prepared(net: O.Net, prep, obs_u8: np.ndarray, vbn_stats=None, is_ref=False):
    """Same arithmetic as oracle.forward for Atari nets, torch end to end.  Returns (logits [B,A], stats)."""
    import torch
    import torch.nn.functional as F
    x = torch.from_numpy(obs_u8).to(torch.float32).div_(255.0).permute(0, 3, 1, 2)      # NCHW
    stats_out, bn_i = [], 0
    for l, d in zip(net.layers, prep):
        if l.kind == "conv":
            total = max((l.hout - 1) * l.stride + l.ksize - l.hin, 0)
            pb, pa = total // 2, total - total // 2
            y = F.conv2d(F.pad(x, (pb, pa, pb, pa)), d["w"], d.get("b"), stride=l.stride)
            cdim = 1
        else:
            if x.dim() == 4:
                x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)                           # flatten (h,w,c)
            y = torch.addmm(d["b"], x, d["w"]) if "b" in d else x @ d["w"]
            cdim = 1
        if l.bn == "tf":
            dims = [i for i in range(y.dim()) if i != cdim]
            shape = [1, -1] + [1] * (y.dim() - 2)
            if is_ref:
                mean = y.mean(dim=dims)
                var = (y - mean.view(shape)).square().mean(dim=dims)
                stats_out.append((mean, var))
            else:
                mean, var = vbn_stats[bn_i]
            bn_i += 1
            y = (y - mean.view(shape)) * torch.rsqrt(var.view(shape) + O.BN_EPS) * d["gamma"].view(shape) + d["beta"].view(shape)
        x = torch.relu(y) if l.act == "relu" else (torch.tanh(y) if l.act == "tanh" else y)
    return x, stats_out


def _worker(args):
    """One reference worker process: evaluates `pairs` antithetic pairs for `T` steps each."""
    import torch
    torch.set_num_threads(1)                                    # es.py:91 / scripts/launch.py:117
    wid, pair_idx, T, sigma, seed = args
    cores = _G.get("cores")
    if cores:                                                   # one worker per core: no migration, and the member weights the
        try:                                                    # worker allocates below are first-touched on its own NUMA node
            os.sched_setaffinity(0, {cores[wid % len(cores)]})
        except (AttributeError, OSError):
            pass
    net, noise, theta = _G["net"], _G["noise"], _G["theta"]
    rs = np.random.RandomState(seed + wid)
    obs_pool = rs.randint(0, 256, size=(16, 1, 84, 84, 4)).astype(np.uint8)
    ref = _G.get("ref_batch")
    steps, out = 0, []
    t_setup = t_steps = 0.0
    for idx in pair_idx:
        ta = time.perf_counter()
        v = np.float32(sigma) * noise[idx:idx + net.num_params]              # es.py:413
        t_setup += time.perf_counter() - ta
        rets = []
        for sign in (+1, -1):
            ta = time.perf_counter()
            prep = prepare(net, (theta + v) if sign > 0 else (theta - v))   # es.py:415,419
            stats = None
            if ref is not None:
                _, stats = forward_prepared(net, prep, ref, is_ref=True)     # policies.py:399
            tb = time.perf_counter()
            t_setup += tb - ta
            ret = 0.0
            for t in range(T):                                               # policies.py:401-424
                logits, _ = forward_prepared(net, prep, obs_pool[t & 15], vbn_stats=stats)
                _ = int(torch.argmax(logits, dim=1)[0])
                ret += 10.0 * float(rs.random_sample() < 0.05)               # stub env.step
                steps += 1
            t_steps += time.perf_counter() - tb
            rets.append(ret)
        out.append(rets)
    return steps, t_setup, t_steps, out


def measure_workers(net_name: str, noise: np.ndarray, theta: np.ndarray, idx: List[int], T: int, sigma: float,
                    n_workers: int, use_ref_batch: bool = False, seed: int = 0):
    """Fork `n_workers` processes (noise shared by fork), split `idx` among them.  Returns
    (total env steps, wall seconds, per-episode set-up seconds [set_trainable_flat (+ VBN pass)], per-env-step
    seconds) -- the last two are busy-time averages over all workers, used to extrapolate the bounded sample to
    full-length episodes (set-up is paid once per episode)."""
    import torch                      # import in the parent so the forked workers do not each pay the cold import
    torch.set_num_threads(1)
    net = O.make_net(net_name)
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cores = None
    _G.update(net=net, noise=noise, theta=theta, cores=cores)
    if use_ref_batch:
        _G["ref_batch"] = np.random.RandomState(seed).randint(0, 256, size=(128, 84, 84, 4)).astype(np.uint8)
    else:
        _G.pop("ref_batch", None)
    chunks = [list(idx[w::n_workers]) for w in range(n_workers)]
    chunks = [c for c in chunks if c]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(chunks)) as pool:
        res = pool.map(_worker, [(w, c, T, sigma, seed) for w, c in enumerate(chunks)])
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    episodes = 2 * sum(len(c) for c in chunks)
    return steps, wall, sum(r[1] for r in res) / episodes, sum(r[2] for r in res) / max(steps, 1)


def measure_master_update(noise: np.ndarray, theta: np.ndarray, idx: np.ndarray, returns_n2: np.ndarray,
                          l2coeff: float = 0.005, stepsize: float = 0.01):
    """es.py:273-301 single process: ranks -> batched_weighted_sum (float32, 500-row slabs) -> Adam.
    Returns (seconds, g)."""
    t0 = time.perf_counter()
    proc = O.compute_centered_ranks(returns_n2)
    g = O.es_gradient(proc, noise, idx, theta.size, dtype=np.float32)
    opt = O.Adam(theta, stepsize)
    opt.update(O.es_update_direction(g, theta, l2coeff))
    return time.perf_counter() - t0, g


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
