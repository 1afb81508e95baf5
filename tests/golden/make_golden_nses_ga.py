"""Generate tests/golden/ref_nses_ga.npz from the REFERENCE's own novelty / selection code and library calls.

Build container only (needs /root/reference and Pillow; nothing here runs on the GPU box):

    python tests/golden/make_golden_nses_ga.py

* es_distributed/nses.py:12-32 (`euclidean_distance`, `compute_novelty_vs_archive`) is imported from /root/reference with
  stub `redis` and `tensorflow` modules (nses.py:4 imports tensorflow only for the session set-up) and
  `np.float = float` (the alias the reference uses at nses.py:24,26 was removed from numpy 1.24+).
* es_distributed/ga.py:145-149: the truncation selection is the literal numpy expression of the reference
  (`np.argpartition(returns, (-population_size, -1))[-1:-population_size-1:-1]`) on tie-free fitness.
* es_distributed/atari_wrappers.py:138-142 (`WarpFrame._observation`): the literal numpy / Pillow expressions on
  seeded RGB frames -- pins the CPU-mode gray + BILINEAR 210x160 -> 84x84 resize of the preprocess kernel.
* gpu_implementation/neuroevolution/display.py:31: the 260-mutation Frostbite genome shipped with the reference,
  parsed from the source file (the only concrete artefact of the GPU path); its theta is materialised by the oracle on
  the REAL 250M-entry noise table and recorded as float64 checksums + sampled coordinates (a GA-materialise KAT).
"""
import ast
import os
import re
import sys
import types

import numpy as np

np.float = float                                   # nses.py:24,26
for name in ("redis", "tensorflow"):
    sys.modules[name] = types.ModuleType(name)
sys.path.insert(0, "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import es_distributed.nses as ref_ns              # noqa: E402

out = {}
rs = np.random.RandomState(20260923)

# --- novelty (nses.py:12-32): ragged uint8 BC sequences, k larger and smaller than the archive ---------------------
D, t_max = 128, 24
def make(nseq):
    lens = rs.randint(1, t_max + 1, size=nseq).astype(np.int32)
    seqs = [rs.randint(0, 256, size=(t, D)).astype(np.uint8) for t in lens]
    pad = np.stack([np.concatenate([s, np.repeat(s[-1:], t_max - len(s), 0)]) for s in seqs])
    return lens, seqs, pad
ql, qs, qp = make(7)
al, as_, ap = make(19)
out["nov_q_len"], out["nov_q_pad"], out["nov_a_len"], out["nov_a_pad"] = ql, qp, al, ap
out["nov_dist"] = np.array([[ref_ns.euclidean_distance(a.astype(float), q.astype(float)) for a in as_] for q in qs])
for k in (1, 10, 19, 40):
    out[f"nov_k{k}"] = np.array([ref_ns.compute_novelty_vs_archive(as_, q, k) for q in qs])
for n_arch in (3,):                                 # archive smaller than k
    out[f"nov_k10_arch{n_arch}"] = np.array([ref_ns.compute_novelty_vs_archive(as_[:n_arch], q, 10) for q in qs])

# --- GA truncation (ga.py:145-149), tie-free fitness -------------------------------------------------------------------
for pop, T in ((1000, 20), (64, 64), (7, 3)):
    fit = (rs.permutation(pop).astype(np.float32) * np.float32(3.7) - np.float32(100.0))
    idx = np.argpartition(fit, (-T, -1))[-1:-T - 1:-1]
    out[f"ga_fit_{pop}_{T}"], out[f"ga_sel_{pop}_{T}"] = fit, idx.astype(np.int64)

# --- CPU-mode preprocess (atari_wrappers.py:138-142) ---------------------------------------------------------------------
from PIL import Image                               # noqa: E402
import PIL                                          # noqa: E402
frames = rs.randint(0, 256, size=(6, 210, 160, 3)).astype(np.uint8)
frames[4] = (np.arange(210)[:, None, None] + np.arange(160)[None, :, None] * 2 + np.arange(3)[None, None, :] * 40) % 256   # smooth
frames[5] = 255
warped, grays = [], []
for obs in frames:
    frame = np.dot(obs.astype('float32'), np.array([0.299, 0.587, 0.114], 'float32'))
    grays.append(frame)
    warped.append(np.array(Image.fromarray(frame).resize((84, 84), resample=Image.BILINEAR), dtype=np.uint8))
out["warp_rgb"], out["warp_gray_f32"], out["warp_out"] = frames, np.stack(grays), np.stack(warped)
out["warp_pillow_version"] = np.array(PIL.__version__)

# --- the reference's Frostbite genome (display.py:31) as a GA-materialise KAT ----------------------------------------------
src = open("/root/reference/gpu_implementation/neuroevolution/display.py").read()
m = re.search(r"^seeds = (\[.*\])\s*$", src, re.M)
genome = ast.literal_eval(m.group(1))
idx0 = int(genome[0])
muts = [(int(i), float(p)) for i, p in genome[1:]]
out["genome_idx0"] = np.int64(idx0)
out["genome_idx"] = np.array([i for i, _ in muts], dtype=np.int64)
out["genome_power"] = np.array([p for _, p in muts], dtype=np.float32)
if os.environ.get("SKIP_GENOME_THETA") != "1":
    from oracle import oracle as O                  # noqa: E402
    sys.path.insert(0, os.path.join(ROOT, "deep-neuroevolution_b200"))
    rsn = np.random.RandomState(123)
    noise = np.empty(250_000_000, dtype=np.float32)
    for s in range(0, len(noise), 1 << 24):          # es.py:60, chunked (the legacy stream is continuous across calls)
        e = min(len(noise), s + (1 << 24))
        noise[s:e] = rsn.randn(e - s)
    net = O.make_net("LargeModel")
    theta = O.ga_materialize_gpu(net, noise, (idx0,) + tuple(muts))
    t64 = theta.astype(np.float64)
    cols = np.random.RandomState(1).randint(0, net.num_params, size=4096)
    cols[:4] = [0, 1, net.num_params - 2, net.num_params - 1]
    out["genome_theta_sum"], out["genome_theta_sumsq"] = np.float64(t64.sum()), np.float64(np.square(t64).sum())
    out["genome_theta_cols"], out["genome_theta_vals"] = cols.astype(np.int64), theta[cols]
    out["genome_noise_checksum"] = np.float64(noise[::1000].astype(np.float64).sum())

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_nses_ga.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: np.asarray(v).shape for k, v in out.items()})
