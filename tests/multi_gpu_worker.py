"""Helper of tests/test_gpu_multi.py: one rank of a (torchrun) job that runs a driver for a few generations on the
deterministic environment and, on rank 0, writes what the generations produced.  Usage:
    python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_worker.py <es|ga|nsr> <out.npz>"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "deep-neuroevolution_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np      # noqa: E402
import torch            # noqa: E402

from dne import shard                        # noqa: E402
from dne.envs import DeterministicAtariEnv    # noqa: E402
from dne.noise import SharedNoiseTable        # noqa: E402

CFG = {"calc_obstat_prob": 0.0, "episodes_per_batch": 24, "eval_prob": 0.0, "l2coeff": 0.005, "noise_stdev": 0.02,
       "snapshot_freq": 0, "timesteps_per_batch": 10, "return_proc_mode": "centered_rank", "episode_cutoff_mode": 5000}
EXPS = {
    "es": {"config": CFG, "env_id": "SyntheticAtariDeterministic", "optimizer": {"args": {"stepsize": 0.01}, "type": "adam"},
           "policy": {"args": {}, "type": "GAAtariPolicy"}},
    "ga": {"config": dict(CFG, episodes_per_batch=16), "env_id": "SyntheticAtariDeterministic", "population_size": 4,
           "num_elites": 1, "policy": {"args": {"nonlin_type": "relu"}, "type": "GAAtariPolicy"}},
    "nsr": {"config": dict(CFG, episodes_per_batch=12, return_proc_mode="centered_sign_rank"),
            "env_id": "SyntheticAtariDeterministic", "algo_type": "nsr",
            "novelty_search": {"k": 3, "population_size": 2, "num_rollouts": 1, "selection_method": "novelty_prob"},
            "optimizer": {"args": {"stepsize": 0.01}, "type": "adam"}, "policy": {"args": {}, "type": "GAAtariPolicy"}},
}


def main():
    algo, out = sys.argv[1], sys.argv[2]
    rank, world, local = shard.init_from_env("nccl")
    torch.cuda.set_device(local)
    noise = SharedNoiseTable(count=6_000_000, device=torch.device("cuda", local))
    env = DeterministicAtariEnv(8, episode_len=6, seed=3)
    exp = json.loads(json.dumps(EXPS[algo]))
    rec = {}

    def on_it(it, stats, extra):
        if algo == "es":
            rec[f"returns_{it}"], rec[f"idx_{it}"] = extra["returns_n2"].copy(), extra["noise_inds_n"].copy()
            rec[f"theta_{it}"], rec[f"g_{it}"] = extra["theta"].cpu().numpy(), extra["g"].cpu().numpy()
        elif algo == "ga":
            rec[f"returns_{it}"] = np.asarray(extra["returns"]).copy()
            rec[f"score_{it}"] = np.asarray(extra["population_score"]).copy()
            rec[f"pop_{it}"] = np.array([json.dumps([int(x) for x in g]) for g in extra["population"]])
            rec[f"elite_{it}"] = extra["elite_theta"].cpu().numpy()
        else:
            rec[f"returns_{it}"], rec[f"novelty_{it}"] = extra["returns_n2"].copy(), extra["novelty_n2"].copy()
            rec[f"theta_{it}"], rec[f"parent_{it}"] = extra["theta"].cpu().numpy(), np.int64(extra["parent"])
            rec[f"g_{it}"] = extra["g"].cpu().numpy()
            rec[f"archive_len_{it}"] = np.int64(len(extra["archive"]))
            rec[f"archive_last_{it}"] = extra["archive"].seqs[-1].copy()
    if algo == "es":
        from es_distributed import es as D
    elif algo == "ga":
        from es_distributed import ga as D
    else:
        from es_distributed import nses as D
    D.run_master(None, None, exp, max_iterations=2, n_slots=8, env=env, noise=noise, seed=7, on_iteration=on_it)
    if rank == 0:
        np.savez(out, world=np.int64(world), **rec)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
