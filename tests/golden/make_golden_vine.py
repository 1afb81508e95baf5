"""Generate tests/golden/ref_vine.npz by EXECUTING the reference's VINE export functions.

Run in the build container only (imports /root/reference; nothing under tests/ reads that path at test time):
    python tests/golden/make_golden_vine.py

es_distributed/es_modified.py `master_extract_cloud` (:179-199) and `master_extract_parent` (:140-177) write the per-generation
behaviour-characterisation files the reference's visual_inspector reads.  They are plain numpy + csv; the module only needs a
stand-in for `redis` to import.  The fixture holds the exact bytes they write for a small seeded input."""
import importlib
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def inputs():
    rs = np.random.RandomState(31)
    Result = types.SimpleNamespace
    results, cloud = [], []
    for w in range(3):                                          # three worker results, two episodes (+/-) each
        pts = []
        for sign in (1, -1):
            bc = rs.randint(0, 256, size=(int(rs.randint(2, 6)), 128)).astype(np.float64)      # RAM trace [t, 128]
            pt = (bc, float(rs.randint(0, 500) * 10), int(rs.randint(50, 900)), int(rs.randint(0, 2 ** 31)), int(rs.randint(0, 1000)), sign)
            pts.append(pt)
            cloud.append(pt)
        results.append(Result(bc_vectors=pts))
    evals = [(rs.randint(0, 256, size=(3, 128)).astype(np.float64), float(r), int(rs.randint(50, 900)), int(rs.randint(0, 1000)), 0.02)
             for r in (120.0, 80.0, 310.0, 150.0)]
    return results, cloud, evals, [e[1] for e in evals]


def main():
    sys.modules["redis"] = types.ModuleType("redis")
    pkg = types.ModuleType("refes")
    pkg.__path__ = ["/root/reference/es_distributed"]
    sys.modules["refes"] = pkg
    M = importlib.import_module("refes.es_modified")
    results, cloud, evals, rets = inputs()
    pol = types.SimpleNamespace(save=lambda fn: open(fn, "wb").close())
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            M.master_extract_cloud(results, 7)
            M.master_extract_parent(evals, rets, 7, pol, np.zeros((2, 3)))
            base = os.path.join(d, "snapshots", "snapshot_gen_0007")
            files = sorted(os.listdir(base))
            off = open(os.path.join(base, "snapshot_offspring_0007.dat"), "rb").read()
            par = open(os.path.join(base, "snapshot_parent_0007.dat"), "rb").read()
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "ref_vine.npz"), files=np.array(files), offspring=np.frombuffer(off, dtype=np.uint8),
                        parent=np.frombuffer(par, dtype=np.uint8))
    print(files, len(off), len(par))
    print(par[:120])


if __name__ == "__main__":
    main()
