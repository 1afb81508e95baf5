"""Dev tool: cycles per tcgen05.mma (tf32, M=128) for three smem operand layouts + numerical check."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-neuroevolution_b200")]
import numpy as np, torch
from dne import _ffi as F
L = F.dev_lib()
rs = np.random.RandomState(0)
cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
names = {0: "chain (4 MMAs/iter)", 1: "6 MMAs + commit + wait / iter", 2: "chain + generic ST traffic (3 warps)",
         3: "chain + st.shared traffic (3 warps)", 4: "ping-pong: stage warp -> full -> 6 MMAs -> commit(empty) / iter"}
def run(N, mode, layout, reps):
    A = torch.from_numpy(rs.randn(128, 32).astype(np.float32)).cuda()
    B = torch.from_numpy(rs.randn(N, 32).astype(np.float32)).cuda()
    Cc = torch.zeros(128, N, device="cuda")
    F.check(L.dne_probe_mma(A.data_ptr(), B.data_ptr(), Cc.data_ptr(), N, (mode << 4) | layout, reps, cyc.data_ptr(), None))
    torch.cuda.synchronize()
    return int(cyc.item()) / reps
for N in (32, 64, 128):
    for layout in (0, 1, 2):
        print(f"N={N} layout={layout} chain: {run(N, 0, layout, 512) / 4:.1f} cycles/MMA", flush=True)
for mode in (1, 2, 3, 4):
    print(f"N=64 mode {mode} [{names[mode]}]: {run(64, mode, 0, 512):.1f} cycles/iter", flush=True)
