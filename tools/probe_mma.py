"""Dev tool: cycles per tcgen05.mma (tf32, M=128) for three smem operand layouts + numerical check."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-neuroevolution_b200")]
import numpy as np, torch
from dne import _ffi as F
L = F.lib()
L.dne_probe_mma.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p, C.c_void_p]
L.dne_probe_mma.restype = C.c_int
rs = np.random.RandomState(0)
for N in (32, 64, 128):
    A = torch.from_numpy(rs.randn(128, 32).astype(np.float32)).cuda()
    B = torch.from_numpy(rs.randn(N, 32).astype(np.float32)).cuda()
    Cc = torch.zeros(128, N, device="cuda")
    cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
    ref = (A.double() @ B.double().T).cpu().numpy()
    for layout in (0, 1, 2):
        F.check(L.dne_probe_mma(A.data_ptr(), B.data_ptr(), Cc.data_ptr(), N, layout, 1, cyc.data_ptr(), None))
        torch.cuda.synchronize()
        err = np.abs(Cc.cpu().numpy() - ref).max() / np.abs(ref).max()
        res = []
        for reps in (64, 512):
            F.check(L.dne_probe_mma(A.data_ptr(), B.data_ptr(), Cc.data_ptr(), N, layout, reps, cyc.data_ptr(), None))
            torch.cuda.synchronize()
            res.append(int(cyc.item()) / (reps * 4))
        print(f"N={N:3d} layout={layout} rel_err(1 rep, tf32-hi)={err:.2e}  cycles/MMA @64 reps={res[0]:.1f} @512 reps={res[1]:.1f}")
