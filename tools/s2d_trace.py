"""Dev tool (GPU box): per-chunk timeline of CTA 0 of the conv_s2d kernels (build with -DDNE_S2D_TRACE into
dne/libdne_trace.so; run with DNE_LIB=.../libdne_trace.so).  Prints clock64 deltas relative to kernel start."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-neuroevolution_b200")]
# build it first: make -C deep-neuroevolution_b200/csrc trace
os.environ.setdefault("DNE_LIB", os.path.join(ROOT, "deep-neuroevolution_b200", "dne", "libdne_trace.so"))
import numpy as np, torch
from dne import _ffi as F, nets
from dne.engine import SlotForward, make_context
from dne.noise import SharedNoiseTable
count = int(os.environ.get("NOISE_COUNT", 250_000_000)); slots = int(os.environ.get("SLOTS", 256))
host = np.random.RandomState(123).randn(count).astype(np.float32) if count <= 60_000_000 else None
ctx = make_context(0, SharedNoiseTable(host_noise=host, device="cuda:0") if host is not None else SharedNoiseTable(count=count, device="cuda:0"))
net = nets.make_net("LargeModel"); P = net.num_params
rs = np.random.RandomState(0)
theta = torch.from_numpy((rs.randn(P) * 0.05).astype(np.float32)).cuda()
pidx = rs.randint(0, count - P + 1, size=slots // 2).astype(np.int64)
sf = SlotForward(ctx, net, slots)
sf.set_slots(np.repeat(pidx, 2), np.tile([0.005, -0.005], slots // 2).astype(np.float32))
obs = torch.randint(0, 256, (slots, 84, 84, 4), dtype=torch.uint8, device="cuda")
for _ in range(3):
    sf.forward(theta, obs, paired=True)
torch.cuda.synchronize()
L = F.lib()
L.dne_debug_s2d_trace.argtypes = [C.c_void_p]
buf = (C.c_longlong * (3 * 512))()
assert L.dne_debug_s2d_trace(buf) == 0
tr = np.array(buf, dtype=np.int64).reshape(3, 512)
NG = [4, 16, 12]   # chunks per member (conv1, conv2, conv3)
names = ["wprod:slot free", "wprod:issued", "conv:raw landed", "conv:tile ready", "mma:A ready", "mma:B ready", "mma:issued", "conv:A staged/begin"]
for l in range(3):
    t = tr[l]; t0 = t[0]
    print(f"=== layer {l}: setup done +{t[1]-t0}, acc_full it0 +{t[2]-t0} it1 +{t[3]-t0}, epi done it0 +{t[4]-t0} it1 +{t[5]-t0}, kernel end +{t[6]-t0}")
    for it in range(2):
        for g in range(NG[l]):
            ev = [int(t[16 + ((it * 16 + g) * 8 + e)] - t0) for e in range(8)]
            print(f"  it{it} g{g:2d}: wfree {ev[0]:7d} wiss {ev[1]:7d} | cbegin {ev[7]:7d} raw {ev[2]:7d} tile {ev[3]:7d} | mmaA {ev[4]:7d} mmaB {ev[5]:7d} iss {ev[6]:7d}")
