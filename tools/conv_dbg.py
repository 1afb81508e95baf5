"""Dev tool: time the conv_tc kernels with parts disabled (dbg mask) to find the bottleneck."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-neuroevolution_b200")]
import numpy as np, torch
from dne import nets, _ffi as F
from dne.engine import SlotForward, make_context
from dne.noise import SharedNoiseTable
count = 60_000_000
host = np.random.RandomState(123).randn(count).astype(np.float32)
ctx = make_context(0, SharedNoiseTable(host_noise=host, device="cuda:0"))
net = nets.make_net("LargeModel"); P = net.num_params
# conv-only net: time full forward minus known others is messy; instead time whole forward per mask and diff
rs = np.random.RandomState(0)
theta = torch.from_numpy((rs.randn(P) * 0.05).astype(np.float32)).cuda()
slots = 256
pidx = rs.randint(0, count - P + 1, size=slots // 2).astype(np.int64)
sf = SlotForward(ctx, net, slots)
sf.set_slots(np.repeat(pidx, 2), np.tile([0.02, -0.02], slots // 2).astype(np.float32))
obs = torch.randint(0, 256, (slots, 84, 84, 4), dtype=torch.uint8, device="cuda")
def t(mask, n=20):
    F.check(F.lib().dne_set_option(b"dbg", mask))
    for _ in range(3): sf.forward(theta, obs, paired=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): sf.forward(theta, obs, paired=True)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1000
base = t(0)
print("full tick us", round(base, 1))
for name, m in (("no MMA", 1), ("no A loads", 2), ("no B loads", 4), ("no epilogue stores", 8), ("no A, no B", 6), ("no MMA/A/B", 7), ("nothing but skeleton", 15)):
    v = t(m); print(f"{name:24s} tick {v:8.1f} us   saved {base - v:7.1f}")
F.lib().dne_set_option(b"dbg", 0)
