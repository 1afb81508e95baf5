"""Dev tool (GPU box): a few forward ticks + one update for ncu captures.  Not a bench."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-neuroevolution_b200")]
import numpy as np, torch
from dne import nets
from dne.engine import SlotForward, ESUpdate, make_context
from dne.noise import SharedNoiseTable
name = os.environ.get("NET", "LargeModel"); slots = int(os.environ.get("SLOTS", 256)); ticks = int(os.environ.get("TICKS", 4))
count = int(os.environ.get("NOISE_COUNT", 60_000_000))
host = np.random.RandomState(123).randn(count).astype(np.float32)
ctx = make_context(0, SharedNoiseTable(host_noise=host, device="cuda:0"))
from dne import _ffi as F
for kv in filter(None, os.environ.get("DNE_OPTS", "").split(",")):
    k, v = kv.split("="); F.check(F.lib().dne_set_option(k.encode(), int(v)))
net = nets.make_net(name); P = net.num_params
rs = np.random.RandomState(0)
theta = torch.from_numpy((rs.randn(P) * 0.05).astype(np.float32)).cuda()
pidx = rs.randint(0, count - P + 1, size=slots // 2).astype(np.int64)
sf = SlotForward(ctx, net, slots)
sf.set_slots(np.repeat(pidx, 2), np.tile([0.02, -0.02], slots // 2).astype(np.float32))
if net.ob_kind == 0:
    obs = torch.randint(0, 256, (slots, 84, 84, 4), dtype=torch.uint8, device="cuda"); kw = {}
else:
    obs = torch.randn(slots, 376, device="cuda"); kw = dict(ob_mean=torch.zeros(376, device="cuda"), ob_std=torch.ones(376, device="cuda"))
for _ in range(ticks):
    sf.forward(theta, obs, paired=True, **kw)
torch.cuda.synchronize()
if os.environ.get("UPDATE", "1") == "1":
    n = 500
    upd = ESUpdate(ctx, theta, "adam", stepsize=0.01)
    gi = torch.from_numpy(rs.randint(0, count - P + 1, size=n).astype(np.int64)).cuda()
    cen, _ = upd.centered_ranks(torch.randn(n, 2, device="cuda"))
    upd.gradient(cen, gi, 2 * n); upd.step(0.005)
    torch.cuda.synchronize()
print("done")
