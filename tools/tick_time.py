"""Dev tool (GPU box): warm average tick time (CUDA events over many ticks) of one LargeModel slot table, launched kernel by
kernel and replayed as a CUDA graph, for a few table sizes.  Not a bench."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-neuroevolution_b200")]
import numpy as np, torch
from dne import _ffi as F, nets
from dne.engine import SlotForward, make_context
from dne.noise import SharedNoiseTable
for kv in filter(None, os.environ.get("DNE_OPTS", "").split(",")):
    k, v = kv.split("="); F.check(F.lib().dne_set_option(k.encode(), int(v)))
count = int(os.environ.get("NOISE_COUNT", 250_000_000))
ctx = make_context(0, SharedNoiseTable(count=count, device="cuda:0"))
net = nets.make_net("LargeModel"); P = net.num_params
rs = np.random.RandomState(0)
theta = torch.from_numpy((rs.randn(P) * 0.05).astype(np.float32)).cuda()
out = {}
for slots in [int(x) for x in os.environ.get("SLOTS_LIST", "124,256").split(",")]:
    pidx = rs.randint(0, count - P + 1, size=slots // 2).astype(np.int64)
    sf = SlotForward(ctx, net, slots)
    sf.set_slots(np.repeat(pidx, 2), np.tile([0.005, -0.005], slots // 2).astype(np.float32))
    pool = torch.randint(0, 256, (4, slots, 84, 84, 4), dtype=torch.uint8, device="cuda")
    def run(n, fn):
        for t in range(8): fn(t)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for t in range(n): fn(t)
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3
    plain = run(400, lambda t: sf.forward(theta, pool[t & 3], paired=True))
    graphs = []
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for r in range(4):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                sf.forward(theta, pool[r], paired=True)
            graphs.append(g)
        graphed = run(400, lambda t: graphs[t & 3].replay())
    out[slots] = dict(plain_us=plain, graph_us=graphed)
    print(slots, json.dumps(out[slots]), flush=True)
