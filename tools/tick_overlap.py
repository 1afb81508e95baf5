"""Dev tool (GPU box): steady-state time of one ROUND (every slot of the GPU ticks once) when the slot table is split into
NT tables on NT streams (CUDA-graph replays, no cross-stream ordering), against the single-table tick.  The HBM-bound
noise GEMV of one table could overlap the tensor-core kernels of another if the GEMV left whole SMs free
(OPT_SETS="a=1,b=2;a=1,b=3" loops over option sets, e.g. gemv_ctas_per_sm=1,gemv_stages=8,gemv_grid=G; SLOTS_LIST, TABLES).
Result: it does not pay -- the GEMV needs 2 CTAs/SM on all SMs.  Not a bench."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-neuroevolution_b200")]
import numpy as np, torch
from dne import _ffi as F, nets
from dne.engine import SlotForward, make_context
from dne.noise import SharedNoiseTable
def set_opts(spec):
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("="); F.check(F.lib().dne_set_option(k.encode(), int(v)))
count = int(os.environ.get("NOISE_COUNT", 250_000_000))
ctx = make_context(0, SharedNoiseTable(count=count, device="cuda:0"))
net = nets.make_net("LargeModel"); P = net.num_params
rs = np.random.RandomState(0)
theta = torch.from_numpy((rs.randn(P) * 0.05).astype(np.float32)).cuda()
N = int(os.environ.get("TICKS", 300))
for OPTS in os.environ.get("OPT_SETS", "").split(";"):
  set_opts(OPTS)
  for total in [int(x) for x in os.environ.get("SLOTS_LIST", "124,256").split(",")]:
      for NT in [int(x) for x in os.environ.get("TABLES", "1,2").split(",")]:
          part = 2 * (-(-(total // 2) // NT))
          tabs = []
          for h in range(NT):
              pidx = rs.randint(0, count - P + 1, size=part // 2).astype(np.int64)
              sf = SlotForward(ctx, net, part)
              sf.set_slots(np.repeat(pidx, 2), np.tile([0.005, -0.005], part // 2).astype(np.float32))
              pool = torch.randint(0, 256, (4, part, 84, 84, 4), dtype=torch.uint8, device="cuda")
              st = torch.cuda.Stream()
              gs = []
              with torch.cuda.stream(st):
                  sf.forward(theta, pool[0], paired=True)
                  torch.cuda.synchronize()
                  for r in range(4):
                      g = torch.cuda.CUDAGraph()
                      with torch.cuda.graph(g, stream=st):
                          sf.forward(theta, pool[r], paired=True)
                      gs.append(g)
              tabs.append((sf, pool, st, gs))
          def rounds(n):
              for t in range(n):
                  for (sf, pool, st, gs) in tabs:
                      with torch.cuda.stream(st):
                          gs[t & 3].replay()
          rounds(16)
          torch.cuda.synchronize()
          a, b = torch.cuda.Event(True), torch.cuda.Event(True)
          a.record()
          for (_, _, st, _) in tabs:
              st.wait_event(a)
          rounds(N)
          cur = torch.cuda.current_stream()
          for (_, _, st, _) in tabs:
              cur.wait_stream(st)
          b.record(); torch.cuda.synchronize()
          us = a.elapsed_time(b) / N * 1e3
          print(json.dumps(dict(slots=part * NT, tables=NT, round_us=round(us, 2), us_per_slot=round(us / (part * NT), 4),
                                opts=OPTS)), flush=True)
          del tabs
