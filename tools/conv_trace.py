"""Dev tool: clock64 timeline of one conv2 CTA (library built with -DDNE_CONV_TRACE) + tick time vs slot count."""
import os, sys, ctypes as C
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
rs = np.random.RandomState(0)
theta = torch.from_numpy((rs.randn(P) * 0.05).astype(np.float32)).cuda()
L = F.lib()
L.dne_debug_conv_trace.argtypes = [C.c_void_p]
buf = (C.c_longlong * 1024)()
for slots in (128, 148, 256, 296, 512):
    pidx = rs.randint(0, count - P + 1, size=slots // 2).astype(np.int64)
    sf = SlotForward(ctx, net, slots)
    sf.set_slots(np.repeat(pidx, 2), np.tile([0.02, -0.02], slots // 2).astype(np.float32))
    obs = torch.randint(0, 256, (slots, 84, 84, 4), dtype=torch.uint8, device="cuda")
    for _ in range(200): sf.forward(theta, obs, paired=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(100): sf.forward(theta, obs, paired=True)
    b.record(); torch.cuda.synchronize()
    assert L.dne_debug_conv_trace(buf) == 0
    t = list(buf)
    cyc, ns = t[3] - t[0], t[5] - t[4]
    print(f"slots={slots}: tick {a.elapsed_time(b) * 10:.1f} us; traced conv2 CTA: {cyc} cycles in {ns} ns -> {cyc / max(ns, 1):.2f} GHz; "
          f"loop start {t[128] - t[0]}, mma_done {t[1] - t[0]}", flush=True)
    if slots == 296:
        r = lambda x: x - t[0]
        print("  MMA warp (c: full_seen, issued): " + " ".join(f"{c}:{r(t[16+2*c])},{r(t[17+2*c])}" for c in range(32)))
        for g in range(4):
            print(f"  g={g}: " + " | ".join(" ".join(str(r(t[128 + (g * 16 + it) * 4 + j])) for j in range(4)) for it in range(8)))
    del sf, obs
