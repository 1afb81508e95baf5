"""Dev tool (GPU box): per-kernel CUDA-event timing of one forward tick + update kernels.  Not a bench."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-neuroevolution_b200")]
import numpy as np, torch
from dne import nets
from dne.engine import SlotForward, ESUpdate, make_context
from dne.noise import SharedNoiseTable, generate_host

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

count = int(os.environ.get("NOISE_COUNT", 250_000_000))
t0 = time.time()
host = np.random.RandomState(123).randn(count).astype(np.float32) if count <= 30_000_000 else generate_host(count)
print("noise gen s", time.time() - t0, flush=True)
table = SharedNoiseTable(host_noise=host, device="cuda:0")
ctx = make_context(0, table)
out = {}
for name, slots in (("LargeModel", 256), ("Model", 256), ("MujocoPolicy", 10000)):
    net = nets.make_net(name)
    P = net.num_params
    rs = np.random.RandomState(0)
    theta = torch.from_numpy((rs.randn(P) * 0.05).astype(np.float32)).cuda()
    pidx = rs.randint(0, count - P + 1, size=slots // 2).astype(np.int64)
    idx, scale = np.repeat(pidx, 2), np.tile([0.02, -0.02], slots // 2).astype(np.float32)
    sf = SlotForward(ctx, net, slots)
    sf.set_slots(idx, scale)
    if net.ob_kind == 0:
        obs = torch.randint(0, 256, (slots, 84, 84, 4), dtype=torch.uint8, device="cuda")
        kw = {}
    else:
        obs = torch.randn(slots, 376, device="cuda")
        kw = dict(ob_mean=torch.zeros(376, device="cuda"), ob_std=torch.ones(376, device="cuda"))
    ms_p = timeit(lambda: sf.forward(theta, obs, paired=True, **kw))
    ms_u = timeit(lambda: sf.forward(theta, obs, paired=False, **kw))
    bytes_step = 4 * P
    out[name] = dict(slots=slots, ms_paired=ms_p, ms_unpaired=ms_u,
                     steps_per_s_paired=slots / ms_p * 1e3, noise_GBs_paired=slots / 2 * bytes_step / ms_p / 1e6,
                     noise_GBs_unpaired=slots * bytes_step / ms_u / 1e6)
    print(name, json.dumps(out[name]), flush=True)
    n = 500 if name != "MujocoPolicy" else 5000
    upd = ESUpdate(ctx, theta, "adam", stepsize=0.01)
    gi = torch.from_numpy(rs.randint(0, count - P + 1, size=n).astype(np.int64)).cuda()
    proc = torch.randn(n, 2, device="cuda")
    ms_g = timeit(lambda: upd.gradient(proc, gi, 2 * n), n=5, warm=1)
    ms_a = timeit(lambda: upd.step(0.005), n=20)
    ret = torch.randn(n, 2, device="cuda")
    ms_r = timeit(lambda: upd.centered_ranks(ret), n=20)
    out[name + "_update"] = dict(n=n, grad_ms=ms_g, grad_GBs=n * 4 * P / ms_g / 1e6, adam_ms=ms_a,
                                 adam_GBs=7 * 4 * P / ms_a / 1e6, rank_ms=ms_r)
    print(name, "update", json.dumps(out[name + "_update"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kernel_times.json"), "w"), indent=1)
