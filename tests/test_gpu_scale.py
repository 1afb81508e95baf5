"""GPU parity at the BENCHMARKED scale (BASELINE.json configs[1]): LargeModel, 256 env slots per GPU = 125 active
antithetic pairs + an inactive tail (the last wave of pop 1000), the 124-slot shard of the 8-GPU run, the multi-table
phase-event schedule, and the ESAtariPolicy virtual-batch-norm pass at the configuration's n_ref = 128.

What these cover that the small-slot tests cannot: every persistent CTA of the noise GEMV walks
several (pair, K-chunk) work items (n_items = 128 groups x 32 chunks = 4096 >> the 296-CTA grid) with inactive items
skipped inside the loop; the conv / theta-GEMM grids run more than one wave; split-K of the shared-theta GEMM at
M = 256.  Oracle on a sample of slots (the oracle is a CPU torch forward: ~20 ms per slot), SIMT-vs-tensor-core cross
check on ALL slots.  Tolerance: |dlogit|_inf <= 2e-5 * max(1, |logit|_inf) per row (tests/test_gpu_parity.py)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import oracle as O            # noqa: E402  (checker only)
from dne import _ffi as F                 # noqa: E402
from dne import nets as N                 # noqa: E402
from dne.engine import SlotForward, make_context   # noqa: E402
from dne.noise import SharedNoiseTable    # noqa: E402

DEV = torch.device("cuda", 0)
NOISE_COUNT = 12_000_000
SIGMA = 0.005                             # configurations/frostbite_es.json:7
GEMV_GRID = 2 * 148                       # persistent CTAs of gemv_bulk_kernel (2 per SM)


@pytest.fixture(scope="module")
def host_noise():
    return O.noise_table(NOISE_COUNT)


@pytest.fixture(scope="module")
def ctx(host_noise):
    return make_context(0, SharedNoiseTable(host_noise=host_noise, device=DEV))


def cuda(x):
    return torch.as_tensor(np.ascontiguousarray(x)).to(DEV).contiguous()


def _row_bound(ref, tol=2e-5):
    return tol * np.maximum(1.0, np.abs(ref).max(axis=1))


def _oracle_rows(net_o, theta, host_noise, idx, scale, obs, rows):
    P = net_o.num_params
    return np.stack([O.forward(net_o, (theta + np.float32(scale[s]) * host_noise[idx[s]:idx[s] + P]).astype(np.float32),
                               obs[s:s + 1])[0][0] for s in rows])


def _setup(rs, n_slots, n_active_pairs, P):
    pidx = rs.randint(0, NOISE_COUNT - P + 1, size=n_slots // 2).astype(np.int64)
    pidx[0] = (pidx[0] // 4) * 4                 # every alignment of the slice start among the first pairs
    pidx[1] = (pidx[1] // 4) * 4 + 1
    pidx[2] = (pidx[2] // 4) * 4 + 2
    pidx[3] = (pidx[3] // 4) * 4 + 3
    idx = np.repeat(pidx, 2)
    scale = np.tile([SIGMA, -SIGMA], n_slots // 2).astype(np.float32)
    active = np.zeros(n_slots, dtype=np.uint8)
    active[:2 * n_active_pairs] = 1
    obs = rs.randint(0, 256, size=(n_slots, 84, 84, 4)).astype(np.uint8)
    return idx, scale, active, obs


def _sample_rows(rs, n_active, k=32):
    rows = set(rs.choice(n_active, size=min(k, n_active), replace=False).tolist())
    rows |= {0, 1, 6, 7, n_active - 2, n_active - 1}       # first pairs (all alignments start there), last active pair
    return sorted(rows)


@pytest.mark.parametrize("n_slots,n_pairs", [(256, 125), (256, 128), (124, 62)])
def test_largemodel_benchmark_scale_vs_oracle(ctx, host_noise, n_slots, n_pairs):
    """configs[1] tick: 256 slots with 125 active pairs + 6 inactive tail slots (last wave of pop 1000), the full
    128-pair wave, and the 124-slot shard of the 8-GPU run.  >= 32 sampled slots vs the oracle; inactive untouched."""
    net, net_o = N.make_net("LargeModel"), O.make_net("LargeModel")
    P = net.num_params
    rs = np.random.RandomState(1000 + n_slots + n_pairs)
    theta = (rs.randn(P) * 0.05).astype(np.float32)
    idx, scale, active, obs = _setup(rs, n_slots, n_pairs, P)
    # the persistent noise-GEMV CTAs must each walk several work items for this test to mean anything
    fc = net.layers[3]
    n_chunks = -(-fc.cin // 242)
    assert (n_slots // 2) * n_chunks >= 4 * GEMV_GRID or n_slots < 256
    sf = SlotForward(ctx, net, n_slots)
    sf.set_slots(idx, scale, active=active if n_pairs * 2 < n_slots else None)
    sf.actions.fill_(-7)
    sf.logits.fill_(123.0)
    d_theta, d_obs = cuda(theta), cuda(obs)
    actions = sf.forward(d_theta, d_obs, paired=True).cpu().numpy()
    logits = sf.logits.cpu().numpy()
    n_active = 2 * n_pairs
    assert (actions[n_active:] == -7).all() and (logits[n_active:] == 123.0).all()
    rows = _sample_rows(rs, n_active)
    assert len(rows) >= 32
    ref = _oracle_rows(net_o, theta, host_noise, idx, scale, obs, rows)
    bound = _row_bound(ref)
    err = np.abs(logits[rows] - ref).max(axis=1)
    assert (err <= bound).all(), (err.max(), bound.min(), np.array(rows)[err > bound])
    srt = np.sort(ref, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > 2 * bound
    assert decided.mean() > 0.5
    np.testing.assert_array_equal(actions[rows][decided], np.argmax(ref, axis=1)[decided])
    # determinism: a second launch on the same inputs is bit-identical (static work lists, no atomics)
    again = sf.forward(d_theta, d_obs, paired=True).cpu().numpy()
    np.testing.assert_array_equal(again, actions)
    np.testing.assert_array_equal(sf.logits.cpu().numpy(), logits)


def test_largemodel_256_slots_tensor_core_vs_simt_all_slots(ctx, host_noise):
    """Every one of the 256 slots: tensor-core convolutions + bulk-copy GEMV (the benchmarked kernels) against the
    plain fp32 SIMT kernels (dne_set_option conv_tc = 0, gemv_bulk = 0) within twice the oracle bound."""
    net = N.make_net("LargeModel")
    P = net.num_params
    rs = np.random.RandomState(4242)
    theta = (rs.randn(P) * 0.05).astype(np.float32)
    idx, scale, active, obs = _setup(rs, 256, 125, P)
    d_theta, d_obs = cuda(theta), cuda(obs)
    L = F.lib()
    out = {}
    try:
        for fast in (1, 0):
            F.check(L.dne_set_option(b"conv_tc", 2 if fast else 0))
            F.check(L.dne_set_option(b"gemv_bulk", fast))
            sf = SlotForward(ctx, net, 256)
            sf.set_slots(idx, scale, active=active)
            a = sf.forward(d_theta, d_obs, paired=True).cpu().numpy()
            out[fast] = (sf.logits.cpu().numpy()[:250], a[:250])
    finally:
        F.check(L.dne_set_option(b"conv_tc", 2))
        F.check(L.dne_set_option(b"gemv_bulk", 1))
    lf, af = out[1]
    ls, as_ = out[0]
    bound = 2 * _row_bound(ls)
    assert (np.abs(lf - ls).max(axis=1) <= bound).all(), np.abs(lf - ls).max()
    srt = np.sort(ls, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > 2 * bound
    np.testing.assert_array_equal(af[decided], as_[decided])


def test_largemodel_four_tables_phase_events_match_one_table(ctx, host_noise):
    """The `--slots 1024`-style schedule: 4 slot tables of 64 on 4 streams with the phase-event hand-off (both modes)
    give bit-identical logits / actions to one 256-slot table."""
    net = N.make_net("LargeModel")
    P = net.num_params
    rs = np.random.RandomState(99)
    theta = cuda((rs.randn(P) * 0.05).astype(np.float32))
    idx, scale, _, obs = _setup(rs, 256, 128, P)
    d_obs = cuda(obs)
    whole = SlotForward(ctx, net, 256)
    whole.set_slots(idx, scale)
    whole.forward(theta, d_obs, paired=True)
    ref_logits, ref_actions = whole.logits.clone(), whole.actions.clone()
    NS, part = 4, 64
    tabs = [SlotForward(ctx, net, part) for _ in range(NS)]
    streams = [torch.cuda.Stream() for _ in range(NS)]
    evs = [torch.cuda.Event() for _ in range(NS)]
    for e in evs:
        e.record()
    for h in range(NS):
        tabs[h].set_slots(idx[h * part:(h + 1) * part], scale[h * part:(h + 1) * part])
    torch.cuda.synchronize()
    for mode in (0, 1):
        for rep in range(3):
            for h in range(NS):
                with torch.cuda.stream(streams[h]):
                    F.check(F.lib().dne_set_phase_events(ctx.handle, C.c_void_p(evs[(h - 1) % NS].cuda_event),
                                                         C.c_void_p(evs[h].cuda_event), mode))
                    tabs[h].forward(theta, d_obs[h * part:(h + 1) * part], paired=True)
        torch.cuda.synchronize()
        got = torch.cat([t.logits for t in tabs])
        # split-K of the shared-theta GEMM depends on the table size (M = 64 vs 256): same math, different partial
        # order, so compare within the forward bound; actions wherever decided
        bound = torch.clamp(ref_logits.abs().max(dim=1).values, min=1.0) * 2e-5
        assert bool(((got - ref_logits).abs().max(dim=1).values <= bound).all()), mode
        srt = ref_logits.sort(dim=1).values
        decided = (srt[:, -1] - srt[:, -2]) > 2 * bound
        assert torch.equal(torch.cat([t.actions for t in tabs])[decided], ref_actions[decided])


@pytest.mark.parametrize("name", ["ESAtariPolicy", "ModelVirtualBN"])
def test_es_atari_policy_vbn_at_config_size(ctx, host_noise, name):
    """frostbite_es.json: ESAtariPolicy with the 128-observation reference batch (es.py:105-113,160-162).  The
    reference pass at n_ref = 128 and the act path on its statistics, 8 slots vs the oracle.  ModelVirtualBN is the
    GPU path's flavour (gpu_implementation/neuroevolution/models/batchnorm.py:50-123): no layer bias, no gamma,
    (x - mean) / sqrt(var + 1e-3) + b."""
    net, net_o = N.make_net(name), O.make_net(name)
    P = net.num_params
    rs = np.random.RandomState(2121)
    theta = (rs.randn(P) * 0.05).astype(np.float32)
    for v in net_o.variables():
        if v.kind == "gamma":
            theta[v.offset:v.offset + v.size] = 1.0 + 0.1 * rs.randn(v.size).astype(np.float32)
    n_slots, n_ref = 8, 128
    pidx = rs.randint(0, NOISE_COUNT - P + 1, size=n_slots // 2).astype(np.int64)
    idx = np.repeat(pidx, 2)
    scale = np.tile([SIGMA, -SIGMA], n_slots // 2).astype(np.float32)
    ref_batch = rs.randint(0, 256, size=(n_ref, 84, 84, 4)).astype(np.uint8)
    obs = rs.randint(0, 256, size=(n_slots, 84, 84, 4)).astype(np.uint8)
    sf = SlotForward(ctx, net, n_slots, n_ref=n_ref)
    sf.set_slots(idx, scale)
    d_theta = cuda(theta)
    sf.vbn_reference_pass(d_theta, cuda(ref_batch))
    actions = sf.forward(d_theta, cuda(obs), paired=True).cpu().numpy()
    logits, vbn = sf.logits.cpu().numpy(), sf.vbn.cpu().numpy()
    ref_logits = []
    for s in range(n_slots):
        th = (theta + np.float32(scale[s]) * host_noise[idx[s]:idx[s] + P]).astype(np.float32)
        _, stats = O.forward(net_o, th, ref_batch, is_ref=True)
        off = 0
        for mean, var in stats:
            c = mean.size
            np.testing.assert_allclose(vbn[s, off:off + c], mean, rtol=2e-4, atol=2e-5)
            np.testing.assert_allclose(vbn[s, off + c:off + 2 * c], var, rtol=5e-4, atol=1e-6)
            off += 2 * c
        ref_logits.append(O.forward(net_o, th, obs[s:s + 1], vbn_stats=stats)[0][0])
    ref = np.stack(ref_logits)
    bound = _row_bound(ref, 5e-4)
    assert (np.abs(logits - ref).max(axis=1) <= bound).all(), np.abs(logits - ref).max()
    srt = np.sort(ref, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > 2 * bound
    np.testing.assert_array_equal(actions[decided], np.argmax(ref, axis=1)[decided])
