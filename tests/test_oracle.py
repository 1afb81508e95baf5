"""Pins oracle/oracle.py against outputs of the reference's own numpy code (tests/golden/ref_numpy.npz,
made by tests/golden/make_golden.py) and cross-checks the TF-restated pieces against an independent
naive implementation.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O


def test_noise_prefix_matches_reference(golden, small_noise):
    assert small_noise.dtype == np.float32 and small_noise.size == int(golden["noise_count"])
    np.testing.assert_array_equal(small_noise[:64], golden["noise_head"])
    np.testing.assert_array_equal(small_noise[-64:], golden["noise_tail"])
    assert small_noise.astype(np.float64).sum() == float(golden["noise_sum64"])
    # es.py:60 first values, quoted in SURVEY.md 8c
    np.testing.assert_allclose(small_noise[:3], [-1.0856307, 0.99734545, 0.2829785], rtol=0, atol=1e-7)


def test_sample_index_stream(golden):
    stream = np.random.RandomState(7)
    got = [O.sample_index(stream, 250_000_000, 4052658) for _ in range(16)]
    np.testing.assert_array_equal(got, golden["sample_index_P4052658"])


@pytest.mark.parametrize("n", [1, 8, 500, 5000])
def test_ranks_bit_exact_tie_free(golden, n):
    x = golden[f"rank_in_{n}"]
    np.testing.assert_array_equal(O.compute_ranks(x.ravel()), golden[f"rank_ranks_{n}"])
    got = O.compute_centered_ranks(x)
    assert got.dtype == np.float32
    if n > 1:   # n == 1 -> size-1 == 1, fine; n==... all finite
        np.testing.assert_array_equal(got, golden[f"rank_centered_{n}"])


def test_ranks_stable_tie_rule():
    x = np.array([10, 0, 10, 0, 20, 10], dtype=np.float32)
    np.testing.assert_array_equal(O.compute_ranks(x), [2, 0, 3, 1, 5, 4])


def test_es_gradient_matches_reference(golden, small_noise):
    P, idx, returns = int(golden["grad_P"]), golden["grad_idx"], golden["grad_returns"]
    proc = O.compute_centered_ranks(returns)
    g32 = O.es_gradient(proc, small_noise, idx, P, dtype=np.float32)
    g64 = O.es_gradient(proc, small_noise, idx, P, dtype=np.float64)
    ref = golden["grad_g"]
    scale = np.abs(ref).max()
    assert np.abs(g32 - ref).max() <= 1e-6 * scale          # same algorithm, BLAS order may differ
    assert np.abs(g64 - ref).max() <= 1e-5 * scale          # float64 referee vs reference float32


def test_optimizers_match_reference(golden):
    theta0, grads = golden["opt_theta0"], golden["opt_grads"]
    adam, sgd = O.Adam(theta0, 0.01), O.SGD(theta0, 0.01, 0.9)
    for k, gk in enumerate(grads):
        r, t = adam.update(O.es_update_direction(gk, adam.theta, 0.005))
        assert t.dtype == np.float32
        np.testing.assert_allclose(t, golden["adam_theta"][k], rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(r, golden["adam_ratio"][k], rtol=1e-5)
        r, t = sgd.update(O.es_update_direction(gk, sgd.theta, 0.005))
        np.testing.assert_allclose(t, golden["sgd_theta"][k], rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(r, golden["sgd_ratio"][k], rtol=1e-5)


def test_running_stat(golden):
    st = O.RunningStat((5,), eps=1e-2)
    obs = golden["rstat_obs"]
    st.increment(obs.sum(axis=0), np.square(obs).sum(axis=0), len(obs))
    np.testing.assert_array_equal(st.mean, golden["rstat_mean"])
    np.testing.assert_array_equal(st.std, golden["rstat_std"])


@pytest.mark.parametrize("name,P", [("LargeModel", 4052658), ("ESAtariPolicy", 1009058),
                                    ("GAAtariPolicy", 1008450), ("Model", 1008450), ("MujocoPolicy", 166673)])
def test_param_counts(name, P):
    assert O.make_net(name).num_params == P


@pytest.mark.parametrize("k,s,h,cin,cout", [(8, 4, 84, 4, 8), (4, 2, 21, 8, 8), (3, 1, 11, 8, 4)])
def test_conv_same_vs_naive(k, s, h, cin, cout):
    rs = np.random.RandomState(0)
    x = rs.rand(2, h, h, cin).astype(np.float32)
    w = rs.randn(k, k, cin, cout).astype(np.float32)
    got = O._conv_same(x, w, s)
    ref = O.forward_naive_conv(x, w, s)
    assert got.shape == ref.shape == (2, -(-h // s), -(-h // s), cout)
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-4)


def test_antithetic_symmetry(small_noise):
    # gpu_implementation/es.py:182-183
    net = O.make_net("MujocoPolicy")
    theta = small_noise[1000:1000 + net.num_params].copy()
    pos, neg = O.perturb(theta, small_noise, 777, 0.02, +1), O.perturb(theta, small_noise, 777, 0.02, -1)
    assert np.max(np.abs((pos + neg) / 2 - theta)) < 1e-5


def test_vbn_reference_pass_normalises():
    net = O.make_net("ESAtariPolicy")
    rs = np.random.RandomState(1)
    theta = (rs.randn(net.num_params) * 0.05).astype(np.float32)
    for v in net.variables():          # gamma = 1 as TF initialises it
        if v.kind == "gamma":
            theta[v.offset:v.offset + v.size] = 1.0
    ref = rs.randint(0, 256, size=(16, 84, 84, 4)).astype(np.uint8)
    logits, stats = O.forward(net, theta, ref, is_ref=True)
    assert len(stats) == 3 and stats[0][0].shape == (16,) and stats[2][0].shape == (256,)
    logits2, _ = O.forward(net, theta, ref, vbn_stats=stats)
    np.testing.assert_allclose(logits, logits2, rtol=1e-4, atol=1e-4)   # decay=0: moving stats == batch stats


def test_ga_paths(small_noise):
    net = O.make_net("GAAtariPolicy", num_actions=6)
    P = net.num_params
    noise = O.noise_table(2 * P + 10)
    th = O.ga_materialize_cpu(net, noise, [3, P, 7], 0.005)
    w = O.unflatten(net, O.ga_reinitialize(net, noise[3:3 + P]))
    np.testing.assert_allclose(np.sqrt(np.square(w[0]["w"].reshape(-1, 16)).sum(0)), 1.0, rtol=1e-5)
    np.testing.assert_allclose(np.sqrt(np.square(w[3]["w"]).sum(0)), 0.1, rtol=1e-5)
    assert np.all(w[0]["b"] == 0)
    assert th.dtype == np.float32
    th2 = O.ga_materialize_gpu(net, noise, (5, (9, 0.002)))
    sb = O.ga_scale_by(net)
    np.testing.assert_allclose(th2, noise[5:5 + P] * sb + np.float32(0.002) * noise[9:9 + P], rtol=1e-6, atol=1e-8)
    fit = np.array([10, 50, 50, 0, 70, 50], dtype=np.float32)
    np.testing.assert_array_equal(O.ga_truncate(fit, 4), [4, 1, 2, 5])


def test_novelty():
    rs = np.random.RandomState(3)
    arch = [rs.randint(0, 256, size=(t, 128)).astype(np.uint8) for t in (5, 9, 3, 7)]
    q = rs.randint(0, 256, size=(6, 128)).astype(np.uint8)
    # padded formulation used on the device: pad both with their last row to max(len) and take L2
    for a in arch:
        T = max(len(a), len(q))
        ap = np.concatenate([a, np.repeat(a[-1:], T - len(a), 0)]).astype(np.float64)
        qp = np.concatenate([q, np.repeat(q[-1:], T - len(q), 0)]).astype(np.float64)
        np.testing.assert_allclose(O.euclidean_distance(a.astype(np.float64), q.astype(np.float64)),
                                   np.linalg.norm(ap - qp), rtol=1e-12)
    nov = O.compute_novelty_vs_archive(arch, q, k=2)
    d = sorted(O.euclidean_distance(a.astype(np.float64), q.astype(np.float64)) for a in arch)
    assert nov == pytest.approx((d[0] + d[1]) / 2)
