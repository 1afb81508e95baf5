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


# ---- size-independent properties (hypothesis): the same properties the GPU suite checks at BASELINE sizes ----------
from hypothesis import given, settings, strategies as st   # noqa: E402
from hypothesis.extra import numpy as hnp                  # noqa: E402

_f32s = st.floats(min_value=-1e4, max_value=1e4, allow_nan=False, width=32)


@settings(max_examples=60, deadline=None)
@given(hnp.arrays(np.float32, st.integers(2, 200), elements=_f32s))
def test_ranks_are_the_stable_sort_permutation(x):
    r = O.compute_ranks(x)
    assert sorted(r.tolist()) == list(range(len(x)))                      # a permutation of 0..n-1
    order = np.argsort(x, kind="stable")
    np.testing.assert_array_equal(r[order], np.arange(len(x)))           # rank of the i-th smallest is i; ties by index
    c = O.compute_centered_ranks(x)
    assert c.dtype == np.float32 and c.min() == np.float32(-0.5) and c.max() == np.float32(0.5)
    y = x * np.float32(3.0) + np.float32(1.0)                            # es.py:77 is invariant under order-preserving maps
    if np.unique(y).size == np.unique(x).size:                           # (unless float32 rounding merged two values)
        np.testing.assert_array_equal(O.compute_ranks(y), r)


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 12), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
def test_es_gradient_linearity_and_antithetic_cancellation(n, dim, seed):
    rs = np.random.RandomState(seed)
    noise = rs.randn(4096).astype(np.float32)
    idx = rs.randint(0, len(noise) - dim + 1, size=n)
    a = rs.randn(n, 2).astype(np.float32)
    b = rs.randn(n, 2).astype(np.float32)
    ga, gb = O.es_gradient(a, noise, idx, dim), O.es_gradient(b, noise, idx, dim)
    gab = O.es_gradient((a.astype(np.float64) + 2.0 * b).astype(np.float64), noise, idx, dim)
    np.testing.assert_allclose(gab, ga + 2.0 * gb, rtol=1e-5, atol=1e-6)
    same = np.repeat(a[:, :1], 2, axis=1)                                 # w+ == w-  ->  zero gradient (es.py:292)
    assert np.all(O.es_gradient(same, noise, idx, dim) == 0)
    # batched_weighted_sum in slabs == one slab (es.py:115-122) up to float32 re-association
    w = (a[:, 0] - a[:, 1]).astype(np.float32)
    s1 = O.batched_weighted_sum(w, noise, idx, dim, batch_size=3)
    s2 = O.batched_weighted_sum(w, noise, idx, dim, batch_size=500)
    np.testing.assert_allclose(s1, s2, rtol=1e-4, atol=1e-4)


@settings(max_examples=60, deadline=None)
@given(hnp.arrays(np.float32, st.integers(1, 120), elements=st.sampled_from([0.0, 10.0, 20.0, 30.0, -10.0, 7.5])), st.data())
def test_ga_truncate_is_pythons_stable_descending_sort(fit, data):
    T = data.draw(st.integers(1, len(fit)))
    want = sorted(range(len(fit)), key=lambda i: fit[i], reverse=True)[:T]     # gpu_implementation/ga.py:180
    np.testing.assert_array_equal(O.ga_truncate(fit, T), want)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 9), st.integers(1, 9), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_bc_distance_padding_and_symmetry(n, m, d, seed):
    rs = np.random.RandomState(seed)
    x = rs.randint(0, 256, size=(n, d)).astype(np.float64)
    y = rs.randint(0, 256, size=(m, d)).astype(np.float64)
    dist = O.euclidean_distance(x, y)
    assert dist == O.euclidean_distance(y, x) and O.euclidean_distance(x, x) == 0.0
    L = max(n, m)                                                        # nses.py:12-20: pad the shorter with its last row
    xp = np.concatenate([x, np.repeat(x[-1:], L - n, axis=0)])
    yp = np.concatenate([y, np.repeat(y[-1:], L - m, axis=0)])
    np.testing.assert_allclose(dist, np.sqrt(np.square(xp - yp).sum()), rtol=1e-12)
    arch = [rs.randint(0, 256, size=(rs.randint(1, 9), d)).astype(np.uint8) for _ in range(5)]
    k = 3
    nov = O.compute_novelty_vs_archive(arch, x.astype(np.uint8), k)
    ds = sorted(O.euclidean_distance(a.astype(np.float64), x) for a in arch)
    np.testing.assert_allclose(nov, np.mean(ds[:k]), rtol=1e-12)


# ---- pins against the reference's own novelty / selection / warp expressions (tests/golden/make_golden_nses_ga.py) ------
@pytest.fixture(scope="module")
def golden2():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_nses_ga.npz"))


def _unpad(pad, lens):
    return [pad[i, :lens[i]] for i in range(len(lens))]


def test_novelty_matches_reference_nses(golden2):
    """nses.py:12-32 imported from the reference (stub tensorflow): ragged lengths, k below / at / above the archive size."""
    qs, as_ = _unpad(golden2["nov_q_pad"], golden2["nov_q_len"]), _unpad(golden2["nov_a_pad"], golden2["nov_a_len"])
    d = np.array([[O.euclidean_distance(a.astype(np.float64), q.astype(np.float64)) for a in as_] for q in qs])
    np.testing.assert_array_equal(d, golden2["nov_dist"])
    for k in (1, 10, 19, 40):
        got = np.array([O.compute_novelty_vs_archive(as_, q, k) for q in qs])
        np.testing.assert_allclose(got, golden2[f"nov_k{k}"], rtol=1e-15)
    got = np.array([O.compute_novelty_vs_archive(as_[:3], q, 10) for q in qs])
    np.testing.assert_allclose(got, golden2["nov_k10_arch3"], rtol=1e-15)


@pytest.mark.parametrize("pop,T", [(1000, 20), (64, 64), (7, 3)])
def test_ga_truncation_matches_reference_argpartition(golden2, pop, T):
    """ga.py:145-149: np.argpartition(returns, (-T, -1))[-1:-T-1:-1] guarantees the best individual first and the top-T
    SET (the order in between is unspecified); the canonical stable-descending rule must agree on both."""
    fit, ref = golden2[f"ga_fit_{pop}_{T}"], golden2[f"ga_sel_{pop}_{T}"]
    got = O.ga_truncate(fit, T)
    assert got[0] == ref[0] and fit[got[0]] == fit.max()
    assert set(got.tolist()) == set(ref.tolist())


def test_warp_frame_cpu_matches_pillow(golden2):
    """atari_wrappers.py:138-142 evaluated with numpy + Pillow (recorded).  The Pillow resize (area-scaled triangle filter,
    double accumulation, float32 intermediate, truncating uint8 cast) is pinned BIT-EXACTLY on the recorded gray frames;
    the gray itself is a BLAS sgemv in the reference and only defined to 1 ulp, which moves at most a handful of output
    pixels by one level."""
    for i, rgb in enumerate(golden2["warp_rgb"]):
        gray_ref = golden2["warp_gray_f32"][i]
        np.testing.assert_array_equal(O.resize_pillow_bilinear(gray_ref), golden2["warp_out"][i])
        gray = O.gray_rgb(rgb)
        assert np.abs(gray - gray_ref).max() <= np.spacing(np.float32(255.0))            # 1 ulp at the top of the range
        full = O.warp_frame_cpu(rgb).astype(np.int32)
        diff = np.abs(full - golden2["warp_out"][i].astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() < 0.01


def test_warp_frame_cpu_live_pillow():
    """Same pin against the Pillow installed next to the test (skipped where Pillow is absent)."""
    Image = pytest.importorskip("PIL.Image")
    rs = np.random.RandomState(5)
    for frame in (rs.rand(210, 160).astype(np.float32) * 255, rs.randint(0, 256, size=(210, 160)).astype(np.float32)):
        ref = np.array(Image.fromarray(frame).resize((84, 84), resample=Image.BILINEAR), dtype=np.uint8)
        np.testing.assert_array_equal(O.resize_pillow_bilinear(frame), ref)


def test_warp_frame_gpu_formula_properties():
    """tf_atari.py:90-92 by formula: align_corners=True maps the four corners onto themselves, a constant image stays
    constant, and the result lies between the min and max of the 2-frame maximum."""
    pal = O.ntsc_gray_palette()
    rs = np.random.RandomState(2)
    idx = (rs.randint(0, 128, size=(2, 210, 160)) * 2).astype(np.uint8)
    out = O.warp_frame_gpu(idx, pal)
    g = np.maximum(pal[idx[0]], pal[idx[1]])
    assert out.shape == (84, 84) and out.dtype == np.float32
    assert out[0, 0] == g[0, 0] and out[0, -1] == g[0, -1] and out[-1, 0] == g[-1, 0] and out[-1, -1] == g[-1, -1]
    assert out.min() >= g.min() - 1e-6 and out.max() <= g.max() + 1e-6
    const = np.full((2, 210, 160), 6, dtype=np.uint8)
    np.testing.assert_array_equal(O.warp_frame_gpu(const, pal), np.full((84, 84), pal[6], dtype=np.float32))


# --------------------------------------------------------------------------------------------------
# Pinned to the reference's OWN model classes, executed under a shape-only TensorFlow stand-in
# (tests/golden/make_golden_models.py -> ref_models.npz): flat layout (8a-3), initialisation scale and seed-chain
# genome materialisation (8a-11) of gpu_implementation/neuroevolution/models/{base,dqn,batchnorm}.py.
# --------------------------------------------------------------------------------------------------
import os  # noqa: E402

_REF_MODELS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_models.npz")


@pytest.mark.parametrize("name", ["Model", "LargeModel", "ModelVirtualBN"])
def test_layout_and_init_scale_match_reference_models(name):
    g = np.load(_REF_MODELS)
    net = O.make_net(name)
    ref_names = [str(s) for s in g[f"{name}.names"]]
    ref_shapes = [tuple(int(d) for d in str(s).split(",")) for s in g[f"{name}.shapes"]]
    assert net.num_params == int(g[f"{name}.num_params"])
    got = net.variables()
    assert len(got) == len(ref_names)
    off = 0
    for v, rn, rsh in zip(got, ref_names, ref_shapes):
        assert rn.split("/")[-1] == ("w" if v.kind == "w" else "b"), (rn, v)            # creation order: w then b per layer
        assert int(np.prod(rsh)) == v.size and tuple(d for d in rsh if d != 1) == tuple(d for d in v.shape if d != 1), (rn, rsh, v.shape)
        assert v.offset == off
        off += v.size
    # per-variable initialisation scale, rounded to float32 as numpy 1.x does when it multiplies the float32 ones vector
    scale = O.ga_scale_by(net)
    for v, s in zip(got, g[f"{name}.var_scale_by"]):
        assert np.all(scale[v.offset:v.offset + v.size] == np.float32(s)), (v.name, float(s))
    assert abs(float(scale.astype(np.float64).sum()) - float(g[f"{name}.scale_by_sum"])) <= 1e-6 * float(g[f"{name}.scale_by_sum"])


@pytest.mark.parametrize("name", ["Model", "LargeModel", "ModelVirtualBN"])
def test_genome_materialisation_matches_reference_models(name):
    """base.py:127-156 executed by the reference itself (float64 under numpy >= 2, see the generator's note) against the
    float32 oracle: every sampled coordinate within float32 rounding of the 4-term chain."""
    g = np.load(_REF_MODELS)
    net = O.make_net(name)
    noise = O.noise_table(24_000_000)
    idx, power = g[f"{name}.seed_idx"], g[f"{name}.seed_power"]
    seeds = (int(idx[0]),) + tuple((int(i), float(p)) for i, p in zip(idx[1:], power[1:]))
    theta = O.ga_materialize_gpu(net, noise, seeds)
    assert theta.dtype == np.float32
    ref = g[f"{name}.theta_samples"]
    np.testing.assert_allclose(theta[::997].astype(np.float64), ref, rtol=3e-7, atol=1e-9)
    t64 = theta.astype(np.float64)
    assert abs(t64.sum() - float(g[f"{name}.theta_sum"])) <= 1e-4
    assert abs(np.square(t64).sum() - float(g[f"{name}.theta_sumsq"])) <= 1e-6 * float(g[f"{name}.theta_sumsq"])


# --------------------------------------------------------------------------------------------------
# Pinned to the reference's CPU-path builders (es_distributed/policies.py `_make_net` + tf_util layer functions) executed
# under the shape-only TensorFlow stand-in of tests/golden/make_golden_policies.py -> ref_policies.npz
# --------------------------------------------------------------------------------------------------
_REF_POLICIES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_policies.npz")


@pytest.mark.parametrize("key,name,kw", [
    ("GAAtariPolicy", "GAAtariPolicy", {}),
    ("MujocoPolicy.continuous", "MujocoPolicy", dict(ob_dim=376, hidden=(256, 256), ac_dim=17)),
    ("MujocoPolicy.uniform10", "MujocoPolicy", dict(ob_dim=376, hidden=(256, 256), ac_dim=170)),
])
def test_flat_layout_matches_reference_policy_builders(key, name, kw):
    """tf_util.GetFlat / SetFromFlat (tf_util.py:224-246) concatenate the trainable variables in creation order."""
    g = np.load(_REF_POLICIES)
    net = O.make_net(name, **kw)
    names = [str(s) for s in g[key + ".names"]]
    shapes = [tuple(int(d) for d in str(s).split(",")) for s in g[key + ".shapes"]]
    assert net.num_params == int(g[key + ".num_params"])
    got = net.variables()
    assert len(got) == len(names)
    off = 0
    for v, rn, rsh in zip(got, names, shapes):
        assert rn.split("/")[-1] == ("w" if v.kind == "w" else "b"), (rn, v)
        assert int(np.prod(rsh)) == v.size and tuple(d for d in rsh if d != 1) == tuple(d for d in v.shape if d != 1), (rn, rsh, v.shape)
        assert v.offset == off
        off += v.size


def test_reinitialize_matches_reference_closures_bit_exactly():
    """Policy.reinitialize (policies.py:42-44): the numpy closure of tf_util._normalize run by the generator on a seeded flat
    vector, variable by variable, against oracle.ga_reinitialize -- every bit."""
    import hashlib
    g = np.load(_REF_POLICIES)
    net = O.make_net("GAAtariPolicy")
    flat = np.random.RandomState(int(g["GAAtariPolicy.reinit_in_seed"])).randn(net.num_params).astype(np.float32)
    out = O.ga_reinitialize(net, flat)
    assert out.dtype == np.float32
    np.testing.assert_array_equal(out[::499], g["GAAtariPolicy.reinit_samples"])
    assert hashlib.sha1(out.tobytes()).hexdigest() == str(g["GAAtariPolicy.reinit_sha1"])
