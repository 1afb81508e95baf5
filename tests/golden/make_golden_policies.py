"""Generate tests/golden/ref_policies.npz by EXECUTING the reference's CPU-path network builders and initialisers.

Run in the build container only (imports /root/reference; nothing under tests/ reads that path at test time):
    python tests/golden/make_golden_policies.py

es_distributed/policies.py + tf_util.py build their graphs with TensorFlow 1.x calls (absent here).  Same device as
make_golden_models.py: a shape-only stand-in for `tensorflow` lets the reference's own `_make_net` methods and `tf_util`
layer functions run unmodified, which pins

  * 8a-3: the creation order, names and shapes of the trainable variables of GAAtariPolicy (policies.py:449-459) and
    MujocoPolicy (policies.py:155-196; 'continuous:' and 'uniform:10' heads) -- tf_util.GetFlat / SetFromFlat concatenate
    them in exactly that order (tf_util.py:224-246);
  * 8a-11: `Policy.reinitialize` (policies.py:42-44): the numpy closure inside tf_util._normalize (tf_util.py:122-130) and the
    bias reset are pulled out of the stand-in's `tf.py_func` / `assign` records and run on a random flat vector -- the
    reference's own arithmetic for `v = reinitialize(noise[seed0])` of ga.py:256-260;
  * tf_util.normc_initializer (tf_util.py:108-119) on the global numpy stream after np.random.seed.

Not covered: ESAtariPolicy (its `layers.batch_norm` creates beta / gamma / moving statistics inside TensorFlow).
"""
import hashlib
import importlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


class Dim(int):
    @property
    def value(self):
        return int(self)


class ShapeList(list):
    def as_list(self):
        return [None if d is None else int(d) for d in self]


class PyFunc:
    """Record of tf.py_func(fn, inputs, dtype): keeps the numpy closure so that the fixture can run it."""
    def __init__(self, fn, inputs):
        self.fn, self.inputs = fn, inputs

    def set_shape(self, shape):
        pass


def _bshape(sa, sb):
    n = max(len(sa), len(sb))
    sa, sb = (1,) * (n - len(sa)) + tuple(sa), (1,) * (n - len(sb)) + tuple(sb)
    out = []
    for x, y in zip(sa, sb):
        if x is None or y is None:
            out.append(None)
        else:
            assert x == y or x == 1 or y == 1, (sa, sb)
            out.append(max(x, y))
    return tuple(out)


class T:
    """Fake tensor / variable: a static shape (None = unknown batch) and whatever attributes the reference hangs on it."""
    def __init__(self, shape, name=None):
        self.shape_ = tuple(None if d is None else int(d) for d in shape)
        self.name = name

    def get_shape(self):
        return ShapeList(None if d is None else Dim(d) for d in self.shape_)

    def _bin(self, other):
        so = other.shape_ if isinstance(other, T) else tuple(np.shape(other))
        return T(_bshape(self.shape_, so))
    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = _bin

    def assign(self, value):
        return ("assign", self, value)


def make_tf():
    tf = types.ModuleType("tensorflow")
    tf._scopes, tf.created = [], []
    tf.float32, tf.int32 = "float32", "int32"

    class Scope:
        def __init__(self, name):
            self.name = name

        def __enter__(self):
            tf._scopes.append(self.name)
            return self

        def __exit__(self, *a):
            tf._scopes.pop()

    def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
        full = "/".join(tf._scopes + [name])
        v = T([int(d) for d in shape], full)
        v.trainable, v.initializer_ = trainable, initializer
        tf.created.append(v)
        return v

    def reshape(x, shape):
        shape = [d if d is None else int(d) for d in shape]
        known = [d for d in x.shape_ if d is not None]
        if None in x.shape_:                             # unknown batch: -1 keeps it unknown
            return T([None if d == -1 else d for d in shape])
        total = int(np.prod(known))
        if -1 in shape:
            shape[shape.index(-1)] = total // int(np.prod([d for d in shape if d != -1]))
        return T(shape)

    def conv2d(x, w, strides, padding):
        n, h, wd, c = x.shape_
        k, s = w.shape_[0], strides[1]
        assert c == w.shape_[2] and padding == "SAME"
        return T((n, -(-h // s), -(-wd // s), w.shape_[3]))

    def matmul(a, b):
        assert a.shape_[-1] == b.shape_[0], (a.shape_, b.shape_)
        return T((a.shape_[0], b.shape_[1]))

    ident = lambda x, *a, **kw: x                        # noqa: E731
    tf.variable_scope = lambda name, *a, **kw: Scope(name)
    tf.get_variable, tf.reshape, tf.matmul = get_variable, reshape, matmul
    tf.placeholder = lambda dtype, shape=None: T(shape)
    tf.py_func = lambda fn, inputs, dtype: PyFunc(fn, inputs)
    tf.zeros_initializer = "zeros_initializer"
    tf.constant_initializer = lambda v: ("constant", v)
    tf.zeros_like = lambda x: ("zeros_like", x)
    tf.assign = lambda v, x: ("assign", v, x)
    tf.tanh = tf.to_float = ident
    tf.clip_by_value = ident
    tf.argmax = lambda x, axis: T(tuple(d for i, d in enumerate(x.shape_) if i != axis))
    tf.shape = lambda x: list(x.shape_)
    tf.nn = types.SimpleNamespace(relu=ident, elu=ident, conv2d=conv2d)
    contrib = types.ModuleType("tensorflow.contrib")
    layers = types.ModuleType("tensorflow.contrib.layers")
    contrib.layers = layers
    tf.contrib = contrib
    return tf, contrib, layers


def load_reference(tf, contrib, layers):
    for k in [k for k in sys.modules if k == "refes" or k.startswith("refes.")]:
        del sys.modules[k]
    sys.modules.update({"tensorflow": tf, "tensorflow.contrib": contrib, "tensorflow.contrib.layers": layers,
                        "h5py": types.ModuleType("h5py")})
    pkg = types.ModuleType("refes")
    pkg.__path__ = [os.path.join(REF, "es_distributed")]
    sys.modules["refes"] = pkg
    return importlib.import_module("refes.tf_util"), importlib.import_module("refes.policies")


def record(out, key, tf):
    tv = [v for v in tf.created if v.trainable]
    out[key + ".names"] = np.array([v.name for v in tv])
    out[key + ".shapes"] = np.array([",".join(str(d) for d in v.shape_) for v in tv])
    out[key + ".num_params"] = np.int64(sum(int(np.prod(v.shape_)) for v in tv))
    out[key + ".non_trainable"] = np.array([v.name for v in tf.created if not v.trainable] or [""])
    return tv


def main():
    out = {}
    Box = types.SimpleNamespace

    # ---- GAAtariPolicy: layout + reinitialize ------------------------------------------------------------------
    tf, contrib, layers = make_tf()
    U, P = load_reference(tf, contrib, layers)
    pol = object.__new__(P.GAAtariPolicy)                # _initialize's graph part, without the session-bound helpers
    pol.nonlin, pol.num_actions, pol.ac_init_std = tf.nn.relu, 18, 0.1
    with tf.variable_scope("GAAtariPolicy"):
        pol._make_net(tf.placeholder(tf.float32, [None, 84, 84, 4]))
    tv = record(out, "GAAtariPolicy", tf)
    rs = np.random.RandomState(99)
    n = int(out["GAAtariPolicy.num_params"])
    flat = rs.randn(n).astype(np.float32)                # stands for noise[seed0 : seed0 + num_params] (ga.py:256)
    pieces, off = [], 0
    for v in tv:                                         # Policy.reinitialize: v.reinitialize.eval() for every trainable variable
        size = int(np.prod(v.shape_))
        cur = flat[off:off + size].reshape(v.shape_).copy()
        op = v.reinitialize
        assert op[0] == "assign" and op[1] is v
        if isinstance(op[2], PyFunc):                    # tf_util._normalize: the reference's numpy closure
            cur = np.asarray(op[2].fn(cur), dtype=np.float32)
        else:                                            # biases: assign(zeros_like(b))
            assert op[2][0] == "zeros_like"
            cur = np.zeros_like(cur)
        pieces.append(cur.reshape(-1))
        off += size
    re = np.concatenate(pieces)
    out["GAAtariPolicy.reinit_in_seed"] = np.int64(99)
    out["GAAtariPolicy.reinit_samples"] = re[::499].copy()
    out["GAAtariPolicy.reinit_sha1"] = np.array(hashlib.sha1(re.tobytes()).hexdigest())

    # ---- MujocoPolicy: layouts of the continuous and the binned head -----------------------------------------------
    for key, bins_ in (("MujocoPolicy.continuous", "continuous:"), ("MujocoPolicy.uniform10", "uniform:10")):
        tf, contrib, layers = make_tf()
        U, P = load_reference(tf, contrib, layers)
        pol = object.__new__(P.MujocoPolicy)
        pol.nonlin, pol.hidden_dims, pol.connection_type, pol.ac_bins = tf.tanh, [256, 256], "ff", bins_
        pol.ac_space = Box(shape=(17,), high=np.ones(17, np.float32), low=-np.ones(17, np.float32))
        with tf.variable_scope("MujocoPolicy"):
            pol._make_net(tf.placeholder(tf.float32, [None, 376]))
        record(out, key, tf)

    # ---- normc_initializer on the global numpy stream -----------------------------------------------------------
    tf, contrib, layers = make_tf()
    U, P = load_reference(tf, contrib, layers)
    for i, (shape, std) in enumerate((((8, 8, 4, 16), 1.0), ((256, 18), 0.1), ((376, 256), 1.0), ((256, 17), 0.01))):
        np.random.seed(1000 + i)
        arr = U.normc_initializer(std)(list(shape)).fn()
        out[f"normc.{i}.shape"] = np.array(shape, dtype=np.int64)
        out[f"normc.{i}.std"] = np.float64(std)
        out[f"normc.{i}.sha1"] = np.array(hashlib.sha1(np.ascontiguousarray(arr, dtype=np.float32).tobytes()).hexdigest())
        out[f"normc.{i}.head"] = np.asarray(arr, dtype=np.float32).reshape(-1)[:16].copy()
        assert arr.dtype == np.float32
    for k in sorted(out):
        if k.endswith(".names"):
            print(k, list(zip(out[k].tolist(), out[k.replace('.names', '.shapes')].tolist())), int(out[k.replace('.names', '.num_params')]))
    np.savez_compressed(os.path.join(HERE, "ref_policies.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_policies.npz"))


if __name__ == "__main__":
    main()
