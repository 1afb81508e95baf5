"""Generate tests/golden/ref_models.npz by EXECUTING the reference's own model classes.

Run in the build container only (imports /root/reference; nothing under tests/ reads that path at test time):
    python tests/golden/make_golden_models.py

gpu_implementation/neuroevolution/models/{base,dqn,batchnorm}.py build their networks with TensorFlow 1.x graph calls, and
TensorFlow is absent from this image.  What the flat parameter layout, the per-parameter initialisation scale
and the seed-chain genome materialisation (8a-11: base.py `make_weights`, `compute_weights_from_seeds`, `compute_mutation`)
depend on is only the ORDER, NAMES and SHAPES of the `tf.get_variable` calls plus plain numpy -- so a shape-only stand-in for
the `tensorflow` module (fake tensors that carry a static shape; variable scopes that build names; no arithmetic) is enough
to run `Model`, `LargeModel` and `ModelVirtualBN` `.make_net()` + `BaseModel.make_weights()` unmodified and record

  * the variables in creation order: scoped name, per-member shape, `scale_by`;
  * `num_params` and the concatenated `scale_by` vector (as the reference builds it);
  * theta = `compute_weights_from_seeds(noise, seeds)` for a 4-entry genome on the reference noise table
    (np.random.RandomState(123).randn, float32 -- tests/golden/make_golden.py pins that table to es.py:51-58),
    stored as strided samples + float64 sums (the vectors themselves are 4 MB .. 16 MB).

numpy >= 2 note (same hazard as the optimizers): `scale_by` is an np.float64 scalar times a float32 ones vector,
which numpy 1.x (the reference's era) keeps float32 and numpy >= 2 promotes to float64, so the recorded theta is float64 here;
the oracle (float32 throughout, numpy-1 semantics) is compared to it at float32 rounding accuracy and to `scale_by` exactly
after rounding to float32.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/gpu_implementation"
HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------------------------
# shape-only stand-in for tensorflow
# ---------------------------------------------------------------------------------------------------------------
class Dim(int):
    """A static dimension: an int with the `.value` attribute of tf.Dimension."""
    @property
    def value(self):
        return int(self)


def _bshape(a, b):
    sa = tuple(a.shape_) if isinstance(a, T) else ()
    sb = tuple(b.shape_) if isinstance(b, T) else ()
    n = max(len(sa), len(sb))
    sa, sb = (1,) * (n - len(sa)) + sa, (1,) * (n - len(sb)) + sb
    out = []
    for x, y in zip(sa, sb):
        assert x == y or x == 1 or y == 1, (sa, sb)
        out.append(max(x, y))
    return tuple(out)


class T:
    """Fake tensor: a static shape and nothing else."""
    def __init__(self, shape, name=None):
        self.shape_ = tuple(int(d) for d in shape)
        self.name = name

    def get_shape(self):
        return [Dim(d) for d in self.shape_]

    def _bin(self, other):
        return T(_bshape(self, other))
    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = _bin

    def __getitem__(self, sl):
        assert isinstance(sl, slice) and len(self.shape_) == 1
        return T((len(range(*sl.indices(self.shape_[0]))),))


class Scope:
    def __init__(self, tfm, name, reuse=False):
        self.tfm, self.name, self.reuse = tfm, name, reuse

    def __enter__(self):
        self.tfm._scopes.append(self.name)
        return self

    def __exit__(self, *a):
        self.tfm._scopes.pop()


def make_tf():
    tf = types.ModuleType("tensorflow")
    tf._scopes, tf._vars, tf.created = [], {}, []
    tf.float32, tf.int32 = "float32", "int32"

    def variable_scope(name, *a, **kw):
        assert isinstance(name, str), "only named scopes are exercised"
        return Scope(tf, name)

    def get_variable(name, shape=None, trainable=True, **kw):
        full = "/".join(tf._scopes + [name])
        if full not in tf._vars:
            tf._vars[full] = T(shape, full)
            tf.created.append(full)
        return tf._vars[full]

    def reshape(x, shape):
        shape = [int(d) for d in shape]
        total = int(np.prod(x.shape_))
        if -1 in shape:
            known = int(np.prod([d for d in shape if d != -1]))
            shape[shape.index(-1)] = total // known
        assert int(np.prod(shape)) == total, (x.shape_, shape)
        return T(shape)

    def extract_image_patches(x, ksizes, strides, rates, padding):
        n, h, w, c = x.shape_
        k, s = ksizes[1], strides[1]
        if padding == "SAME":
            oh, ow = -(-h // s), -(-w // s)
        else:
            oh, ow = (h - k) // s + 1, (w - k) // s + 1
        return T((n, oh, ow, k * k * c))

    def matmul(a, b):
        assert a.shape_[-1] == b.shape_[-2], (a.shape_, b.shape_)
        lead = _bshape(T(a.shape_[:-2]), T(b.shape_[:-2]))
        return T(lead + (a.shape_[-2], b.shape_[-1]))

    tf.variable_scope, tf.get_variable, tf.reshape, tf.matmul = variable_scope, get_variable, reshape, matmul
    tf.extract_image_patches = extract_image_patches
    tf.shape = lambda x: [int(d) for d in x.shape_]
    tf.get_default_graph = lambda: object()
    tf.placeholder = lambda dtype, shape: T(shape)
    tf.scatter_update = lambda v, idx, val: v
    tf.group = lambda *a: None
    tf.expand_dims = lambda x, axis: T(x.shape_[:axis] + (1,) + x.shape_[axis:])
    tf.gather = lambda x, idx: x
    tf.nn = types.SimpleNamespace(relu=lambda x: x)
    return tf


def load_reference_models(tf):
    """neuroevolution/models/{base,dqn,batchnorm}.py as modules of a stand-in package (the real models/__init__.py also imports
    dqn_xavier / simple, which need tf.contrib initialisers)."""
    for k in [k for k in sys.modules if k == "refmodels" or k.startswith("refmodels.")]:
        del sys.modules[k]                              # a fresh import per model: the modules bind `tf` at import time
    sys.modules["tensorflow"] = tf
    sys.modules["tabular_logger"] = types.ModuleType("tabular_logger")
    gym_tf = types.ModuleType("gym_tensorflow")
    ops = types.ModuleType("gym_tensorflow.ops")
    ops.indexed_matmul = lambda *a, **kw: (_ for _ in ()).throw(AssertionError("indexed_matmul: indices is None in this run"))
    gym_tf.ops = ops
    sys.modules["gym_tensorflow"], sys.modules["gym_tensorflow.ops"] = gym_tf, ops
    pkg = types.ModuleType("refmodels")
    pkg.__path__ = [os.path.join(REF, "neuroevolution", "models")]
    sys.modules["refmodels"] = pkg
    return {m: importlib.import_module("refmodels." + m) for m in ("base", "dqn", "batchnorm")}


class Noise:
    """What the model methods use of SharedNoiseTable (es.py:51-67 / gpu_implementation/es.py): get(i, dim)."""
    def __init__(self, table):
        self.noise = table

    def get(self, i, dim):
        return self.noise[i:i + dim]


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import oracle as O
    count = 24_000_000
    noise = Noise(O.noise_table(count))
    out = {}
    for cls_name, mod in (("Model", "dqn"), ("LargeModel", "dqn"), ("ModelVirtualBN", "batchnorm")):
        tf = make_tf()
        mods = load_reference_models(tf)
        cls = getattr(mods[mod], cls_name)
        m = cls()
        x = T((1, 1, 84, 84, 4))                        # Policies x Batch x Height x Width x Feature (base.py:62)
        m.make_net(x, 18, batch_size=1, ref_batch=T((1, 84, 84, 4)) if m.requires_ref_batch else None)
        mods["base"].BaseModel.make_weights(m)          # the base part: num_params, scale_by (VBN's override re-traces the net)
        names = [v.name for v in m.variables]
        shapes = [list(v.shape_[1:]) for v in m.variables]
        out[f"{cls_name}.names"] = np.array(names)
        out[f"{cls_name}.shapes"] = np.array([",".join(map(str, s)) for s in shapes])
        out[f"{cls_name}.var_scale_by"] = np.array([float(v.scale_by) for v in m.variables], dtype=np.float64)
        out[f"{cls_name}.created"] = np.array(tf.created)          # every get_variable in order (incl. VBN mean / var)
        out[f"{cls_name}.num_params"] = np.int64(m.num_params)
        sb = np.asarray(m.scale_by)
        out[f"{cls_name}.scale_by_dtype"] = np.array(str(sb.dtype))
        out[f"{cls_name}.scale_by_sum"] = np.float64(sb.astype(np.float64).sum())
        P = int(m.num_params)
        rs = np.random.RandomState(2024)
        seeds = (int(rs.randint(0, count - P)),) + tuple((int(rs.randint(0, count - P)), p) for p in (0.002, 0.005, 0.002))
        theta = m.compute_weights_from_seeds(noise, seeds)
        out[f"{cls_name}.seed_idx"] = np.array([seeds[0]] + [s[0] for s in seeds[1:]], dtype=np.int64)
        out[f"{cls_name}.seed_power"] = np.array([0.0] + [s[1] for s in seeds[1:]], dtype=np.float64)
        out[f"{cls_name}.theta_dtype"] = np.array(str(theta.dtype))
        out[f"{cls_name}.theta_samples"] = np.asarray(theta[::997], dtype=np.float64)
        out[f"{cls_name}.theta_sum"] = np.float64(np.asarray(theta, dtype=np.float64).sum())
        out[f"{cls_name}.theta_sumsq"] = np.float64(np.square(np.asarray(theta, dtype=np.float64)).sum())
        print(cls_name, "P =", P, "vars:", list(zip(names, shapes)), "theta dtype", theta.dtype)
    np.savez_compressed(os.path.join(HERE, "ref_models.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_models.npz"))


if __name__ == "__main__":
    main()
