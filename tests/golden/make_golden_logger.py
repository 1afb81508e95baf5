"""Generate tests/golden/ref_logger.npz by EXECUTING the reference's tabular logger.

Run in the build container only (imports /root/reference):  python tests/golden/make_golden_logger.py

es_distributed/tabular_logger.py formats the per-iteration metric table (`record_tabular` / `dump_tabular`, :131-152) and free
text (`log`, :154-179) into <dir>/log.txt.  Only its TensorBoard writer needs TensorFlow; with stand-ins for those four imports
the text path runs unmodified.  The fixture holds the exact bytes of log.txt for the scripted calls of `script()`."""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def script(L):
    """The same calls are replayed on the package's logger by tests/test_host.py."""
    L.log("********** Iteration 3 **********")
    L.record_tabular("EpRewMean", 123.456789)
    L.record_tabular("EpLenMean", np.float32(812.5))
    L.record_tabular("EpCount", 2000)
    L.record_tabular("Norm", 1.5e-7)
    L.record_tabular("AVeryLongMetricNameThatExceedsThirtyThreeCharacters", 1.0)
    L.record_tabular("Big", 123456789.0)
    L.dump_tabular()
    L.log("two ", "parts")
    L.record_tabular("TimeElapsed", 98765.4321)
    L.record_tabular("Neg", -0.000123456)
    L.dump_tabular()


def main():
    tf = types.ModuleType("tensorflow")
    tf.Summary = type("Summary", (), {"__init__": lambda self, value=None: None, "Value": staticmethod(lambda **kw: None)})
    core, util, python, putil = (types.ModuleType(n) for n in ("tensorflow.core", "tensorflow.core.util", "tensorflow.python",
                                                               "tensorflow.python.util"))
    ev = types.ModuleType("tensorflow.core.util.event_pb2")
    ev.Event = lambda **kw: types.SimpleNamespace(step=0)
    pw = types.ModuleType("tensorflow.python.pywrap_tensorflow")
    pw.EventsWriter = lambda path: types.SimpleNamespace(WriteEvent=lambda e: None, Flush=lambda: None, Close=lambda: None)
    compat = types.ModuleType("tensorflow.python.util.compat")
    compat.as_bytes = lambda s: s.encode()
    util.event_pb2, python.pywrap_tensorflow, putil.compat = ev, pw, compat
    sys.modules.update({"tensorflow": tf, "tensorflow.core": core, "tensorflow.core.util": util, "tensorflow.core.util.event_pb2": ev,
                        "tensorflow.python": python, "tensorflow.python.pywrap_tensorflow": pw, "tensorflow.python.util": putil,
                        "tensorflow.python.util.compat": compat})
    spec = importlib.util.spec_from_file_location("ref_tabular_logger", "/root/reference/es_distributed/tabular_logger.py")
    L = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(L)
    with tempfile.TemporaryDirectory() as d:
        L.start(d)
        script(L)
        L.stop()
        data = open(os.path.join(d, "log.txt"), "rb").read()
    np.savez_compressed(os.path.join(HERE, "ref_logger.npz"), log_txt=np.frombuffer(data, dtype=np.uint8))
    print(data.decode())


if __name__ == "__main__":
    main()
