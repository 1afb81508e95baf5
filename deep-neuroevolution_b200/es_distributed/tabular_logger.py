"""Key-compatible stand-in for the reference's ``tabular_logger`` (es_distributed/tabular_logger.py:60-153):
``start / log / record_tabular / dump_tabular / stop``; rows go to stdout and ``<dir>/log.txt`` in the same
boxed layout.  No TensorFlow event writer ("next")."""
from __future__ import annotations

import os
import sys
import time
from collections import OrderedDict

_state = {"dir": None, "file": None, "row": OrderedDict(), "tstart": time.time(), "quiet": False, "rows": []}


def start(dir):
    stop()
    _state["dir"] = dir
    if dir:
        os.makedirs(dir, exist_ok=True)
        _state["file"] = open(os.path.join(dir, "log.txt"), "at")
    _state["tstart"] = time.time()


def stop():
    if _state["file"]:
        _state["file"].close()
    _state["file"] = None
    _state["dir"] = None


def set_quiet(q=True):
    _state["quiet"] = q


def log(*args):
    msg = " ".join(str(a) for a in args)
    if not _state["quiet"]:
        sys.stdout.write(msg + "\n")
        sys.stdout.flush()
    if _state["file"]:
        _state["file"].write(msg + "\n")
        _state["file"].flush()


def record_tabular(key, val):
    _state["row"][key] = val


def dump_tabular():
    row = _state["row"]
    if not row:
        return
    items = [(k, ("%-8.3g" % v) if hasattr(v, "__float__") else str(v)) for k, v in row.items()]
    kw = max(len(k) for k, _ in items)
    vw = max(len(v) for _, v in items)
    dashes = "-" * (kw + vw + 7)
    lines = [dashes] + ["| %s%s | %s%s |" % (k, " " * (kw - len(k)), v, " " * (vw - len(v))) for k, v in items] + [dashes]
    log("\n".join(lines))
    _state["rows"].append(dict(row))
    row.clear()


def history():
    """All dumped rows (tests / bench read the metrics back from here)."""
    return _state["rows"]
