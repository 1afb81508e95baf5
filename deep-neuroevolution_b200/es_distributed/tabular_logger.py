"""Key-compatible stand-in for the reference's ``tabular_logger`` (es_distributed/tabular_logger.py:60-153):
``start / log / record_tabular / dump_tabular / stop``; rows go to stdout and ``<dir>/log.txt`` in the same
boxed layout, and -- like the reference's ``TbWriter`` (tabular_logger.py:17-52) -- every dumped row is appended to a
TensorBoard event file ``<dir>/events.out.tfevents.*`` as scalar summaries.  The event file is written without TensorFlow:
TFRecord framing (length, masked CRC-32C of the length, payload, masked CRC-32C of the payload) around hand-encoded
``Event { wall_time, step, summary { value { tag, simple_value } } }`` protobuf messages."""
from __future__ import annotations

import os
import sys
import time
from collections import OrderedDict

_state = {"dir": None, "file": None, "row": OrderedDict(), "tstart": time.time(), "quiet": False, "rows": [], "tb": None}


# ---- TensorBoard event file (no TensorFlow dependency) -----------------------------------------------------------------
def _crc32c_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TAB = _crc32c_table()


def _crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TAB[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked_crc(data: bytes) -> int:
    c = _crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _pb_bytes(field: int, payload: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


class TbWriter(object):
    """tabular_logger.py:17-52: ``write_values(key2val)`` appends one Event with a scalar summary per key; step counts the
    calls starting at 1."""

    def __init__(self, dir, prefix="events"):
        import socket
        import struct
        self._struct = struct
        self.dir, self.step = dir, 1
        self.path = os.path.join(os.path.abspath(dir), "%s.out.tfevents.%010d.%s" % (prefix, int(time.time()), socket.gethostname()))
        self.f = open(self.path, "wb")
        self._write(struct.pack("<Bd", 0x09, time.time()) + _pb_bytes(3, b"brain.Event:2"))     # file_version record

    def _write(self, event: bytes):
        st = self._struct
        header = st.pack("<Q", len(event))
        self.f.write(header + st.pack("<I", _masked_crc(header)) + event + st.pack("<I", _masked_crc(event)))
        self.f.flush()

    def write_values(self, key2val):
        st = self._struct
        values = b""
        for k, v in key2val.items():
            if not hasattr(v, "__float__"):
                continue
            val = _pb_bytes(1, str(k).encode()) + st.pack("<Bf", 0x15, float(v))                  # tag = 1, simple_value = 2
            values += _pb_bytes(1, val)
        event = st.pack("<Bd", 0x09, time.time()) + b"\x10" + _varint(self.step) + _pb_bytes(5, values)
        self._write(event)
        self.step += 1

    def close(self):
        if self.f:
            self.f.close()
            self.f = None


def read_tb_events(path):
    """Parse an event file written by TbWriter back into [(step, {tag: value})] (tests; checks both CRCs)."""
    import struct
    out = []
    data = open(path, "rb").read()
    pos = 0

    def rd_varint(buf, i):
        n = shift = 0
        while True:
            b = buf[i]
            i += 1
            n |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                return n, i

    def fields(buf):
        i = 0
        while i < len(buf):
            key, i = rd_varint(buf, i)
            f, wt = key >> 3, key & 7
            if wt == 0:
                v, i = rd_varint(buf, i)
            elif wt == 1:
                v, i = buf[i:i + 8], i + 8
            elif wt == 5:
                v, i = buf[i:i + 4], i + 4
            else:
                n, i = rd_varint(buf, i)
                v, i = buf[i:i + n], i + n
            yield f, wt, v
    while pos < len(data):
        (n,) = struct.unpack_from("<Q", data, pos)
        assert struct.unpack_from("<I", data, pos + 8)[0] == _masked_crc(data[pos:pos + 8])
        ev = data[pos + 12:pos + 12 + n]
        assert struct.unpack_from("<I", data, pos + 12 + n)[0] == _masked_crc(ev)
        pos += 16 + n
        step, vals = 0, {}
        for f, wt, v in fields(ev):
            if f == 2:
                step = v
            elif f == 5:
                for f2, _, val in fields(v):
                    if f2 == 1:
                        d = {ff: vv for ff, _, vv in fields(val)}
                        vals[d[1].decode()] = struct.unpack("<f", d[2])[0]
        if vals:
            out.append((step, vals))
    return out


def start(dir):
    stop()
    _state["dir"] = dir
    if dir:
        os.makedirs(dir, exist_ok=True)
        _state["file"] = open(os.path.join(dir, "log.txt"), "at")
        _state["tb"] = TbWriter(dir=dir, prefix="events")                      # tabular_logger.py:123
    _state["tstart"] = time.time()


def stop():
    if _state["file"]:
        _state["file"].close()
    if _state["tb"]:
        _state["tb"].close()
    _state["file"] = _state["tb"] = None
    _state["dir"] = None


def set_quiet(q=True):
    _state["quiet"] = q


def log(*args):
    # tabular_logger.py:171-179: the strings are written back to back (no separator), then a newline; non-string arguments
    # (an extension: the reference's f.write would raise) are joined with spaces
    msg = "".join(args) if all(isinstance(a, str) for a in args) else " ".join(str(a) for a in args)
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
    def trunc(t):                                                       # tabular_logger.py:180-184
        return t[:30] + "..." if len(t) > 33 else t
    items = [(trunc(k), trunc(("%-8.3g" % v) if hasattr(v, "__float__") else str(v))) for k, v in row.items()]
    kw = max(len(k) for k, _ in items)
    vw = max(len(v) for _, v in items)
    dashes = "-" * (kw + vw + 7)
    lines = [dashes] + ["| %s%s | %s%s |" % (k, " " * (kw - len(k)), v, " " * (vw - len(v))) for k, v in items] + [dashes]
    log("\n".join(lines))
    _state["rows"].append(dict(row))
    if _state["tb"]:
        _state["tb"].write_values(row)                                        # tabular_logger.py:150-152
    row.clear()


def history():
    """All dumped rows (tests / bench read the metrics back from here)."""
    return _state["rows"]
