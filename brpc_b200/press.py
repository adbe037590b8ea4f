"""ctypes binding of tools/libb2press.so — the synthetic rpc_press traffic source."""
import ctypes as C
import os
import subprocess

import numpy as np

from .abi import RUN_DT

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "tools", "libb2press.so")


class Spec(C.Structure):
    _fields_ = [("service", C.c_char_p), ("method", C.c_char_p), ("payload_bytes", C.c_uint32),
                ("attachment_bytes", C.c_uint32), ("payload_kind", C.c_int32), ("checksum_type", C.c_int32),
                ("seed", C.c_uint64)]


def _load():
    if not os.path.exists(_SO):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _SO,
                               os.path.join(_HERE, "tools", "rpc_press.cpp")])
    l = C.CDLL(_SO)
    l.b2press_frame.restype = C.c_size_t
    l.b2press_frame.argtypes = [C.POINTER(Spec), C.c_uint64, C.c_void_p, C.c_size_t]
    l.b2press_fill_run.restype = C.c_uint64
    l.b2press_fill_run.argtypes = [C.POINTER(Spec), C.POINTER(C.c_uint64), C.c_void_p, C.c_size_t]
    return l


lib = _load()


def spec(payload_bytes=1024, attachment_bytes=0, payload_kind=0, checksum_type=0, seed=20260921,
         service=b"example.EchoService", method=b"Echo"):
    return Spec(service, method, payload_bytes, attachment_bytes, payload_kind, checksum_type, seed)


def frame(sp, index):
    cap = 2048 + sp.payload_bytes + sp.attachment_bytes
    buf = C.create_string_buffer(cap)
    n = lib.b2press_frame(C.byref(sp), index, buf, cap)
    return buf.raw[:n]


def fill_batch(sp, out, n_sockets, run_bytes, start_index=0):
    """Fill `out` (np.uint8 array, >= n_sockets * stride) with one run per socket; runs are 16 B aligned.
    Socket s streams frames s*2^32 + start_index ....  Returns (runs, n_complete_frames)."""
    stride = (run_bytes + 15) // 16 * 16
    assert out.nbytes >= n_sockets * stride
    runs = np.zeros(n_sockets, dtype=RUN_DT)
    total = 0
    base = out.ctypes.data
    for s in range(n_sockets):
        idx = C.c_uint64((s << 32) + start_index)
        total += lib.b2press_fill_run(C.byref(sp), C.byref(idx), base + s * stride, run_bytes)
        runs[s] = (s, s * stride, run_bytes, -1, 0)
    return runs, total


def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80); v >>= 7
    out.append(v)
    return bytes(out)


def stream_frame(stream_id, source_stream_id, data=b"", frame_type=3, has_continuation=None):
    """Client mirror of PackStreamMessage (policy/streaming_rpc_protocol.cpp:42-58) with the StreamFrameMeta
    fields Stream::AppendIfNotFull/Write sets on DATA frames (stream.cpp:181-186,199-204):
    "STRM" + BE32(meta+data) + BE32(meta) + StreamFrameMeta + data."""
    meta = b"\x08" + _varint(stream_id) + b"\x10" + _varint(source_stream_id) + b"\x18" + _varint(frame_type)
    if has_continuation is not None:
        meta += b"\x20" + (b"\x01" if has_continuation else b"\x00")
    return b"STRM" + (len(meta) + len(data)).to_bytes(4, "big") + len(meta).to_bytes(4, "big") + meta + bytes(data)
