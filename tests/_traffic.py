"""Seeded byte-stream builders for the parity tests (uses the ORACLE's packers)."""
import random

import numpy as np

import _oracle as O

SEED = 20260921
T62 = b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"


def rnd62(rng, n):
    return bytes(rng.choice(T62) for _ in range(n))


def echo_frame(rng, i, payload=None, **kw):
    if payload is None:
        payload = b"r" * rng.choice([0, 1, 11, 16, 64, 100, 127, 128, 1024])
    args = dict(log_id=i & 0x3fff, correlation_id=((i & 0xfffff) << 32) | (i % 7 + 1), message=payload)
    args.update(kw)
    return O.pack_echo_request(**args)


def mixed_frames(rng, n, big=False):
    """A list of frames exercising every branch of ProcessRpcRequest + stream frames."""
    out = []
    for i in range(n):
        c = rng.random()
        size = rng.choice([0, 1, 5, 16, 63, 64, 65, 127, 128, 129, 1000, 1024, 4096] + ([16384, 70000] if big else []))
        msg = rnd62(rng, size) if rng.random() < 0.5 else b"r" * size
        if c < 0.45:
            out.append(echo_frame(rng, i, msg))
        elif c < 0.55:
            out.append(echo_frame(rng, i, msg, attachment=rnd62(rng, rng.choice([1, 7, 16, 100, 1000]))))
        elif c < 0.62:
            out.append(echo_frame(rng, i, msg, checksum_type=1))
        elif c < 0.66:
            out.append(echo_frame(rng, i, msg, checksum_type=1, attachment=rnd62(rng, 33)))
        elif c < 0.70:
            out.append(echo_frame(rng, i, msg, service=b"EchoService"))                       # short service name
        elif c < 0.73:
            out.append(echo_frame(rng, i, msg, service=b"NoSuchService"))                     # ENOSERVICE
        elif c < 0.76:
            out.append(echo_frame(rng, i, msg, service=b"example.Nope", method=b"Echo"))      # ENOMETHOD
        elif c < 0.78:
            out.append(echo_frame(rng, i, msg, method=b"Echo2"))
        elif c < 0.80:
            out.append(echo_frame(rng, i, msg, log_id=None, correlation_id=rng.randrange(-(1 << 63), 1 << 63)))
        elif c < 0.83:
            out.append(echo_frame(rng, i, msg, trace=(rng.randrange(1 << 62), 5, 6), request_id=b"req-%d" % i, timeout_ms=rng.randrange(1, 100000)))
        elif c < 0.85:
            out.append(echo_frame(rng, i, msg, compress_type=rng.choice([2, 3])))            # gzip/zlib: out of scope
        elif c < 0.87:
            out.append(echo_frame(rng, i, msg, compress_type=rng.choice([4, 9, 2000, -3])))  # no handler -> EREQUEST
        elif c < 0.89:
            out.append(echo_frame(rng, i, msg, content_type=rng.choice([1, 2, 3])))           # json: out of scope
        elif c < 0.91:
            out.append(echo_frame(rng, i, msg, checksum_type=rng.choice([2, 77, -1])))        # unknown checksum: accepted
        elif c < 0.93:
            f = bytearray(echo_frame(rng, i, msg, checksum_type=1))
            if len(f) > 70:
                f[-1] ^= 0x5a                                                                 # CRC mismatch -> EREQUEST
            out.append(bytes(f))
        elif c < 0.97:
            out.append(O.pack_stream_frame(rng.randrange(1 << 40), rng.choice([-1, rng.randrange(1 << 40)]),
                                           rng.randrange(0, 5), rng.choice([None, True, False]), rnd62(rng, rng.choice([0, 10, 500]))))
        else:
            f = bytearray(echo_frame(rng, i, msg))
            meta_size = int.from_bytes(f[8:12], "big")
            pos = 12 + rng.randrange(meta_size)
            f[pos] = rng.randrange(256)                                                       # corrupt the meta
            out.append(bytes(f))
    return out


def patch_meta(frame, fn):
    """Rebuild a frame with its RpcMeta bytes replaced by fn(meta)."""
    body = int.from_bytes(frame[4:8], "big"); meta = int.from_bytes(frame[8:12], "big")
    m = fn(frame[12:12 + meta])
    rest = frame[12 + meta:12 + body]
    return b"PRPC" + (len(m) + len(rest)).to_bytes(4, "big") + len(m).to_bytes(4, "big") + m + rest


def split_runs(rng, streams, cut_tail=True):
    """Each stream (list of frames) becomes one run; optionally cut at a random point of the last frame."""
    chunks = []
    for frames in streams:
        b = b"".join(frames)
        if cut_tail and frames and rng.random() < 0.7:
            cut = rng.randrange(0, len(frames[-1]))
            b = b[:len(b) - len(frames[-1]) + cut]
        chunks.append(b)
    return chunks


# ---- hand-built baidu_std frames (any body bytes, any compress_type): what a peer using another codec puts on the wire ----
def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80); v >>= 7
    out.append(v)
    return bytes(out)


def _fld(num, wt, payload):
    return _varint((num << 3) | wt) + payload


def echo_pb(message):
    """EchoRequest / EchoResponse {message}"""
    return _fld(1, 2, _varint(len(message)) + message)


def raw_request_frame(body, correlation_id, compress_type=0, checksum_value=None, checksum_type=0, attachment=b"",
                      service=b"example.EchoService", method=b"Echo", log_id=None):
    """PackRpcRequest's frame (baidu_rpc_protocol.cpp:1045-1133) around an already serialized (compressed) body."""
    req = _fld(1, 2, _varint(len(service)) + service) + _fld(2, 2, _varint(len(method)) + method)
    if log_id is not None:
        req += _fld(3, 0, _varint(log_id))
    meta = _fld(1, 2, _varint(len(req)) + req) + _fld(3, 0, _varint(compress_type)) + _fld(4, 0, _varint(correlation_id))
    if attachment:
        meta += _fld(5, 0, _varint(len(attachment)))
    meta += _fld(10, 0, _varint(0)) + _fld(11, 0, _varint(checksum_type))
    meta += _fld(12, 2, _varint(len(checksum_value or b"")) + (checksum_value or b""))
    return b"PRPC" + (len(meta) + len(body) + len(attachment)).to_bytes(4, "big") + len(meta).to_bytes(4, "big") + meta + body + attachment


def raw_response_frame(body, correlation_id, compress_type=0, checksum_value=None, checksum_type=0, attachment=b""):
    """SendRpcResponse's frame (baidu_rpc_protocol.cpp:339-352) around an already serialized (compressed) body."""
    resp = _fld(1, 0, _varint(0))
    meta = _fld(2, 2, _varint(len(resp)) + resp) + _fld(3, 0, _varint(compress_type)) + _fld(4, 0, _varint(correlation_id))
    if attachment:
        meta += _fld(5, 0, _varint(len(attachment)))
    meta += _fld(10, 0, _varint(0)) + _fld(11, 0, _varint(checksum_type))
    meta += _fld(12, 2, _varint(len(checksum_value or b"")) + (checksum_value or b""))
    return b"PRPC" + (len(meta) + len(body) + len(attachment)).to_bytes(4, "big") + len(meta).to_bytes(4, "big") + meta + body + attachment
