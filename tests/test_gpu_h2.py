"""GPU: h2 building blocks (a15).  HPACK decode vs the RFC 7541 Appendix C vectors the reference asserts
(test/brpc_hpack_unittest.cpp) and vs the oracle on generated and mutated header blocks with per-connection
dynamic tables; h2 frame-head scan vs the oracle on gRPC-shaped connection streams (example/grpc_c++:
preface + SETTINGS, then HEADERS + DATA(5-byte gRPC prefix + 4 KB pb) per call)."""
import json
import os
import random
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
import _oracle as O  # noqa: E402


def hp_int(v, prefix, first=0):
    m = (1 << prefix) - 1
    if v < m:
        return bytes([first | v])
    out = [first | m]; v -= m
    while v >= 128:
        out.append((v & 0x7f) | 0x80); v >>= 7
    out.append(v)
    return bytes(out)


def hp_str(s):
    return hp_int(len(s), 7) + s


def gen_block(rng, names):
    """A header block using indexed, incremental, non-indexed and never-indexed forms + size updates."""
    b = bytearray()
    for _ in range(rng.randrange(1, 12)):
        c = rng.random()
        if c < 0.25: b += hp_int(rng.randrange(1, 62), 7, 0x80)
        elif c < 0.35: b += hp_int(rng.randrange(62, 70), 7, 0x80)                       # may or may not exist
        elif c < 0.6: b += hp_int(0, 6, 0x40) + hp_str(rng.choice(names)) + hp_str(os.urandom(rng.randrange(0, 40)))
        elif c < 0.7: b += hp_int(rng.randrange(1, 62), 6, 0x40) + hp_str(os.urandom(rng.randrange(0, 300)))
        elif c < 0.8: b += hp_int(rng.randrange(0, 62), 4, 0x00) + (hp_str(b"X-Mixed-Case") if rng.random() < 0.5 else b"") + hp_str(b"v" * rng.randrange(0, 20))
        elif c < 0.9: b += hp_int(0, 4, 0x10) + hp_str(b"secret-%d" % rng.randrange(10)) + hp_str(os.urandom(8))
        else: b += hp_int(rng.choice([0, 100, 4096, 4097, 256]), 5, 0x20)
    return bytes(b)


def test_hpack_rfc_vectors_and_generated_blocks(oracle):
    import brpc_b200
    ctx = brpc_b200.Context(device=0, max_batch_bytes=8 << 20, max_msgs=1 << 16, max_runs=64, max_resp_bytes=64 << 20)
    vec = json.load(open(os.path.join(HERE, "golden", "hpack_vectors.json")))
    # RFC 7541 C.2-C.6: one connection per test, steps in order
    buf = bytearray(); blocks = []; want = []
    for conn, t in enumerate(vec["unittest"]):
        ctx.hpack_reset(conn, t["max_table_size"])
        for s in t["steps"]:
            b = bytes.fromhex(s["bytes_hex"])
            blocks.append((conn, len(buf), len(b))); buf += b
            want.append((0, [(a.lower().encode(), v.encode()) for a, v in s["headers"]]))
    got = ctx.hpack_decode_batch(np.frombuffer(bytes(buf), np.uint8), blocks)
    assert got == want
    # generated + mutated blocks on 300 connections, several rounds: the dynamic tables persist across calls
    rng = random.Random(20260921)
    names = [b"custom-key", b"x-trace", b"Grpc-Timeout", b"te", b"a", b"user-agent-extra-long-name-0123456789"]
    n_conn = 300
    orcs = [oracle.HPack(4096) for _ in range(n_conn)]
    for c in range(n_conn): ctx.hpack_reset(1000 + c, 4096)
    seeds = [bytes.fromhex(s["bytes_hex"]) for t in vec["unittest"] for s in t["steps"]] + [bytes.fromhex(s["hex"]) for s in vec["seed_corpus"]]
    for rnd in range(4):
        buf = bytearray(); blocks = []; want = []
        for c in range(n_conn):
            for _ in range(rng.randrange(1, 4)):
                b = gen_block(rng, names) if rng.random() < 0.7 else bytearray(rng.choice(seeds))
                if rng.random() < 0.2 and len(b):
                    b = bytearray(b); b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
                b = bytes(b)
                blocks.append((1000 + c, len(buf), len(b))); buf += b
                want.append(orcs[c].decode_block(b))
        got = ctx.hpack_decode_batch(np.frombuffer(bytes(buf) + b"\0", np.uint8), blocks, per_block_cap=8192)
        bad = [i for i in range(len(want)) if got[i] != want[i]]
        assert not bad, (rnd, bad[:3], got[bad[0]], want[bad[0]])
        assert sum(1 for w in want if w[0] == -1) > 5 and sum(1 for w in want if w[0] == 0) > 100


def grpc_connection(rng, calls, preface=True):
    def frame(t, flags, sid, payload):
        return struct.pack(">I", len(payload))[1:] + bytes([t, flags]) + struct.pack(">I", sid) + payload
    b = bytearray()
    if preface:
        b += b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n" + frame(4, 0, 0, b"\x00\x03\x00\x00\x00\x64")
    sid = 1
    for _ in range(calls):
        hdr = b"\x83\x86\x44\x1e/helloworld.Greeter/SayHello" + os.urandom(rng.randrange(0, 30))
        b += frame(1, 4, sid, hdr)
        pb = b"\x0a\x80\x20" + b"n" * 4096
        data = b"\x00" + struct.pack(">I", len(pb)) + pb
        while data:
            chunk, data = data[:16384], data[16384:]
            b += frame(0, 1 if not data else 0, sid, chunk)
        sid += 2
    return bytes(b)


def test_h2_frame_scan_vs_oracle(oracle):
    import brpc_b200
    ctx = brpc_b200.Context(device=0, max_batch_bytes=64 << 20, max_msgs=1 << 16, max_runs=1024, max_resp_bytes=64 << 20)
    rng = random.Random(4)
    chunks, flags = [], []
    for c in range(256):
        pre = rng.random() < 0.5
        b = bytearray(grpc_connection(rng, rng.randrange(1, 6), preface=pre))
        k = rng.random()
        if k < 0.2: b = b[:rng.randrange(0, len(b))]                                # cut mid-frame
        elif k < 0.3: b[rng.randrange(min(len(b), 24))] ^= 0x55                       # broken preface or first head
        elif k < 0.4 and len(b) > 60: b[len(b) // 2] = 0xff                           # garbage in the middle
        chunks.append(bytes(b)); flags.append(2 if pre else 0)
    chunks += [b"", b"PRI", b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n", b"\x00\x40\x01\x00\x00\x00\x00\x00\x01", b"\x00\x00\x00\x00\x00\x80\x00\x00\x01"]
    flags += [2, 2, 2, 0, 0]
    data, runs = brpc_b200.make_runs(chunks)
    runs["flags"] = flags
    frames, nf, cons, err = ctx.h2_scan_batch(data, runs, max_frame_size=16384, cap_per_run=64)
    for i, ch in enumerate(chunks):
        body = ch
        if flags[i] & 2:
            pre = b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n"
            if ch[:24] != pre[:len(ch[:24])]:
                assert (nf[i], cons[i], err[i]) == (0, 0, 1); continue
            if len(ch) < 24:
                assert (nf[i], cons[i], err[i]) == (0, 0, 2); continue
            body = ch[24:]
        ofr, oc, oe = oracle.h2_scan(body, 16384)
        shift = len(ch) - len(body)
        assert nf[i] == len(ofr) and cons[i] == oc + shift and err[i] == oe, i
        d = frames[i]
        assert np.array_equal(d["type"], ofr["type"][:64]) and np.array_equal(d["flags"], ofr["flags"][:64])
        assert np.array_equal(d["stream_id"], ofr["stream_id"][:64]) and np.array_equal(d["payload_len"], ofr["payload_len"][:64])
        assert np.array_equal(d["payload_off"].astype(np.int64) - int(runs["offset"][i]) - shift, ofr["payload_off"][:64].astype(np.int64))
    assert (err == 5).sum() > 3 and (err == 2).sum() > 100
