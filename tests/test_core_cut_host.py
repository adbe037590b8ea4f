"""CPU: the PRODUCT's CutInputMessage chain — brpc_b200/csrc/b2_core.cuh, the source the kernels compile, built for the host
(tests/cpp/core_host.cc, a test harness) — against the oracle's restatement of input_messenger.cpp:84-179,206-322 on every run of mixed
five-protocol traffic: server and client sockets, every preferred index, handler masks, corruptions, truncation at random points, and
rpc_dump files.  The GPU tests compare the kernels with the oracle; this one runs the kernels' cut rules themselves where no GPU is."""
import ctypes as C
import os
import random
import struct
import subprocess

import numpy as np
import pytest

import _oracle as O
from _traffic import SEED, echo_frame, mixed_frames, rnd62

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ALL = (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4) | (1 << 12)


@pytest.fixture(scope="module")
def core():
    so = os.path.join(HERE, "cpp", "libcore_host.so")
    src = os.path.join(HERE, "cpp", "core_host.cc")
    hdr = os.path.join(ROOT, "brpc_b200", "csrc", "b2_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", so, src])
    lib = C.CDLL(so)
    lib.core_cut_run.restype = C.c_uint32
    lib.core_cut_run.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
    return lib


def hulu(meta, payload):
    return b"HULU" + struct.pack("<II", len(meta) + len(payload), len(meta)) + meta + payload


def sofa(meta, payload):
    return b"SOFA" + struct.pack("<IQQ", len(meta), len(payload), len(meta) + len(payload)) + meta + payload


def nshead(body, log_id=7):
    return struct.pack("<HHI16sIII", 1, 2, log_id, b"b2-test", 0xfb709394, 0, len(body)) + body


def five_protocol_stream(rng, n):
    out = []
    for i in range(n):
        c = rng.random()
        body = rnd62(rng, rng.choice([0, 1, 20, 100, 1000]))
        meta = rnd62(rng, rng.choice([0, 5, 40]))
        if c < 0.22: out.append(hulu(meta, body))
        elif c < 0.40: out.append(sofa(meta, body))
        elif c < 0.58: out.append(nshead(body, log_id=i))
        elif c < 0.84: out.append(echo_frame(rng, i, body))
        elif c < 0.88: out.append(b"HULU" + struct.pack("<II", 10, 50) + bytes(10))           # meta_size > body_size: popped, TRY_OTHERS
        elif c < 0.92: out.append(O.pack_stream_frame(rng.randrange(1 << 30), -1, 3, None, body))
        elif c < 0.95: out.append(sofa(meta, body)[:16] + struct.pack("<Q", 12345) + meta + body)   # msg_size mismatch
        elif c < 0.97: out.append(rnd62(rng, rng.randrange(1, 40)))                             # garbage: the connection dies here
        else: out.append(b"PRP" + rnd62(rng, 3))                                              # a magic prefix that goes wrong
    return out


def compare(core, chunks, mask, preferred=-1, flags=0, max_body=0):
    from brpc_b200.messenger import make_runs
    data, runs = make_runs(chunks)
    runs["preferred_proto"] = preferred; runs["flags"] = flags
    cfg = O.make_config(protocols=mask, max_body_size=max_body)
    rs, msgs, resp = O.process_batch(cfg, data, runs)
    buf = bytes(data)
    out = (C.c_uint32 * 3)(); offs = (C.c_uint32 * 4096)()
    n_total = 0
    for r in range(len(runs)):
        off, ln = int(runs["offset"][r]), int(runs["length"][r])
        n = core.core_cut_run(buf[off:off + ln], ln, preferred, max_body or (64 << 20), 1 if flags & 1 else 0, mask, flags, out, offs, 4096)
        got = (n, out[0], out[1], C.c_int32(out[2]).value)
        want = (int(rs["n_msgs"][r]), int(rs["consumed"][r]), int(rs["parse_error"][r]), int(rs["preferred_proto"][r]))
        assert got == want, ("run", r, got, want, mask, preferred, flags)
        f0 = int(rs["first_msg"][r])
        for k in range(min(n, 4096)):
            assert (offs[k] & 0x7fffffff) + off == int(msgs["frame_off"][f0 + k]), (r, k)
            if not flags & 4:                                                     # (a dump record's `protocol` is the sample's, not the handler's)
                assert (offs[k] >> 31) == (0 if int(msgs["protocol"][f0 + k]) == 1 else 1), (r, k, int(msgs["protocol"][f0 + k]))
        n_total += n
    return n_total


def test_cut_chain_on_baidu_std_and_streaming_traffic(core):
    rng = random.Random(SEED + 901)
    total = 0
    for trial in range(6):
        streams = [b"".join(mixed_frames(rng, rng.randrange(1, 60), big=trial == 0)) for _ in range(40)]
        chunks = [s[:rng.randrange(len(s) + 1)] if rng.random() < 0.5 else s for s in streams]
        for pref in (-1, 1, 2):
            total += compare(core, chunks, (1 << 1) | (1 << 2), preferred=pref)
        total += compare(core, chunks, (1 << 1) | (1 << 2), preferred=-1, flags=1)              # client sockets: the baidu_std <-> streaming fallback
        total += compare(core, chunks, (1 << 1) | (1 << 2), preferred=1, max_body=3000)         # TOO_BIG_DATA
    assert total > 5000


def test_cut_chain_on_five_protocols(core):
    rng = random.Random(SEED + 902)
    total = 0
    for trial in range(8):
        streams = [b"".join(five_protocol_stream(rng, rng.randrange(1, 50))) for _ in range(40)]
        chunks = [s[:rng.randrange(len(s) + 1)] if rng.random() < 0.5 else s for s in streams]
        for mask in (ALL, (1 << 1) | (1 << 3), (1 << 1) | (1 << 2) | (1 << 12), (1 << 4) | (1 << 12), 1 << 12):
            for pref in (-1, 1, 3, 4, 12):
                total += compare(core, chunks, mask, preferred=pref)
    assert total > 10000


def test_cut_chain_on_rpc_dump_files(core):
    import json
    vec = json.load(open(os.path.join(HERE, "golden", "dump_vectors.json")))
    files = [bytes.fromhex(v["file_hex"]) for v in vec["files"]]
    rng = random.Random(SEED + 903)
    chunks = files + [f[:rng.randrange(len(f) + 1)] for f in files] + [f[:10] + b"\xff" + f[11:] for f in files if len(f) > 20]
    assert compare(core, chunks, (1 << 1) | (1 << 2), preferred=-1, flags=4) > 0


def test_product_decoders_against_the_oracle_on_golden_and_fuzz_vectors(core):
    """decode_rpc_meta / decode_stream_meta / decode_echo_request of b2_core.cuh (host build) field by field against the oracle, which the
    python-protobuf vectors pin (tests/test_oracle_golden.py): 3 000+ valid, mutated and truncated metas."""
    import json
    core.core_decode_rpc_meta.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_void_p]
    core.core_decode_stream_meta.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p]
    core.core_decode_echo_request.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    gold = lambda n: json.load(open(os.path.join(HERE, "golden", n)))
    metas = [bytes.fromhex(r["hex"]) for r in gold("rpc_meta_vectors.json")["rpc_meta"]] + [bytes.fromhex(r["hex"]) for r in gold("pb_fuzz_vectors.json")["rpc_meta"]]
    rng = random.Random(SEED + 904)
    metas += [m[:rng.randrange(len(m) + 1)] for m in metas[:600]]
    out = (C.c_longlong * 14)()
    n_ok = 0
    for b in metas:
        ok_o, m = O.parse_rpc_meta(b)
        ok_p = core.core_decode_rpc_meta(b, len(b), 0, out)
        assert bool(ok_p) == ok_o, b.hex()
        if not ok_o:
            continue
        n_ok += 1
        assert out[0] == m.has and out[1] == m.correlation_id and out[3] == m.compress_type and out[4] == m.attachment_size, b.hex()
        assert out[5] == m.checksum_type and out[6] == m.content_type and out[13] == m.error_code, b.hex()
        if m.has & (1 << 12): assert out[2] == m.log_id
        if m.has_service_name: assert (out[7], out[8]) == (m.service_name.off, m.service_name.len)
        if m.has_method_name: assert (out[9], out[10]) == (m.method_name.off, m.method_name.len)
        if m.has & (1 << 11): assert (out[11], out[12]) == (m.checksum_value.off, m.checksum_value.len)
    assert n_ok > 1000
    frames = [bytes.fromhex(r["hex"]) for r in gold("pb_fuzz_vectors.json")["stream_frame_meta"]]
    frames += [f[:rng.randrange(len(f) + 1)] for f in frames[:300]]
    so = (C.c_longlong * 5)()
    for b in frames:
        ok_o, m = O.parse_stream_meta(b)
        assert bool(core.core_decode_stream_meta(b, len(b), so)) == ok_o, b.hex()
        if ok_o:
            assert (so[0], so[1], so[4]) == (m.has, m.stream_id, m.frame_type) and so[2] == m.source_stream_id and so[3] == m.consumed_size, b.hex()
    off = C.c_uint32(); ln = C.c_uint32()
    bodies = [b"", b"\x0a\x00", b"\x0a\x03abc", b"\x0a\x03ab", b"\x12\x01x\x0a\x02hi", b"\x0a\x01a\x0a\x02bc", b"\x08\x01", b"\x0a\x80\x01" + b"x" * 128, b"\x0a\xff\xff\xff\xff\x0f",
              b"\x0b\x0c", b"\x0a\x02hi\x00"] + [bytes(rng.getrandbits(8) for _ in range(rng.randrange(0, 12))) for _ in range(2000)]
    for b in bodies:
        ok_o, (o_off, o_len) = O.parse_echo_request(b)
        assert bool(core.core_decode_echo_request(b, len(b), C.byref(off), C.byref(ln))) == ok_o, b.hex()
        if ok_o:
            assert (off.value, ln.value) == (o_off, o_len), b.hex()


def test_cut_chain_and_decoders_on_arbitrary_bytes(core):
    """hypothesis: ANY byte string as a run / as a meta — seeded with the magics and header shapes so that the interesting branches are
    reached — gives the same verdicts from the product's rules and the oracle's."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    core.core_decode_rpc_meta.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_void_p]
    magic = st.sampled_from([b"PRPC", b"STRM", b"HULU", b"SOFA", b"PRP", b"ST", b"\xfb\x70\x93\x94", b"", b"P", b"HUL"])
    u32 = st.integers(0, 2 ** 32 - 1)
    small = st.integers(0, 80)
    def frame_like(m, a, b, tail):
        return m + struct.pack(">II", a, b) + tail
    piece = st.one_of(st.binary(max_size=48),
                      st.builds(frame_like, magic, small, small, st.binary(max_size=96)),
                      st.builds(frame_like, magic, u32, u32, st.binary(max_size=16)),
                      st.builds(lambda m, a, b, t: m + struct.pack("<II", a, b) + t, magic, small, small, st.binary(max_size=96)),
                      st.builds(lambda body: struct.pack("<HHI16sIII", 1, 2, 3, b"x", 0xfb709394, 0, len(body)) + body, st.binary(max_size=40)))
    runs = st.lists(piece, min_size=0, max_size=6).map(b"".join)

    @settings(max_examples=400, deadline=None, suppress_health_check=list(HealthCheck))
    @given(runs, st.sampled_from([-1, 1, 2, 3, 4, 12]), st.sampled_from([ALL, (1 << 1) | (1 << 2), (1 << 1) | (1 << 12), 1 << 3]), st.sampled_from([0, 1]), st.sampled_from([0, 50]))
    def cut(run, pref, mask, flags, max_body):
        compare(core, [run], mask, preferred=pref, flags=flags, max_body=max_body)
    cut()

    out = (C.c_longlong * 14)()

    @settings(max_examples=600, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.lists(st.one_of(st.binary(max_size=12), st.sampled_from([b"\x0a", b"\x12", b"\x18", b"\x20", b"\x28", b"\x32", b"\x3a", b"\x42", b"\x4a", b"\x50", b"\x58", b"\x62",
                                                                          b"\x0a\x02\x0a\x00", b"\x12\x02\x08\x00", b"\xff\xff\xff\xff\x0f", b"\x80"])), max_size=12).map(b"".join))
    def meta(b):
        ok_o, m = O.parse_rpc_meta(b)
        assert bool(core.core_decode_rpc_meta(b, len(b), 0, out)) == ok_o, b.hex()
        if ok_o:
            assert out[0] == m.has and out[1] == m.correlation_id and out[3] == m.compress_type and out[13] == m.error_code, b.hex()
    meta()
