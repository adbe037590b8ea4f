"""CPU: k_small — the WHOLE latency path as one kernel (per-run cut loop, frame table, decode rounds with their shared-memory row staging, slot scan,
pack; the code k_ring runs per batch) — executed on an emulated thread block of 512 host threads (tests/cpp/block_emul_prelude.h: warp collectives
and __syncthreads() as barriers) out of the generated host-compilable copy of b2_kernels.cuh, and compared with the oracle exactly like the GPU
tests compare the device: run status, descriptors, reply bytes and the layout invariants of the reply region (tests/_compare.assert_same)."""
import ctypes as C
import os
import random
import subprocess
import sys
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import _oracle as O  # noqa: E402
from _compare import assert_same  # noqa: E402
from _traffic import SEED, echo_frame, echo_pb, mixed_frames, raw_request_frame, rnd62, split_runs  # noqa: E402
from brpc_b200.abi import ECHO_METHOD, MSG_DT, REF_DT, RUN_STATUS_DT  # noqa: E402
from brpc_b200.messenger import make_runs  # noqa: E402

LONG = os.environ.get("B2_LONG_TESTS") == "1"


@pytest.fixture(scope="module")
def sh():
    cpp = os.path.join(HERE, "cpp")
    so = os.path.join(cpp, "libsmall_host.so")
    deps = [os.path.join(cpp, f) for f in ("gen_kernels_host.py", "block_emul_prelude.h", "small_host.cc")] + \
           [os.path.join(ROOT, "brpc_b200", "csrc", f) for f in ("b2_kernels.cuh", "b2_core.cuh", "b2_inflate.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call([sys.executable, os.path.join(cpp, "gen_kernels_host.py")])
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-w", "-pthread", "-I", os.path.join(cpp, "stub"), "-I", os.path.join(ROOT, "include"),
                               "-o", so, os.path.join(cpp, "small_host.cc")])
    lib = C.CDLL(so)
    lib.sh_create.restype = C.c_void_p
    lib.sh_create.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p]
    lib.sh_destroy.argtypes = [C.c_void_p]
    lib.sh_add_method.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.sh_k_small.restype = C.c_uint32
    lib.sh_k_small.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    return lib


def make(sh, methods, identity=None, mask=(1 << 1) | (1 << 2), by_ref=0, stream_handler=0):
    k = sh.sh_create(0, mask, by_ref, stream_handler, identity)
    for m in methods:
        sh.sh_add_method(k, m["service_full_name"], m["service_name"], m["method_name"], m["request_type_name"], m["handler"], m["echo_attachment"],
                         m["response_checksum_type"], m["response_compress_type"])
    return k


def k_small(sh, k, chunks, preferred=-1, flags=0, max_msgs=1024):
    data, runs = make_runs(chunks)
    runs["preferred_proto"] = preferred; runs["flags"] = flags
    assert len(runs) <= 512 and len(data) <= 128 << 10                      # the shape the library sends down this path
    buf = np.concatenate([np.asarray(data, np.uint8), np.zeros(1024, np.uint8)])
    max_resp = len(data) + max_msgs * 80 + 2048                              # b2_batch_upload: small_resp
    rs = np.zeros(len(runs), RUN_STATUS_DT); msgs = np.zeros(max_msgs, MSG_DT); refs = np.zeros(max_msgs, REF_DT); resp = np.zeros(max_resp + 64, np.uint8)
    nm = C.c_uint32(); rb = C.c_uint32()
    flg = sh.sh_k_small(k, buf.ctypes.data, runs.ctypes.data, len(runs), max_msgs, max_resp, rs.ctypes.data, msgs.ctypes.data, refs.ctypes.data, resp.ctypes.data,
                        C.byref(nm), C.byref(rb))
    return flg, (rs, msgs[:nm.value], resp[:max(rb.value, 1)]), (data, runs)


def test_k_small_on_mixed_batches(sh):
    rng = random.Random(SEED + 981)
    total = 0
    for r_cks, r_cmp, identity in [(0, 0, None), (1, 0, b"10.0.0.1:8000"), (0, 1, None)]:
        ms = [dict(ECHO_METHOD, response_checksum_type=r_cks, response_compress_type=r_cmp)]
        k = make(sh, ms, identity); cfg = O.make_config(methods=ms, server_identity=identity)
        for trial in range(6 if LONG else 2):
            streams = [mixed_frames(rng, rng.randrange(1, 9)) for _ in range(rng.randrange(1, 40))]
            chunks = split_runs(rng, streams)
            for pref in ((-1, 1, 2) if LONG else (rng.choice([-1, 1, 2]),)):
                flg, dev, (data, runs) = k_small(sh, k, chunks, preferred=pref)
                assert flg == 0
                assert_same(dev, O.process_batch(cfg, data, runs), "k_small trial %d pref %d" % (trial, pref))
                total += len(dev[1])
        sh.sh_destroy(k)
    assert total > (1500 if LONG else 200)


def test_k_small_bench_shape_codecs_and_client_side(sh):
    rng = random.Random(SEED + 982)
    ms = [dict(ECHO_METHOD)]
    k = make(sh, ms); cfg = O.make_config(methods=ms)
    # 64 connections x 1 request: the batch bench.py times on the ring
    chunks = [echo_frame(rng, s, rnd62(rng, 1024)) for s in range(64)]
    flg, dev, (data, runs) = k_small(sh, k, chunks, preferred=1)
    assert flg == 0 and len(dev[1]) == 64
    assert_same(dev, O.process_batch(cfg, data, runs), "64 x 1")
    # snappy / gzip / zlib / CRC32C requests, some corrupted
    fr = []
    for i in range(24 if LONG else 12):
        msg = rnd62(rng, rng.choice([0, 10, 600, 3000]))
        kind = rng.choice(["snappy", "gzip", "zlib", "crc"])
        if kind in ("gzip", "zlib"):
            c = zlib.compressobj(rng.choice([0, 6]), zlib.DEFLATED, 31 if kind == "gzip" else 15)
            body = c.compress(echo_pb(msg)) + c.flush()
            f = raw_request_frame(body, 900 + i, compress_type=2 if kind == "gzip" else 3)
        else:
            f = echo_frame(rng, i, msg, compress_type=1 if kind == "snappy" else 0, checksum_type=1, attachment=rng.choice([b"", b"att"]))
        if rng.random() < 0.2:
            b = bytearray(f); b[-1 - rng.randrange(min(30, len(b) - 50))] ^= 0x11; f = bytes(b)
        fr.append(f)
    flg, dev, (data, runs) = k_small(sh, k, [b"".join(fr[i::4]) for i in range(4)])
    assert flg == 0
    assert_same(dev, O.process_batch(cfg, data, runs), "codecs")
    # the replies as client-side input
    orc = O.process_batch(cfg, data, runs)
    replies = [bytes(orc[2][int(m["resp_off"]):int(m["resp_off"]) + int(m["resp_len"])]) for m in orc[1] if int(m["resp_len"])]
    flg, dev, (data, runs) = k_small(sh, k, [b"".join(replies[i::3]) for i in range(3)], flags=1)
    assert flg == 0
    assert_same(dev, O.process_batch(cfg, data, runs), "client side")
    sh.sh_destroy(k)


def empty_reply_frames():
    """client-side replies whose EchoResponse is EMPTY (body b"" or 0a 00) under a CRC32C checksum, right and wrong: Crc32cVerify comes before
    the parse (DeserializeRpcMessage), so a wrong checksum fails the call although there is nothing to hand over"""
    from _traffic import raw_response_frame
    out = []
    for body in (b"", b"\x0a\x00", b"\x10\x05"):                 # nothing / an empty string / one unknown field (`message` is a required field: only the second parses)
        good = O.lib.orc_crc32c_mask(O.crc32c(body)).to_bytes(4, "big")
        bad = bytes([good[0] ^ 0x40]) + good[1:]
        for cks in (good, bad, b"", good + b"\x00"):
            out.append(raw_response_frame(body, 77 + len(out), checksum_type=1, checksum_value=cks))
        out.append(raw_response_frame(body, 99, checksum_type=1, checksum_value=bad, attachment=b"tail"))
    return out


def test_k_small_verifies_the_checksum_of_empty_replies(sh):
    k = make(sh, [dict(ECHO_METHOD)]); cfg = O.make_config()
    fr = empty_reply_frames()
    flg, dev, (data, runs) = k_small(sh, k, [b"".join(fr[i::2]) for i in range(2)], flags=1)
    assert flg == 0
    orc = O.process_batch(cfg, data, runs)
    assert_same(dev, orc, "empty replies")
    assert int((dev[1]["error_code"] == 1003).sum()) == 14 and int((dev[1]["error_code"] == 0).sum()) == 1     # only {0a 00} with the right checksum succeeds
    sh.sh_destroy(k)


def test_k_small_capacity_flags(sh):
    """more messages / reply bytes than the compact block holds: the kernel must flag it (the library then takes the tile pipeline), not overrun"""
    rng = random.Random(SEED + 983)
    k = make(sh, [dict(ECHO_METHOD)])
    chunks = [b"".join(echo_frame(rng, i, b"") for i in range(40)) for _ in range(4)]          # 160 tiny messages
    flg, dev, _ = k_small(sh, k, chunks, max_msgs=100)
    assert flg & 1
    sh.sh_destroy(k)
