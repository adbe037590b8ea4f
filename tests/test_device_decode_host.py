"""CPU: the per-message device functions of the hot path — decode_one (what a lane of k_decode / k_small / k_ring runs) and fused_fast_echo
(what a lane of k_fused runs for the exact PackRpcRequest shape) — from a host-compilable copy of brpc_b200/csrc/b2_kernels.cuh
(tests/cpp/gen_kernels_host.py: inline PTX dropped, a warp of one lane), message by message against the oracle: descriptor fields, error
classes, reply lengths, the reply prefix of the bandwidth path, and the whole in-place reply of the fused path."""
import ctypes as C
import os
import random
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import _oracle as O  # noqa: E402
from _traffic import SEED, echo_frame, mixed_frames, rnd62, split_runs  # noqa: E402
from brpc_b200.abi import ECHO_METHOD, MSG_DT, RUN_DT  # noqa: E402
from brpc_b200.messenger import make_runs  # noqa: E402

AUX_DT = np.dtype([(n, "<u4") for n in ("msg_off", "msg_len", "att_len", "att_off", "cks_off", "cks_len", "svc_off", "svc_len", "mth_off", "mth_len", "pad", "err_kind")])
JOB_DT = np.dtype([("src_off", "<u4"), ("bulk_len", "<u4"), ("head_len", "<u2"), ("pad", "u1"), ("fast", "u1"), ("slot_len", "<u4")])
OUT_DT = np.dtype([("d", MSG_DT), ("a", AUX_DT), ("job", JOB_DT), ("slot", "<u4"), ("ref", "<u4", 4), ("head", "u1", 96),
                   ("fast_ok", "<u4"), ("fast_prefix", "<u4"), ("fast_rs", "<u4"), ("fast_d", MSG_DT)])
DESC_FIELDS = ["frame_off", "body_size", "meta_size", "correlation_id", "log_id", "attachment_size", "compress_type", "checksum_type", "has_bits", "protocol",
               "content_type", "method_idx"]


@pytest.fixture(scope="module")
def kh():
    cpp = os.path.join(HERE, "cpp")
    so = os.path.join(cpp, "libdecode_host.so")
    deps = [os.path.join(cpp, f) for f in ("gen_kernels_host.py", "kernels_host_prelude.h", "h2_host_prelude.h", "decode_host.cc")] + \
           [os.path.join(ROOT, "brpc_b200", "csrc", f) for f in ("b2_kernels.cuh", "b2_core.cuh", "b2_inflate.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call([sys.executable, os.path.join(cpp, "gen_kernels_host.py")])
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-w", "-I", os.path.join(cpp, "stub"), "-I", os.path.join(ROOT, "include"),
                               "-o", so, os.path.join(cpp, "decode_host.cc")])
    lib = C.CDLL(so)
    lib.kh_create.restype = C.c_void_p
    lib.kh_create.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p]
    lib.kh_destroy.argtypes = [C.c_void_p]
    lib.kh_add_method.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.kh_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.kh_fast_reply.restype = C.c_uint32
    lib.kh_fast_reply.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    lib.kh_error_reply.restype = C.c_uint32
    lib.kh_error_reply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.kh_walks.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    assert lib.kh_sizeof_out() == OUT_DT.itemsize, (lib.kh_sizeof_out(), OUT_DT.itemsize)
    return lib


def make(kh, methods, identity=None, by_ref=0, mask=(1 << 1) | (1 << 2)):
    k = kh.kh_create(0, mask, by_ref, 0, identity)
    for m in methods:
        kh.kh_add_method(k, m["service_full_name"], m["service_name"], m["method_name"], m["request_type_name"], m["handler"], m["echo_attachment"],
                         m["response_checksum_type"], m["response_compress_type"])
    return k


def compare(kh, k, cfg, chunks, flags=0):
    data, runs = make_runs(chunks)
    runs["flags"] = flags
    rs, msgs, resp = O.process_batch(cfg, data, runs)
    buf = np.concatenate([np.asarray(data, np.uint8), np.zeros(1024, np.uint8)])
    out = np.zeros(1, OUT_DT); reply = np.zeros(1 << 20, np.uint8)
    stats = {"n": 0, "fast": 0, "fused": 0, "err": 0}
    for i in range(len(msgs)):
        m = msgs[i]
        run = runs[int(m["run_idx"]):int(m["run_idx"]) + 1].copy()
        fo = int(m["frame_off"])
        status = int(m["status"])
        if status in (9,):                       # framed-only protocols: decode_one is not what handles their payload
            continue
        fo_raw = fo | ((0 if int(m["protocol"]) == 1 or flags & 4 else 1) << 31)
        kh.kh_decode(k, buf.ctypes.data, fo_raw, run.ctypes.data, out.ctypes.data)
        d = out["d"][0]; a = out["a"][0]; job = out["job"][0]
        tag = (i, status, int(d["status"]))
        for f in DESC_FIELDS:
            if f == "method_idx" and status in (3, 4, 5):
                continue
            assert d[f] == m[f], (tag, f, d[f], m[f])
        dstatus = int(d["status"])
        if dstatus == 0 and (int(d["checksum_type"]) == 1 or int(d["compress_type"]) != 0):
            assert status in (0, 1), tag                                     # the pack stage verifies / decompresses: it may still fail there
            if status == 1:
                assert int(m["error_code"]) == 1003, tag
        elif dstatus == 7 and status in (7, 8):
            pass                                                             # client side: the pack stage verifies / decompresses as well
        else:
            assert dstatus == status and int(d["error_code"]) == int(m["error_code"]), (tag, int(d["error_code"]), int(m["error_code"]))
        want = bytes(resp[int(m["resp_off"]):int(m["resp_off"]) + int(m["resp_len"])])
        if dstatus == 1:                                                     # error reply: its exact length is known at decode time,
            assert int(d["resp_len"]) == len(want), tag                       # and lane 0 of the pack stage writes it: pack_error_reply
            n = kh.kh_error_reply(k, buf.ctypes.data, out.ctypes.data, reply.ctypes.data)
            assert n == len(want) and bytes(reply[:n]) == want, tag
            stats["err"] += 1
        if int(job["fast"]) == 1 and status == 0:                            # the bandwidth path: the pre-built prefix in front of the payload
            prefix = len(want) - int(a["msg_len"]) - int(a["att_len"])
            head = bytes(out["head"][0][int(a["pad"]):int(a["pad"]) + prefix])
            assert int(d["resp_len"]) == len(want) and head == want[:prefix], tag
            assert want[prefix:] == bytes(buf[fo + int(a["msg_off"]):fo + int(a["msg_off"]) + int(a["msg_len"]) + int(a["att_len"])]), tag
            stats["fast"] += 1
        if out["fast_ok"][0]:                                                # k_fused's exact-shape decoder took it: whole reply, in place
            assert status == 0 and int(job["fast"]) == 1, tag
            n = kh.kh_fast_reply(k, buf.ctypes.data, fo, reply.ctypes.data, reply.nbytes)
            assert n == len(want) and bytes(reply[:n]) == want, tag
            fd = out["fast_d"][0]
            for f in DESC_FIELDS:
                assert fd[f] == m[f], (tag, "fused", f)
            assert int(fd["status"]) == 0 and int(fd["resp_len"]) == len(want)
            stats["fused"] += 1
        stats["n"] += 1
    return stats


def test_decode_one_and_fused_fast_echo_on_mixed_traffic(kh):
    rng = random.Random(SEED + 951)
    total = {"n": 0, "fast": 0, "fused": 0, "err": 0}
    for identity in (None, b"10.1.2.3:8000"):
        for rck in (0, 1):
            ms = [dict(ECHO_METHOD, response_checksum_type=rck, echo_attachment=rng.choice([0, 1]))]
            k = make(kh, ms, identity=identity)
            cfg = O.make_config(methods=ms, server_identity=identity)
            for trial in range(3):
                streams = [mixed_frames(rng, rng.randrange(20, 120), big=trial == 0) for _ in range(12)]
                st = compare(kh, k, cfg, split_runs(rng, streams))
                for key in total: total[key] += st[key]
            kh.kh_destroy(k)
    assert total["n"] > 4000 and total["fast"] > 800 and total["err"] > 150, total
    # plain pipelined echoes: the shape k_fused answers itself
    k = make(kh, [dict(ECHO_METHOD)]); cfg = O.make_config()
    streams = [[echo_frame(rng, s * 1000 + i, rnd62(rng, rng.choice([0, 1, 16, 127, 128, 1024, 20000]))) for i in range(60)] for s in range(10)]
    st = compare(kh, k, cfg, split_runs(rng, streams))
    assert st["fused"] > 300 and st["fused"] <= st["fast"], st        # (it declines what does not fit in place: tiny payloads)
    kh.kh_destroy(k)


def test_decode_one_on_client_sockets_and_dump_files(kh):
    rng = random.Random(SEED + 952)
    k = make(kh, [dict(ECHO_METHOD)]); cfg = O.make_config()
    # server replies (the oracle's own, which python-protobuf pins) fed back as client-side input
    streams = [mixed_frames(rng, 80) for _ in range(8)]
    data, runs = make_runs(split_runs(rng, streams, cut_tail=False))
    rs, msgs, resp = O.process_batch(cfg, data, runs)
    replies = [bytes(resp[int(m["resp_off"]):int(m["resp_off"]) + int(m["resp_len"])]) for m in msgs if int(m["resp_len"]) and int(m["status"]) in (0, 1)]
    chunks = [b"".join(replies[i::6]) for i in range(6)]
    st = compare(kh, k, cfg, chunks, flags=1)
    assert st["n"] > 300
    kh.kh_destroy(k)


def test_speculative_walk_equals_the_plain_walk_from_any_entry(kh):
    """walk_tile_spec (k_tile_walk: steps decided from registers, next header prefetched) against walk_tile<true> (the CutInputMessage
    restatement step by step) from true frame starts, from positions inside payloads, on five-protocol traffic with garbage, for several tile ends."""
    import struct
    rng = random.Random(SEED + 953)
    from test_core_cut_host import five_protocol_stream, ALL
    out = (C.c_uint32 * 8)(); oa = (C.c_uint32 * 256)(); ob = (C.c_uint32 * 256)()
    checked = 0
    for trial in range(60):
        frames = five_protocol_stream(rng, rng.randrange(5, 60)) if trial % 2 else mixed_frames(rng, rng.randrange(5, 60))
        run = b"".join(frames)
        run = run[:rng.randrange(len(run) // 2, len(run) + 1)]
        padded = run + bytes(64)
        starts = [0]
        for f in frames: starts.append(starts[-1] + len(f))
        entries = [s_ for s_ in starts if s_ < len(run)] + [rng.randrange(len(run)) for _ in range(20)]
        for mask in (ALL, (1 << 1) | (1 << 2)):
            for entry in entries:
                tile_end = min(len(run) + 100, entry + rng.choice([1, 100, 2000, 8192, 1 << 20]))
                kh.kh_walks(padded, len(run), entry, tile_end, 64 << 20, 0, mask, out, oa, ob, 256)
                assert tuple(out[0:4]) == tuple(out[4:8]), (trial, entry, tile_end, mask, tuple(out))
                n = min(out[1], 256)
                assert list(oa[:n]) == list(ob[:n]), (trial, entry)
                checked += 1
    assert checked > 2000


def test_the_exactness_gate_with_adversarial_speculation(kh):
    """"k_tile_search / k_tile_walk only propose; k_resolve verifies the chain from the run start, so the result is by construction
    the sequential one."  Here the proposals are hostile — true frame starts, random positions (inside payloads, inside headers), frame
    look-alikes carried in payloads, none at all — on five-protocol traffic with garbage and truncation, for several tile sizes: k_tile_walk ->
    k_resolve -> k_frame_table (the kernels' own code, host build) must give the oracle's runs and frame table every time."""
    import struct
    from test_core_cut_host import five_protocol_stream, ALL
    kh.kh_front.restype = C.c_int
    kh.kh_front.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                            C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    from brpc_b200.abi import RUN_STATUS_DT
    rng = random.Random(SEED + 954)
    fake = echo_frame(rng, 5, b"x" * 30)
    n_cases = 0; n_rew = 0
    for trial in range(40):
        kind = trial % 4
        if kind == 0: streams = [b"".join(mixed_frames(rng, rng.randrange(5, 80), big=rng.random() < 0.3)) for _ in range(10)]
        elif kind == 1: streams = [b"".join(five_protocol_stream(rng, rng.randrange(5, 80))) for _ in range(10)]
        elif kind == 2: streams = [b"".join(echo_frame(rng, i, (fake * 40)[:rng.choice([100, 1024, 5000])]) for i in range(rng.randrange(5, 60))) for _ in range(10)]
        else: streams = [b"".join(echo_frame(rng, i, rnd62(rng, rng.choice([0, 10, 1024, 30000]))) for i in range(rng.randrange(1, 40))) for _ in range(10)]
        chunks = [s[:rng.randrange(len(s) + 1)] if rng.random() < 0.5 else s for s in streams]
        mask = ALL if kind == 1 else (1 << 1) | (1 << 2)
        data, runs = make_runs(chunks)
        runs["preferred_proto"] = rng.choice([-1, 1, 2])
        cfg = O.make_config(protocols=mask)
        rs, msgs, resp = O.process_batch(cfg, data, runs)
        buf = np.concatenate([np.asarray(data, np.uint8), np.zeros(1024, np.uint8)])
        for shift in (9, 11, 13):
            tile = 1 << shift
            nt = int(sum((int(l) + tile - 1) >> shift for l in runs["length"]))
            true_starts = set((int(m["run_idx"]), int(m["frame_off"]) - int(runs["offset"][int(m["run_idx"])])) for m in msgs)
            for mode in range(4):
                entries = np.full(max(nt, 1), 0xffffffff, np.uint32)
                t = 0
                for r in range(len(runs)):
                    ln = int(runs["length"][r])
                    for k in range((ln + tile - 1) >> shift):
                        lo, hi = k * tile, min((k + 1) * tile, ln)
                        firsts = sorted(p for (rr, p) in true_starts if rr == r and lo <= p < hi)
                        if k == 0:
                            e = 0
                        elif mode == 0: e = firsts[0] if firsts else 0xffffffff                 # what a perfect search proposes
                        elif mode == 1: e = rng.randrange(lo, hi)                                  # anywhere
                        elif mode == 2: e = 0xffffffff if rng.random() < 0.5 else (firsts[-1] if firsts else rng.randrange(lo, hi))   # nothing, or a LATER true start
                        else: e = rng.choice([lo, hi - 1, (firsts[0] + 1) if firsts and firsts[0] + 1 < hi else lo])
                        entries[t] = e; t += 1
                rs_d = np.zeros(len(runs), RUN_STATUS_DT); fo = np.zeros(len(msgs) + 64, np.uint32); fr = np.zeros(len(msgs) + 64, np.uint32)
                nm = C.c_uint32(); rew = C.c_uint32()
                rc = kh.kh_front(buf.ctypes.data, runs.ctypes.data, len(runs), shift, mask, 0, entries.ctypes.data, nt, rs_d.ctypes.data, fo.ctypes.data, fr.ctypes.data,
                                 len(fo), C.byref(nm), C.byref(rew))
                assert rc == 0 and nm.value == len(msgs), (trial, shift, mode, rc, nm.value, len(msgs))
                for f in ("consumed", "parse_error", "n_msgs", "first_msg", "preferred_proto"):
                    assert np.array_equal(rs_d[f], rs[f]), (trial, shift, mode, f)
                assert np.array_equal(fo[:nm.value] & 0x7fffffff, msgs["frame_off"]) and np.array_equal(fr[:nm.value], msgs["run_idx"]), (trial, shift, mode)
                assert np.array_equal(fo[:nm.value] >> 31, (msgs["protocol"] != 1).astype(np.uint32)), (trial, shift, mode)
                n_cases += 1; n_rew += rew.value
    assert n_cases == 40 * 3 * 4 and n_rew > 1000
