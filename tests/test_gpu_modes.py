"""GPU: the input / response modes of the C ABI (b2_set_modes) give the SAME results as the default copy path:
B2_INPUT_PULL (kernels read the pinned batch buffer in place over PCIe) and B2_RESP_BY_REF (an OK echo reply is
{prefix, reference into the request bytes}; gathering prefix + reference must reproduce the oracle's reply byte for byte,
SendRpcResponse's append-by-reference, baidu_rpc_protocol.cpp:383-389)."""
import ctypes as C
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import _oracle as O  # noqa: E402
from _compare import MSG_FIELDS, RUN_FIELDS  # noqa: E402
from _traffic import SEED, echo_frame, mixed_frames, rnd62, split_runs  # noqa: E402


def replies(data, msgs, resp, refs):
    out = []
    for i in range(len(msgs)):
        o, n = int(msgs["resp_off"][i]), int(msgs["resp_len"][i])
        if msgs["status"][i] == 7:                       # client-side result reported in place
            out.append(b""); continue
        if refs is not None and refs["src_len"][i]:
            p, so, sl = int(refs["prefix_len"][i]), int(refs["src_off"][i]), int(refs["src_len"][i])
            assert p + sl == n and p <= 64
            out.append(bytes(resp[o:o + p]) + bytes(data[so:so + sl]))
        else:
            if refs is not None:
                assert int(refs["prefix_len"][i]) in (0, n) or n == 0
            out.append(bytes(resp[o:o + n]))
    return out


def check(b2, ctx, chunks, cfg, modes, what):
    from brpc_b200.abi import PinnedBuffer
    data, runs = b2.make_runs(chunks)
    pin = PinnedBuffer(max(len(data), 16)); pin.array[:len(data)] = data
    o_rs, o_msgs, o_resp = O.process_batch(cfg, data, runs)
    want = replies(data, o_msgs, o_resp, None)
    for im, rm in modes:
        ctx.set_modes(im, rm)
        rs, msgs, resp, info = ctx.process_batch_ptr(pin.ptr, len(data), runs)
        tag = "%s input=%d resp=%d" % (what, im, rm)
        for f in RUN_FIELDS:
            assert np.array_equal(rs[f], o_rs[f]), tag + " run." + f
        assert len(msgs) == len(o_msgs), tag
        for f in MSG_FIELDS:
            assert np.array_equal(msgs[f], o_msgs[f]), tag + " msgs." + f
        refs = info["refs"]
        assert (refs is not None) == (rm == 1) or len(msgs) == 0
        if rm == 2 and len(msgs):
            # B2_RESP_IOVEC: the device-written gather list, host addresses into the pinned reply block and the request bytes
            iov = info["iov"]
            assert iov is not None and len(iov) == 2 * len(msgs)
            answered = (msgs["status"] == 0) | (msgs["status"] == 1)
            for k in range(len(msgs)):
                g = b"".join(C.string_at(int(iov["base"][j]), int(iov["len"][j])) for j in (2 * k, 2 * k + 1) if iov["len"][j])
                assert g == (want[k] if answered[k] else b""), "%s iovec reply %d differs (status %d)" % (tag, k, msgs["status"][k])
            lo, hi = pin.ptr, pin.ptr + len(data)
            second = iov[1::2]
            assert np.all((second["len"] == 0) | ((second["base"] >= lo) & (second["base"] + second["len"] <= hi)))
            for r in range(len(rs)):
                f0, n = int(rs["first_msg"][r]), int(rs["n_msgs"][r])
                assert int(rs["n_unanswered"][r]) == int((~answered[f0:f0 + n]).sum()), tag
            continue
        got = replies(pin.array, msgs, resp, refs)
        for k, (g, w) in enumerate(zip(got, want)):
            assert g == w, "%s reply %d differs (status %d)" % (tag, k, msgs["status"][k])
        if rm == 1 and len(msgs):
            ok_echo = (msgs["status"] == 0) & (msgs["compress_type"] == 0)
            # plain echoes really are by reference (nothing but the prefix came back for them)
            assert np.all(refs["src_len"][ok_echo & (msgs["resp_len"] > 64)] > 0) or cfg.methods[0].response_checksum_type or cfg.methods[0].response_compress_type
    ctx.set_modes(0, 0)


ALL = [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2)]


def test_modes_on_mixed_traffic_small_and_large_batches(monkeypatch):
    import brpc_b200 as b2
    rng = random.Random(SEED + 77)
    ctx = b2.Context(device=0, max_batch_bytes=64 << 20, max_msgs=1 << 18, max_runs=4096)
    cfg = O.make_config()
    # latency path (k_small) ...
    streams = [mixed_frames(rng, rng.randrange(1, 12)) for _ in range(24)]
    check(b2, ctx, split_runs(rng, streams), cfg, ALL, "small")
    # ... and the tile pipeline
    streams = [mixed_frames(rng, rng.randrange(50, 400), big=True) for _ in range(40)]
    check(b2, ctx, split_runs(rng, streams), cfg, ALL, "large")
    monkeypatch.setenv("B2_SMALL", "off")
    ctx2 = b2.Context(device=0, max_batch_bytes=64 << 20, max_msgs=1 << 18, max_runs=4096, tile_bytes=1024)
    streams = [mixed_frames(rng, rng.randrange(1, 30)) for _ in range(64)]
    check(b2, ctx2, split_runs(rng, streams), cfg, ALL, "tiny tiles")


def test_modes_with_response_checksum_and_client_side():
    import brpc_b200 as b2
    rng = random.Random(SEED + 78)
    m = dict(b2.ECHO_METHOD); m["response_checksum_type"] = 1
    ctx = b2.Context(device=0, max_batch_bytes=32 << 20, max_msgs=1 << 16, max_runs=1024, methods=(m,), server_identity=b"10.1.2.3:8000")
    om = dict(O.ECHO_METHOD); om["response_checksum_type"] = 1
    cfg = O.make_config([om], server_identity=b"10.1.2.3:8000")
    streams = [mixed_frames(rng, rng.randrange(20, 200)) for _ in range(30)]
    check(b2, ctx, split_runs(rng, streams), cfg, ALL, "resp crc")
    m2 = dict(b2.ECHO_METHOD); m2["echo_attachment"] = 0
    ctx3 = b2.Context(device=0, max_batch_bytes=32 << 20, max_msgs=1 << 16, max_runs=1024, methods=(m2,))
    om2 = dict(O.ECHO_METHOD); om2["echo_attachment"] = 0
    check(b2, ctx3, split_runs(rng, streams), O.make_config([om2]), ALL, "no echo attachment")


def test_bench_shaped_batch_by_ref_pull():
    """64 connections x 1 MiB of pipelined 1 KB requests: every reply is a reference, results equal the oracle's."""
    import brpc_b200 as b2
    from brpc_b200 import press
    from brpc_b200.abi import PinnedBuffer
    sp = press.spec(payload_bytes=1024, payload_kind=1)
    run_bytes = (1 << 20) - 112
    pin = PinnedBuffer(64 * (1 << 20))
    runs, n_full = press.fill_batch(sp, pin.array, 64, run_bytes)
    ctx = b2.Context(device=0, max_batch_bytes=(64 << 20) + (1 << 20), max_msgs=n_full + 4096, max_runs=64)
    o_rs, o_msgs, o_resp = O.process_batch(O.make_config(), pin.array, runs)
    want = replies(pin.array, o_msgs, o_resp, None)
    for im, rm in ALL:
        ctx.set_modes(im, rm)
        for rep in range(2):                              # (second pass: tile size adapted to the message size)
            rs, msgs, resp, info = ctx.process_batch_ptr(pin.ptr, 64 << 20, runs)
        assert len(msgs) == n_full == len(o_msgs)
        for f in MSG_FIELDS:
            assert np.array_equal(msgs[f], o_msgs[f]), f
        if rm == 2:
            iov = info["iov"]
            assert np.all(iov["len"][1::2] == 1024) and np.all(iov["len"][0::2] <= 48) and np.all(rs["n_unanswered"] == 0)
            got = [C.string_at(int(iov["base"][2 * k]), int(iov["len"][2 * k])) + C.string_at(int(iov["base"][2 * k + 1]), 1024) for k in range(0, n_full, 97)]
            assert got == want[::97]
            continue
        got = replies(pin.array, msgs, resp, info["refs"])
        assert got == want
        if rm == 1:
            assert np.all(info["refs"]["src_len"] == 1024) and np.all(info["refs"]["prefix_len"] <= 48)
            assert int(rs["resp_bytes"].sum()) <= 64 * n_full
