"""GPU: b2_h2_process_batch (the server side of ParseH2Message on the device) against the oracle: many connections,
their byte streams delivered in random pieces over many batches, gRPC calls with CONTINUATION / padding / trailers /
interleaved streams, control frames, and protocol violations whose exact (quirky) aftermath must match."""
import random

import numpy as np
import pytest

import _oracle as O
import _h2traffic as T

pytestmark = pytest.mark.gpu
SEED = 20260921


def _msg_tuple(m, blob):
    return (int(m["stream_id"]), int(m["n_headers"]), bytes(blob[m["headers_off"]:m["headers_off"] + m["headers_len"]]),
            bytes(blob[m["body_off"]:m["body_off"] + m["body_len"]]), int(m["http_method"]), int(m["content_type"]), int(m["flags"]),
            int(m["method_idx"]), bytes(blob[m["msg_off"]:m["msg_off"] + m["msg_len"]]), bytes(blob[m["path_off"]:m["path_off"] + m["path_len"]]))


def _run(n_conns, n_calls, violations, seed, step_choices):
    import brpc_b200
    rng = random.Random(seed)
    ctx = brpc_b200.Context(device=0, max_batch_bytes=32 << 20, max_msgs=1 << 16, max_runs=1024, max_resp_bytes=64 << 20)
    streams = [b"".join(T.connection_script(rng, n_calls=n_calls, violations=violations if i % 2 else 0.0)) for i in range(n_conns)]
    orc = [O.H2Conn() for _ in range(n_conns)]
    for i in range(n_conns):
        ctx.h2_conn_reset(i)
    fed = [0] * n_conns; buf = [b""] * n_conns; alive = [True] * n_conns
    total_msgs = 0; total_ctrl = 0; errors = {}
    while any(alive[i] and (fed[i] < len(streams[i])) for i in range(n_conns)):
        live = [i for i in range(n_conns) if alive[i] and fed[i] < len(streams[i])]
        batch = [i for i in live if rng.random() < 0.8] or live[:1]
        for i in batch:
            k = rng.choice(step_choices)
            buf[i] += streams[i][fed[i]:fed[i] + k]; fed[i] += k
        data, runs = brpc_b200.make_runs([buf[i] for i in batch])
        runs["socket_id"] = np.array(batch, dtype=np.uint64)
        rs, msgs, out = ctx.h2_process_batch(data, runs)
        for j, i in enumerate(batch):
            e, cons, omsgs, octrl, oblob, mfs, sws = orc[i].consume(buf[i])
            st = rs[j]
            assert (int(st["parse_error"]), int(st["consumed"]), int(st["n_msgs"])) == (e, cons, len(omsgs)), (i, fed[i])
            assert bytes(out[st["ctrl_off"]:st["ctrl_off"] + st["ctrl_len"]]) == octrl, (i, fed[i])
            assert (int(st["remote_max_frame_size"]), int(st["remote_stream_window_size"])) == (mfs, sws)
            dm = msgs[st["first_msg"]:st["first_msg"] + st["n_msgs"]]
            for a, b in zip(dm, omsgs):
                assert int(a["run_idx"]) == j
                assert _msg_tuple(a, out) == _msg_tuple(b, oblob), (i, fed[i], int(b["stream_id"]))
            total_msgs += len(omsgs); total_ctrl += len(octrl)
            buf[i] = buf[i][cons:]
            if e != 2:
                alive[i] = False; errors[e] = errors.get(e, 0) + 1
    return total_msgs, total_ctrl, errors


def test_grpc_connections_clean_traffic():
    msgs, ctrl, errors = _run(n_conns=48, n_calls=20, violations=0.0, seed=SEED, step_choices=[1, 9, 100, 1500, 5000, 20000])
    assert msgs > 48 * 15 and ctrl > 48 * 43 and not errors


def test_grpc_connections_with_protocol_violations():
    msgs, ctrl, errors = _run(n_conns=64, n_calls=24, violations=0.25, seed=SEED + 1, step_choices=[3, 50, 700, 4000, 30000])
    assert msgs > 300 and ctrl > 64 * 43


def test_device_limits_end_the_run_with_no_resource():
    """More concurrent streams / bigger bodies than the device keeps per connection: PARSE_ERROR_NO_RESOURCE (documented limit)."""
    import brpc_b200
    rng = random.Random(3)
    ctx = brpc_b200.Context(device=0, max_batch_bytes=8 << 20, max_msgs=4096, max_runs=64, max_resp_bytes=16 << 20)
    enc = T.HpackEncoder(rng)
    ctx.h2_conn_reset(0); ctx.h2_conn_reset(1)
    many = T.PREFACE + b"".join(T.request_frames(rng, enc, 1 + 2 * k, message=b"m")[0] for k in range(12))     # 12 HEADERS, no END_STREAM
    enc2 = T.HpackEncoder(rng)
    big = T.PREFACE + b"".join(T.request_frames(rng, enc2, 1, message=b"z" * 14000, chunk=4000))
    data, runs = brpc_b200.make_runs([many, big])
    rs, msgs, out = ctx.h2_process_batch(data, runs)
    assert list(rs["parse_error"]) == [4, 4] and len(msgs) == 0
