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


from _h2server_scenario import run_scenario  # noqa: E402


def _run(n_conns, n_calls, violations, seed, step_choices):
    import brpc_b200
    mk = lambda: brpc_b200.Context(device=0, max_batch_bytes=32 << 20, max_msgs=1 << 16, max_runs=1024, max_resp_bytes=64 << 20)
    return run_scenario(mk, brpc_b200.make_runs, n_conns, n_calls, violations, seed, step_choices)


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
    big = T.PREFACE + b"".join(T.request_frames(rng, enc2, 1, message=b"z" * 70000, chunk=9000))
    data, runs = brpc_b200.make_runs([many, big])
    rs, msgs, out = ctx.h2_process_batch(data, runs)
    assert list(rs["parse_error"]) == [4, 4] and len(msgs) == 0


def test_grpc_echo_responses_packed_on_the_device():
    """process -> echo -> b2_h2_pack_responses: HPACK encoder state, DATA splitting, trailers, flow control and the
    deferred connection WINDOW_UPDATE against the oracle's AppendAndDestroySelf / PackH2Message, over several batches."""
    import brpc_b200
    from brpc_b200.abi import H2_RESPONSE_DT
    rng = random.Random(SEED + 7)
    n_conns = 40
    ctx = brpc_b200.Context(device=0, max_batch_bytes=32 << 20, max_msgs=1 << 16, max_runs=1024, max_resp_bytes=64 << 20)
    orc = [O.H2Conn() for _ in range(n_conns)]
    first = [T.settings(rng.choice([(), ((1, 0),), ((5, 20000),), ((4, 100000),)])) for _ in range(n_conns)]
    scripts = []
    for i in range(n_conns):
        ctx.h2_conn_reset(i)
        enc = T.HpackEncoder(rng)
        calls = []
        for k in range(14):
            msg = bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz0123456789") for _ in range(rng.choice([0, 10, 300, 4096, 9000])))
            calls.append(b"".join(T.request_frames(rng, enc, 1 + 2 * k, message=msg, content_type=rng.choice([b"application/grpc", b"application/grpc+proto"]))))
        scripts.append([T.PREFACE + first[i]] + calls)
    n_resp = 0; n_rst = 0; n_wu = 0
    for rnd in range(15):
        streams = [scripts[i][rnd] for i in range(n_conns)]
        data, runs = brpc_b200.make_runs(streams)
        rs, msgs, out = ctx.h2_process_batch(data, runs)
        blob_parts = []; resps = []; expect = []
        for i in range(n_conns):
            e, cons, omsgs, octrl, oblob, _, _ = orc[i].consume(streams[i])
            st = rs[i]
            assert int(st["n_msgs"]) == len(omsgs) and bytes(out[st["ctrl_off"]:st["ctrl_off"] + st["ctrl_len"]]) == octrl
            for m in msgs[st["first_msg"]:st["first_msg"] + st["n_msgs"]]:
                assert m["flags"] & 2
                src = data if (m["flags"] & 16) else out
                body = bytes(src[m["msg_off"]:m["msg_off"] + m["msg_len"]])          # echo: the reply message is the request message
                hb = bytes(out[m["headers_off"]:m["headers_off"] + m["headers_len"]])
                ct = dict(O.parse_header_records(hb))[b"content-type"]
                ct_off = int(m["headers_off"]) + hb.index(ct)                        # the request's own content-type value, inside out
                fail = rng.random() < 0.15
                gm = b"Fail%20to%20find%20method%20" + str(rnd * 7 + i).encode() if fail else b""
                base = sum(len(x) for x in blob_parts)
                if fail:
                    body = b""; blob_parts += [gm]
                    resps.append((i, int(m["stream_id"]), 200, 1 | 8, ct_off, len(ct), 0, 0, 12, base, len(gm), 0))
                elif rng.random() < 0.15:                                            # a reply built on the host: larger than a frame / than the window
                    body = body * 6; blob_parts += [body]
                    resps.append((i, int(m["stream_id"]), 200, 1 | 8, ct_off, len(ct), base, len(body), 0, 0, 0, 0))
                else:                                                                # zero copy: body still on the device (input or out buffer)
                    resps.append((i, int(m["stream_id"]), 200, 1 | 8 | (2 if (m["flags"] & 16) else 4), ct_off, len(ct), int(m["msg_off"]), len(body), 0, 0, 0, 0))
                expect.append(orc[i].pack_response(int(m["stream_id"]), body, 200, ct, True, 12 if fail else 0, gm))
        if not resps:
            continue
        blob = np.frombuffer(b"".join(blob_parts) + b"\0", np.uint8)
        got = ctx.h2_pack_responses(blob, np.array(resps, dtype=H2_RESPONSE_DT))
        for k, (g, x) in enumerate(zip(got, expect)):
            assert g == x, (rnd, resps[k][:2])
        n_resp += len(resps); n_rst += sum(1 for x in expect if len(x) == 13 and x[3] == 3); n_wu += sum(1 for x in expect if x[-13:-9] == b"\x00\x00\x04\x08")
    assert n_resp > 400 and n_rst > 0
