"""GPU: the client mirror — b2_pack_requests (PackRpcRequest / PackStreamMessage on the device) against the oracle's
encoders byte for byte, then the frames go through the server path and come back as responses the client side accepts."""
import random

import numpy as np
import pytest

import _oracle as O
from _traffic import rnd62

pytestmark = pytest.mark.gpu
SEED = 20260921


def test_pack_requests_bit_exact_and_round_trip():
    import brpc_b200
    from brpc_b200.abi import REQUEST_DT
    rng = random.Random(SEED + 31)
    ctx = brpc_b200.Context(device=0, max_batch_bytes=64 << 20, max_msgs=1 << 16, max_runs=256, max_resp_bytes=96 << 20)
    blob = bytearray(); reqs = []; expect = []
    for i in range(1500):
        kind = 1 if rng.random() < 0.2 else 0
        n = rng.choice([0, 1, 15, 16, 100, 127, 128, 1024, 4096, 70000])
        payload = rng.choice([b"r" * n, rnd62(rng, min(n, 20000)), (rnd62(rng, 50) * (n // 50 + 1))[:n]])
        att = rnd62(rng, rng.choice([0, 0, 0, 7, 300]))
        p_off = len(blob); blob += payload; a_off = len(blob); blob += att
        cid = rng.choice([0, 1, (rng.randrange(1 << 20) << 32) | rng.randrange(1, 8), (1 << 62) + 5, -7])
        if kind == 1:
            has_src = rng.random() < 0.8; cont = rng.choice([None, False, True]); ft = rng.choice([1, 2, 3, 4])
            src_id = rng.randrange(1 << 40)
            reqs.append((1, (1 if has_src else 0) | (0 if cont is None else 2 | (4 if cont else 0)), -1, 0, cid, src_id, 0, 0, ft, p_off, len(payload), 0, 0, 0))
            expect.append(O.pack_stream_frame(cid, src_id if has_src else -1, ft, cont, payload))
        else:
            has_log = rng.random() < 0.7; log_id = rng.choice([0, 5, 12345, 1 << 40, -1]); to = rng.choice([0, 0, 100, 60000])
            comp = rng.choice([0, 0, 1]); cks = rng.choice([0, 1])
            reqs.append((0, (1 if has_log else 0) | (2 if to > 0 else 0), 0, to, cid, log_id, comp, cks, 0, p_off, len(payload), a_off, len(att), 0))
            expect.append(O.pack_echo_request(log_id=log_id if has_log else None, correlation_id=cid, message=payload, attachment=att,
                                              compress_type=comp, checksum_type=cks, timeout_ms=to))
    data = np.frombuffer(bytes(blob) + b"\0" * 16, np.uint8)
    got = ctx.pack_requests(data, np.array(reqs, dtype=REQUEST_DT))
    for i, (g, x) in enumerate(zip(got, expect)):
        assert g == x, (i, reqs[i][:9], len(g), len(x))
    # the packed baidu_std requests are real traffic: the server path echoes them, the client path accepts the replies
    frames = [g for g, r in zip(got, reqs) if r[0] == 0]
    streams = [b"".join(frames[k::8]) for k in range(8)]
    data2, runs2 = brpc_b200.make_runs(streams)
    rs, msgs, resp, _ = ctx.process_batch(data2, runs2)
    assert len(msgs) == len(frames) and np.all(msgs["status"] == 0)
    replies = [b"".join(bytes(resp[m["resp_off"]:m["resp_off"] + m["resp_len"]]) for m in msgs[rs[k]["first_msg"]:rs[k]["first_msg"] + rs[k]["n_msgs"]]) for k in range(8)]
    data3, runs3 = brpc_b200.make_runs(replies)
    runs3["flags"] = 1                                                   # B2_RUN_CLIENT
    rs3, msgs3, resp3, _ = ctx.process_batch(data3, runs3)
    assert len(msgs3) == len(frames) and np.all(msgs3["status"] == 7) and np.all(msgs3["error_code"] == 0)
