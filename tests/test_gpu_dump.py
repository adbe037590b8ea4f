"""GPU: rpc_dump files as a byte-stream source (B2_RUN_RPC_DUMP): records cut like SampleIterator::Pop, baidu_std samples re-packed
as the request frames rpc_replay sends — against the python-protobuf golden frames, against the oracle on larger files, and
end to end: the replayed frames go through the server path and come back echoed."""
import json
import os
import random
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import _oracle as O  # noqa: E402
from _compare import assert_same  # noqa: E402
from test_oracle_dump import check, load  # noqa: E402


def test_device_replays_golden_dump_files(monkeypatch):
    import brpc_b200 as b2
    ctx = b2.Context(device=0, max_batch_bytes=16 << 20, max_msgs=1 << 14, max_runs=64)
    files = load()
    for f in files:                                                # one file per batch (latency path) ...
        data = np.frombuffer(bytes.fromhex(f["file_hex"]) + bytes(64), np.uint8)
        runs = np.zeros(1, b2.RUN_DT); runs[0] = (f["base_cid"], 0, len(data) - 64, -1, 4)
        rs, msgs, resp, _ = ctx.process_batch(data, runs)
        check(rs, msgs, resp, f)
    monkeypatch.setenv("B2_SMALL", "off")                          # ... and all of them in one batch through the tile pipeline, beside a live socket
    ctx2 = b2.Context(device=0, max_batch_bytes=16 << 20, max_msgs=1 << 14, max_runs=64, tile_bytes=2048)
    blobs = [bytes.fromhex(f["file_hex"]) for f in files] + [O.pack_echo_request(message=b"live" * 50, correlation_id=77)]
    data, runs = b2.make_runs(blobs)
    for k, f in enumerate(files):
        runs["socket_id"][k] = f["base_cid"]; runs["flags"][k] = 4
    dev = ctx2.process_batch(data, runs)
    assert_same(dev, O.process_batch(O.make_config(), data, runs), "dump files + a socket")
    assert dev[1]["status"][-1] == 0


def test_replayed_requests_are_served():
    """A dump of 3000 echo requests -> replayed frames -> the server path echoes every one of them (rpc_replay against a server)."""
    import brpc_b200 as b2
    rng = random.Random(20260921)
    dump = b""
    msgs_in = []
    for i in range(3000):
        text = bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.choice([0, 7, 100, 1024])))
        v, ln = len(text), b""
        while v >= 0x80: ln += bytes([v & 0x7f | 0x80]); v >>= 7
        body = b"\x0a" + ln + bytes([v]) + text
        meta = b"\x0a\x13example.EchoService\x12\x04Echo\x28\x01"           # service_name, method_name, protocol_type = baidu_std
        dump += b"PRPC" + struct.pack(">II", len(meta) + len(body), len(meta)) + meta + body
        msgs_in.append(text)
    ctx = b2.Context(device=0, max_batch_bytes=16 << 20, max_msgs=1 << 14, max_runs=8)
    data, runs = b2.make_runs([dump]); runs["flags"] = 4; runs["socket_id"] = 1000
    rs, msgs, resp, _ = ctx.process_batch(data, runs)
    assert len(msgs) == 3000 and np.all(msgs["status"] == 10) and int(rs["consumed"][0]) == len(dump)
    wire = b"".join(bytes(resp[int(m["resp_off"]):int(m["resp_off"]) + int(m["resp_len"])]) for m in msgs)
    data2, runs2 = b2.make_runs([wire])
    rs2, msgs2, resp2, _ = ctx.process_batch(data2, runs2)
    assert len(msgs2) == 3000 and np.all(msgs2["status"] == 0) and np.array_equal(msgs2["correlation_id"], 1000 + np.arange(3000))
    for k in (0, 1, 17, 2999):
        m = msgs2[k]; reply = bytes(resp2[int(m["resp_off"]):int(m["resp_off"]) + int(m["resp_len"])])
        assert reply.endswith(msgs_in[k])
