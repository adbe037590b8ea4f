"""GPU: the device path against the python-protobuf / reference-leaf golden frames of
tests/golden/encode_vectors.json DIRECTLY (no oracle in between): every request frame is served through the C ABI
under the vector's server configuration and the reply bytes must equal the golden response frame; the device client
mirror (b2_pack_requests) must reproduce the golden request frames."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    with open(os.path.join(HERE, "golden", "encode_vectors.json")) as f:
        return json.load(f)


def test_device_replies_equal_golden_response_frames():
    import brpc_b200
    vec = load()["rpc"]
    groups = {}
    for v in vec:
        s = v["server"]
        groups.setdefault((s["echo_attachment"], s["response_checksum"], s["response_compress"], s["identity"]), []).append(v)
    n = 0
    for (ea, rc, rz, ident), vs in groups.items():
        m = dict(brpc_b200.ECHO_METHOD); m["echo_attachment"] = ea; m["response_checksum_type"] = rc; m["response_compress_type"] = rz
        ctx = brpc_b200.Context(device=0, max_batch_bytes=16 << 20, max_msgs=1 << 14, max_runs=1024, methods=(m,),
                                server_identity=ident.encode() if ident else None)
        # one connection per vector AND all of them pipelined on one connection
        wires = [bytes.fromhex(v.get("wire_request_hex", v["request_hex"])) for v in vs]
        for streams in (wires, [b"".join(wires)]):
            data, runs = brpc_b200.make_runs(streams)
            rs, msgs, resp, _ = ctx.process_batch(data, runs)
            assert len(msgs) == len(vs)
            for k, v in enumerate(vs):
                got = bytes(resp[int(msgs["resp_off"][k]):int(msgs["resp_off"][k]) + int(msgs["resp_len"][k])])
                assert int(msgs["error_code"][k]) == v["error_code"]
                assert got.hex() == v["response_hex"], "golden response %d (error %d) differs" % (k, v["error_code"])
                n += 1
        ctx.close()
    assert n >= 800


def test_device_pack_requests_equal_golden_request_frames():
    import brpc_b200
    from brpc_b200.abi import REQUEST_DT
    ctx = brpc_b200.Context(device=0, max_batch_bytes=16 << 20, max_msgs=1 << 14, max_runs=64)
    blob = bytearray(); reqs = []; expect = []
    for v in load()["rpc"]:
        if v["trace"] or v["request_id"] or v["service"] != "example.EchoService" or v["method"] != "Echo":
            continue                                       # b2_pack_requests packs registered methods, no tracing fields (b2rpc.h)
        msg = bytes.fromhex(v["message_hex"]); att = bytes.fromhex(v["attachment_hex"])
        p = len(blob); blob += msg; a = len(blob); blob += att
        has_log = v["log_id"] is not None
        reqs.append((0, (1 if has_log else 0) | (2 if v["timeout_ms"] > 0 else 0), 0, v["timeout_ms"], v["correlation_id"],
                     v["log_id"] if has_log else 0, v["compress"], v["checksum"], 0, p, len(msg), a, len(att), 0))
        expect.append(v["request_hex"])
    for v in load()["stream"]:
        d = bytes.fromhex(v["data_hex"]); p = len(blob); blob += d
        hc = v["has_continuation"]
        fl = (1 if v["source_stream_id"] is not None else 0) | (0 if hc is None else 2 | (4 if hc else 0))
        reqs.append((1, fl, -1, 0, v["stream_id"], v["source_stream_id"] or 0, 0, 0, v["frame_type"], p, len(d), 0, 0, 0))
        expect.append(v["frame_hex"])
    got = ctx.pack_requests(np.frombuffer(bytes(blob) + bytes(16), np.uint8), np.array(reqs, dtype=REQUEST_DT))
    assert len(got) > 250
    for i, (g, x) in enumerate(zip(got, expect)):
        assert g.hex() == x, "golden request %d differs" % i
