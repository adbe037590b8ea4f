"""GPU: the device h2 server path against a REAL grpcio client.
(1) replay: the client bytes recorded in tests/golden/h2_grpcio_capture.json (gRPC C-core talking to a brpc-style server)
    go through the oracle and through the device in the same recv() chunks: consumed bytes, parse status, number of
    requests and every byte written back (control frames + reply frames) must be identical.
(2) live: a grpcio channel talks to a TCP loop whose engine is the device (b2_h2_process_batch + b2_h2_pack_responses):
    1000 unary calls of 0 B - 64 KB, more than 8 (100+) of them in flight on one connection, all come back OK, echoed."""
import json
import os

import pytest

import _oracle as O
from _h2loop import DeviceEngine, H2LoopServer, OracleEngine

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def make_ctx(pending=192):
    import brpc_b200
    ctx = brpc_b200.Context(device=0, max_batch_bytes=16 << 20, max_msgs=1 << 14, max_runs=64, max_resp_bytes=64 << 20)
    ctx.h2_configure(max_conns=16, max_pending=pending, stream_bytes=4096 + (256 << 10))
    return ctx


@pytest.mark.parametrize("kind", ["sequential", "concurrent"])
def test_replayed_grpcio_client_bytes_device_equals_oracle(kind):
    with open(os.path.join(HERE, "golden", "h2_grpcio_capture.json")) as f:
        cap = json.load(f)[kind]
    chunks = [bytes.fromhex(c) for c in cap["chunks"]]
    eo, ed = OracleEngine(O), DeviceEngine(make_ctx())
    eo.open(0); ed.open(0)
    po = pd = b""; n_o = n_d = 0
    for i, ch in enumerate(chunks):
        po += ch; pd += ch
        co, oo, erro, no = eo.feed(0, po)
        cd, od, errd, nd = ed.feed(0, pd)
        assert (cd, errd, nd) == (co, erro, no), "chunk %d: device (consumed, err, n) %r oracle %r" % (i, (cd, errd, nd), (co, erro, no))
        assert od == oo, "chunk %d: bytes written back differ (device %d B, oracle %d B)" % (i, len(od), len(oo))
        po = po[co:]; pd = pd[cd:]; n_o += no; n_d += nd
    assert n_d == cap["n_requests"]


def test_live_grpcio_client_against_the_device():
    grpc = pytest.importorskip("grpc")
    from test_oracle_h2_grpcio import SIZES, _channel, _echo, pb
    srv = H2LoopServer(DeviceEngine(make_ctx()))
    try:
        with _channel(srv.port) as ch:
            call = _echo(ch)
            for i in range(200):                                 # one at a time
                msg = pb(SIZES[i % len(SIZES)] % 66000, bytes([97 + i % 26]))
                assert call(msg, timeout=30) == msg
            import threading
            gate = threading.BoundedSemaphore(128)               # 128 calls in flight on the one connection (the pool holds 192 streams)
            futs = []
            for i in range(800):
                msg = pb(SIZES[(7 * i) % len(SIZES)] % 66000, bytes([65 + i % 26]))
                gate.acquire()
                f = call.future(msg, timeout=120, metadata=(("x-trace", "t%d" % (i % 9)),))
                f.add_done_callback(lambda _f: gate.release())
                futs.append((msg, f))
            for msg, f in futs:
                try:
                    got = f.result()
                except grpc.RpcError as e:
                    raise AssertionError("call failed: %s; server loop errors: %r" % (e, srv.errors))
                assert got == msg
            with pytest.raises(grpc.RpcError) as e:
                ch.unary_unary("/example.EchoService/Nope", request_serializer=lambda b: b, response_deserializer=lambda b: b)(b"x", timeout=30)
            assert e.value.code() == grpc.StatusCode.UNIMPLEMENTED
        assert srv.n_requests == 1001 and not srv.errors, srv.errors
    finally:
        srv.close()
