"""CPU: pins the h2 framing oracle (oracle/b2_oracle_h2.c) with an INDEPENDENT h2/gRPC implementation: a real grpcio
client (gRPC C-core) talks over TCP to a loop whose engine is the oracle — ParseH2Message's side effects (SETTINGS,
WINDOW_UPDATE, acks, pongs) and H2UnsentResponse/PackH2Message's reply framing (HEADERS, DATA split at the peer's
max_frame_size, gRPC trailers, HPACK with the encoder table) must all be acceptable to it, and every call must come back
OK with the echoed message."""
import threading

import pytest

import _oracle as O
from _h2loop import H2LoopServer, OracleEngine

grpc = pytest.importorskip("grpc")

SIZES = [0, 1, 5, 100, 1000, 4096, 16384 - 5, 16384, 16385, 40000, 65536, 200000]


def _channel(port, **opt):
    return grpc.insecure_channel("127.0.0.1:%d" % port, options=[("grpc.max_receive_message_length", 1 << 24),
                                                                 ("grpc.max_send_message_length", 1 << 24)] + list(opt.items()))


def _echo(ch):
    return ch.unary_unary("/example.EchoService/Echo", request_serializer=lambda b: b, response_deserializer=lambda b: b)


def pb(n, fill=b"g"):
    body = fill * n                                              # EchoRequest{message}: 0a <len> <bytes>
    v, ln = n, b""
    while v >= 0x80:
        ln += bytes([v & 0x7f | 0x80]); v >>= 7
    return b"\x0a" + ln + bytes([v]) + body


def test_grpcio_unary_calls_against_the_oracle():
    srv = H2LoopServer(OracleEngine(O))
    try:
        with _channel(srv.port) as ch:
            call = _echo(ch)
            for i in range(300):
                msg = pb(SIZES[i % len(SIZES)], bytes([97 + i % 26]))
                assert call(msg, timeout=20) == msg
            # an unknown method: the loop answers grpc-status 12 with a grpc-message trailer
            with pytest.raises(grpc.RpcError) as e:
                ch.unary_unary("/example.EchoService/Nope", request_serializer=lambda b: b, response_deserializer=lambda b: b)(b"x", timeout=20)
            assert e.value.code() == grpc.StatusCode.UNIMPLEMENTED and e.value.details() == "unimplemented"
        assert srv.n_requests == 301 and not srv.errors, srv.errors
    finally:
        srv.close()


def test_grpcio_concurrent_streams_against_the_oracle():
    """100+ calls in flight on ONE connection (grpc's default concurrency), mixed sizes, metadata that exercises the HPACK
    decoder's dynamic table (literal with incremental indexing, then indexed)."""
    srv = H2LoopServer(OracleEngine(O))
    try:
        with _channel(srv.port) as ch:
            call = _echo(ch)
            futs = []
            for i in range(700):
                msg = pb(SIZES[(7 * i) % len(SIZES)] % 70000, bytes([65 + i % 26]))
                md = (("x-trace", "t%d" % (i % 9)), ("x-fixed", "same-value-every-time"))
                futs.append((msg, call.future(msg, timeout=60, metadata=md)))
            for msg, f in futs:
                assert f.result() == msg
        assert srv.n_requests == 700 and not srv.errors, srv.errors
        assert len(srv.capture) == 1                             # one TCP connection carried them all
    finally:
        srv.close()


def test_grpcio_many_connections_against_the_oracle():
    srv = H2LoopServer(OracleEngine(O))
    errs = []

    def client(k):
        try:
            with _channel(srv.port, **{"grpc.use_local_subchannel_pool": 1}) as ch:
                call = _echo(ch)
                for i in range(40):
                    msg = pb((k * 131 + i * 977) % 30000, bytes([48 + k]))
                    assert call(msg, timeout=30) == msg
        except Exception as e:                                   # noqa: BLE001
            errs.append(repr(e))
    try:
        ts = [threading.Thread(target=client, args=(k,)) for k in range(8)]
        [t.start() for t in ts]; [t.join() for t in ts]
        assert not errs and not srv.errors, (errs, srv.errors)
        assert srv.n_requests == 320 and len(srv.capture) == 8
    finally:
        srv.close()
