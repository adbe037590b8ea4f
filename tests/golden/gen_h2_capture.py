#!/usr/bin/env python
"""Record what a REAL grpcio client (gRPC C-core h2) sends to a brpc-style h2 server: tests/golden/h2_grpcio_capture.json.
The server side of the recording session is the CPU oracle behind tests/_h2loop.py; only the CLIENT bytes (per TCP
connection, in recv() chunks) are stored.  Replaying them through the oracle and through the device must give the same
control bytes, request descriptors and reply frames (tests/test_gpu_h2_grpcio.py); the oracle side is itself pinned by
having completed these calls with the real client (tests/test_oracle_h2_grpcio.py)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import grpc  # noqa: E402
import _oracle as O  # noqa: E402
from _h2loop import H2LoopServer, OracleEngine  # noqa: E402
from test_oracle_h2_grpcio import pb, _channel, _echo, SIZES  # noqa: E402


def record(kind):
    srv = H2LoopServer(OracleEngine(O))
    with _channel(srv.port) as ch:
        call = _echo(ch)
        if kind == "sequential":
            for i in range(60):
                msg = pb(SIZES[i % len(SIZES)] % 50000, bytes([97 + i % 26]))
                assert call(msg, timeout=20, metadata=(("x-n", str(i % 4)),)) == msg
            try:
                ch.unary_unary("/example.EchoService/Nope", request_serializer=lambda b: b, response_deserializer=lambda b: b)(b"x", timeout=20)
            except grpc.RpcError:
                pass
        else:
            futs = []
            for i in range(160):
                msg = pb(SIZES[(7 * i) % len(SIZES)] % 20000, bytes([65 + i % 26]))
                futs.append((msg, call.future(msg, timeout=60, metadata=(("x-trace", "t%d" % (i % 9)),))))
            for msg, f in futs:
                assert f.result() == msg
    srv.close()
    assert not srv.errors
    return [c.hex() for c in srv.capture[0]], srv.n_requests


def main():
    out = {}
    for kind in ("sequential", "concurrent"):
        chunks, n = record(kind)
        out[kind] = {"chunks": chunks, "n_requests": n}
        print(kind, "chunks", len(chunks), "bytes", sum(len(c) // 2 for c in chunks), "requests", n)
    with open(os.path.join(HERE, "h2_grpcio_capture.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
