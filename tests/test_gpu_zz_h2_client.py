"""GPU: the client side of h2 (b2_h2_pack_requests = H2UnsentRequest::New + AppendAndDestroySelf + PackH2Message) against the oracle
(pinned by a real grpcio server, tests/test_oracle_h2_client_grpcio.py): preface + settings with the first request, stream ids, HPACK
encoder state across batches (names and whole headers indexed as they repeat, eviction with long values), DATA split at the peer's
max_frame_size after b2_h2_conn_peer_update, never-indexed headers for header_table_size 0, flow control (ELIMIT), id exhaustion
(EH2RUNOUTSTREAMS)."""
import random

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu
SEED = 20260922
GRPC_EXTRA = ((b"te", b"trailers"), (b"grpc-accept-encoding", b"identity,gzip"))


def _ctx():
    import brpc_b200
    return brpc_b200.Context(device=0, max_batch_bytes=32 << 20, max_msgs=1 << 14, max_runs=256, max_resp_bytes=64 << 20)


def _call(rng, conn):
    body = bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.choice([0, 0, 7, 300, 4096, 16379, 16380, 40000])))
    flags = rng.choice([1 | 8 | 16, 1 | 8 | 16, 1 | 8 | 16, 8 | 16, 0, 2 | 8, 4 | 16 | 1])
    path = rng.choice([b"/example.EchoService/Echo", b"/example.EchoService/Echo", b"/a.B/C%d" % rng.randrange(40), b"/" + b"p" * rng.randrange(1, 300)])
    extra = list(GRPC_EXTRA) if flags & 1 else []
    for _ in range(rng.randrange(0, 4)):
        extra.append((rng.choice([b"x-trace", b"X-Mixed-Case", b"grpc-timeout", b"authorization", b"x-" + b"n" * rng.randrange(1, 60)]),
                      rng.choice([b"", b"1S", b"same-value-every-time", b"v" * rng.randrange(1, 400), b"t%d" % rng.randrange(6)])))
    ct = rng.choice([b"application/grpc", b"application/grpc", b"application/json", b""])
    return (conn, flags, path, b"10.1.2.%d:8000" % (conn % 7), ct, body, tuple(extra))


def test_h2_client_requests_packed_on_the_device():
    from brpc_b200.abi import H2_REQUEST_DT
    rng = random.Random(SEED)
    ctx = _ctx()
    n_conns = 24
    orc = [O.H2Conn() for _ in range(n_conns)]
    for i in range(n_conns):
        ctx.h2_conn_reset(i)
    n = 0; n_pre = 0; n_multi = 0
    for rnd in range(12):
        calls = []
        for i in rng.sample(range(n_conns), rng.randrange(1, n_conns + 1)):
            calls += [_call(rng, i) for _ in range(rng.randrange(1, 6))]
        blob, reqs = O.h2_request_blob(calls)
        res, got = ctx.h2_pack_requests(np.frombuffer(blob, np.uint8), reqs.astype(H2_REQUEST_DT))
        for k, c in enumerate(calls):
            st, sid, want = orc[c[0]].pack_request(c[2], c[3], c[5], content_type=c[4], flags=c[1], extra=c[6])
            assert (int(res[k]["status"]), int(res[k]["stream_id"])) == (st, sid), (rnd, k)
            assert got[k] == want, (rnd, k, c[:5])
            n += 1; n_pre += want.startswith(b"PRI"); n_multi += len(c[5]) > 16384
    assert n > 200 and n_pre == n_conns and n_multi > 10


def test_h2_client_requests_follow_the_peers_settings_and_windows():
    """The host parses the server's frames (the receiving half of a client connection is not on the device) and mirrors SETTINGS /
    WINDOW_UPDATE with b2_h2_conn_peer_update; what the client packs afterwards follows them: max_frame_size 20000, header_table_size 0
    (never indexed), a small stream window and a drained connection window (ELIMIT), updates between batches."""
    import brpc_b200
    from brpc_b200.abi import H2_REQUEST_DT
    rng = random.Random(SEED + 1)
    ctx = _ctx()
    first = -(0x7fffffff - 65535)                                  # OnSettings, first SETTINGS frame (:884)
    peers = [dict(max_frame_size=20000, conn_window_add=first), dict(header_table_size=0, conn_window_add=first),
             dict(stream_window_size=1000, conn_window_add=first), dict(max_frame_size=70000, stream_window_size=1 << 20, conn_window_add=first), {}]
    orc = [O.H2Conn() for _ in peers]
    for i, p in enumerate(peers):
        ctx.h2_conn_reset(i)
        if p:
            ctx.h2_conn_peer_update(i, **p); assert orc[i].peer_update(**p) == 0
    n_elimit = 0; n_ok = 0
    for rnd in range(6):
        calls = [_call(rng, i) for i in range(len(peers)) for _ in range(5)]
        blob, reqs = O.h2_request_blob(calls)
        res, got = ctx.h2_pack_requests(np.frombuffer(blob, np.uint8), reqs.astype(H2_REQUEST_DT))
        for k, c in enumerate(calls):
            st, sid, want = orc[c[0]].pack_request(c[2], c[3], c[5], content_type=c[4], flags=c[1], extra=c[6])
            assert (int(res[k]["status"]), int(res[k]["stream_id"]), got[k]) == (st, sid, want), (rnd, k)
            n_elimit += st == 1; n_ok += st == 0
        if rnd == 2:                                               # the peer opens its windows / changes its mind between batches
            for i in (0, 2, 3):
                u = dict(conn_window_add=rng.randrange(1, 1 << 20), stream_window_size=65535 + 1000 * i, header_table_size=256 * i)
                ctx.h2_conn_peer_update(i, **u); assert orc[i].peer_update(**u) == 0
    assert n_elimit > 10 and n_ok > 60
    # AddWindowSize past 2^31 - 1 is FLOW_CONTROL_ERROR in OnWindowUpdate; out-of-range settings are what ParseH2Settings refuses
    for bad in (dict(conn_window_add=0x7fffffff), dict(max_frame_size=16383), dict(max_frame_size=1 << 24), dict(stream_window_size=1 << 31)):
        with pytest.raises(brpc_b200.B2Error):
            ctx.h2_conn_peer_update(4, **bad)
        assert orc[4].peer_update(**bad) == -1


def test_h2_client_stream_ids_run_out():
    from brpc_b200.abi import H2_REQUEST_DT
    ctx = _ctx()
    ctx.h2_conn_reset(3)
    ctx.h2_conn_set_next_stream_id(3, 0x7ffffffd)
    conn = O.H2Conn(); conn.set_next_stream_id(0x7ffffffd)
    calls = [(3, 1 | 8 | 16, b"/a/b", b"h:1", b"application/grpc", b"x", GRPC_EXTRA)] * 4
    blob, reqs = O.h2_request_blob(calls)
    res, got = ctx.h2_pack_requests(np.frombuffer(blob, np.uint8), reqs.astype(H2_REQUEST_DT))
    want = [conn.pack_request(b"/a/b", b"h:1", b"x", extra=GRPC_EXTRA) for _ in calls]
    assert [(int(r["status"]), int(r["stream_id"])) for r in res] == [(w[0], w[1]) for w in want] == [(0, 0x7ffffffd), (0, 0x7fffffff), (2, 0), (2, 0)]
    assert got == [w[2] for w in want]


def test_h2_client_bad_descriptors_are_refused():
    import brpc_b200
    from brpc_b200.abi import H2_REQUEST_DT
    ctx = _ctx()
    blob, reqs = O.h2_request_blob([(0, 1, b"/a/b", b"h:1", b"application/grpc", b"x", ()), (1, 1, b"/a/b", b"h:1", b"", b"", ()), (0, 1, b"/a/b", b"h:1", b"", b"", ())])
    with pytest.raises(brpc_b200.B2Error):                        # requests of one connection must be adjacent
        ctx.h2_pack_requests(np.frombuffer(blob, np.uint8), reqs.astype(H2_REQUEST_DT))
    blob, reqs = O.h2_request_blob([(0, 1, b"/" + b"p" * 3000, b"h:1", b"", b"", ())])
    with pytest.raises(brpc_b200.B2Error):                        # a header block beyond the kernel's fragment buffer
        ctx.h2_pack_requests(np.frombuffer(blob, np.uint8), reqs.astype(H2_REQUEST_DT))
    reqs["body_len"] = 1 << 30
    with pytest.raises(brpc_b200.B2Error):
        ctx.h2_pack_requests(np.frombuffer(blob, np.uint8), reqs.astype(H2_REQUEST_DT))
