"""CPU: pins the CLIENT side of the h2 oracle (orc_h2_pack_request = H2UnsentRequest::New + AppendAndDestroySelf + PackH2Message)
with an independent implementation on the other end: a real grpcio SERVER (gRPC C-core) reads the oracle's bytes from a TCP
socket — preface, SETTINGS, WINDOW_UPDATE, HEADERS with the encoder's dynamic table, DATA split at max_frame_size — and must
answer every call with the echoed message and grpc-status 0."""
from concurrent import futures
import socket

import pytest

import _oracle as O

grpc = pytest.importorskip("grpc")

SIZES = [0, 1, 5, 100, 1000, 4096, 16384 - 5, 16384, 16385, 40000, 65536, 200000]
GRPC_EXTRA = ((b"te", b"trailers"), (b"grpc-accept-encoding", b"identity,gzip"))


class _Echo(grpc.GenericRpcHandler):
    def __init__(self):
        self.seen = []

    def service(self, details):
        if details.method != "/example.EchoService/Echo":
            return None
        md = dict(details.invocation_metadata)
        self.seen.append(md)
        return grpc.unary_unary_rpc_method_handler(lambda req, ctx: req, request_deserializer=lambda b: b, response_serializer=lambda b: b)


def _server():
    h = _Echo()
    srv = grpc.server(futures.ThreadPoolExecutor(max_workers=8), handlers=[h],
                      options=[("grpc.max_receive_message_length", 1 << 24), ("grpc.max_send_message_length", 1 << 24)])
    port = srv.add_insecure_port("127.0.0.1:0")
    srv.start()
    return srv, port, h


class FrameReader:
    """Splits the server's bytes into h2 frames; answers SETTINGS / PING like a client must so that the server keeps going."""
    def __init__(self, sock, conn=None):
        self.s = sock; self.buf = b""; self.conn = conn; self.first = True; self.mirrored = []

    def mirror(self, ftype, flags, sid, payload):
        """What the host's H2Context would hand to b2_h2_conn_peer_update: the server's SETTINGS and connection WINDOW_UPDATEs."""
        if self.conn is None:
            return
        if ftype == 4 and not (flags & 1):
            u = {}
            for k in range(0, len(payload), 6):
                ident, v = int.from_bytes(payload[k:k + 2], "big"), int.from_bytes(payload[k + 2:k + 6], "big")
                name = {1: "header_table_size", 4: "stream_window_size", 5: "max_frame_size"}.get(ident)
                if name:
                    u[name] = v
            if self.first:
                u["conn_window_add"] = -(0x7fffffff - 65535); self.first = False
            if u:
                assert self.conn.peer_update(**u) == 0; self.mirrored.append(u)
        elif ftype == 8 and sid == 0:
            u = dict(conn_window_add=int.from_bytes(payload, "big") & 0x7fffffff)
            assert self.conn.peer_update(**u) == 0; self.mirrored.append(u)

    def frames(self, until):
        while not until():
            while len(self.buf) >= 9:
                n = int.from_bytes(self.buf[:3], "big")
                if len(self.buf) < 9 + n:
                    break
                f = (self.buf[3], self.buf[4], int.from_bytes(self.buf[5:9], "big") & 0x7fffffff, self.buf[9:9 + n])
                self.buf = self.buf[9 + n:]
                if f[0] == 4 and not (f[1] & 1):
                    self.s.sendall(b"\0\0\0\x04\x01\0\0\0\0")              # SETTINGS ack
                if f[0] == 6 and not (f[1] & 1):
                    self.s.sendall(b"\0\0\x08\x06\x01\0\0\0\0" + f[3])       # PING ack
                self.mirror(*f)
                yield f
                if until():
                    return
            d = self.s.recv(1 << 16)
            if not d:
                raise AssertionError("server closed the connection")
            self.buf += d


def _run_calls(conn, sock, calls, window_updates=True, rd=None):
    """Sends every call through the oracle's client side, collects {stream: [DATA bytes, ended]} from the server."""
    rd = rd or FrameReader(sock)
    got = {}; bad = []; sent = []

    def done():
        return bad or (len(sent) == len(calls) and all(got.get(s, [b"", False])[1] for s in sent))
    it = rd.frames(done)
    for path, body, extra in calls:
        st, sid, b = conn.pack_request(path, b"127.0.0.1:1", body, extra=extra)
        assert st == 0, st
        sock.sendall(b); sent.append(sid)
    for ftype, flags, sid, payload in it:
        if ftype == 0:
            got.setdefault(sid, [b"", False])[0] += payload
            if window_updates and payload:                                       # give the server its send window back
                inc = len(payload).to_bytes(4, "big")
                sock.sendall(b"\0\0\x04\x08\0\0\0\0\0" + inc + b"\0\0\x04\x08\0" + sid.to_bytes(4, "big") + inc)
        elif ftype == 1:
            got.setdefault(sid, [b"", False])
        elif ftype in (3, 7):
            bad.append((ftype, sid, payload))
        if flags & 1 and ftype in (0, 1):
            got[sid][1] = True
    assert not bad, bad
    return sent, got


def test_oracle_client_requests_against_a_grpcio_server():
    srv, port, h = _server()
    try:
        conn = O.H2Conn()
        with socket.create_connection(("127.0.0.1", port)) as s:
            s.settimeout(20)
            bodies = [bytes([97 + i % 26]) * SIZES[i % len(SIZES)] for i in range(60)]
            calls = [(b"/example.EchoService/Echo", b, GRPC_EXTRA + ((b"x-call", b"c%d" % (i % 5)), (b"x-fixed", b"same-value-every-time")))
                     for i, b in enumerate(bodies)]
            # one at a time first (the server's SETTINGS are not fed back: the oracle keeps the default peer settings, which every
            # h2 server must accept), then a burst of concurrent streams
            sent = []; got = {}
            for c in calls[:24]:
                s1, g1 = _run_calls(conn, s, [c]); sent += s1; got.update(g1)
            s2, g2 = _run_calls(conn, s, calls[24:]); sent += s2; got.update(g2)
        assert sent == [1 + 2 * i for i in range(60)]
        for sid, body in zip(sent, bodies):
            data, ended = got[sid]
            assert ended and data == b"\0" + len(body).to_bytes(4, "big") + body, (sid, len(body), len(data))
        assert len(h.seen) == 60
        for i, md in enumerate(h.seen[:24]):                                     # what the C-core HPACK decoder made of the header blocks
            assert md["x-call"] == "c%d" % (i % 5) and md["x-fixed"] == "same-value-every-time" and md["user-agent"].startswith("brpc/1.0 curl/7.0")
    finally:
        srv.stop(0)


def test_oracle_client_stream_ids_run_out_and_windows():
    conn = O.H2Conn()
    conn.set_next_stream_id(0x7ffffffd)
    st, sid, b = conn.pack_request(b"/a/b", b"h:1", b"x")
    assert (st, sid) == (0, 0x7ffffffd) and b.startswith(b"PRI * HTTP/2.0")
    st, sid, b = conn.pack_request(b"/a/b", b"h:1", b"x")
    assert (st, sid) == (0, 0x7fffffff) and not b.startswith(b"PRI")
    st, sid, b = conn.pack_request(b"/a/b", b"h:1", b"x")
    assert (st, sid, b) == (2, 0, b"")                                           # EH2RUNOUTSTREAMS
    # the connection window: 2^31 - 1 until the peer says otherwise; each body is charged to it
    conn = O.H2Conn()
    big = bytes(1 << 20)
    n_ok = 0
    for i in range(2100):
        st, sid, b = conn.pack_request(b"/a/b", b"h:1", big, flags=8 | 16)
        assert sid == 1 + 2 * i
        if st != 0:
            assert st == 1 and (b == b"" or i == 0)
            break
        n_ok += 1
    assert n_ok == 2047                                                          # floor((2^31 - 1) / 2^20)


def test_oracle_client_follows_the_servers_settings():
    """The same, with the server's SETTINGS and connection WINDOW_UPDATEs mirrored into the oracle's connection (what a host-side parser
    hands to b2_h2_conn_peer_update): bodies beyond the server's initial 65535-byte windows wait for nothing here because C-core opens
    its windows at once; a body larger than the stream window is refused with ELIMIT instead of being sent."""
    srv, port, h = _server()
    try:
        conn = O.H2Conn()
        with socket.create_connection(("127.0.0.1", port)) as s:
            s.settimeout(20)
            rd = FrameReader(s, conn)
            sent, got = _run_calls(conn, s, [(b"/example.EchoService/Echo", b"first", GRPC_EXTRA)], rd=rd)
            assert rd.mirrored and "conn_window_add" in rd.mirrored[0]           # the server's SETTINGS came with the first reply
            bodies = [bytes([65 + i % 26]) * n for i, n in enumerate([10, 30000, 16384, 100, 50000, 0, 7])]
            s2, g2 = _run_calls(conn, s, [(b"/example.EchoService/Echo", b, GRPC_EXTRA) for b in bodies], rd=rd)
            for sid, body in zip(s2, bodies):
                assert g2[sid] == [b"\0" + len(body).to_bytes(4, "big") + body, True]
            assert conn.peer_update(stream_window_size=1000) == 0                 # (as if the server had shrunk its stream windows)
            st, sid, b = conn.pack_request(b"/example.EchoService/Echo", b"127.0.0.1:1", bytes(2000), extra=GRPC_EXTRA)
            assert st == 1 and b == b""                                          # refused with ELIMIT instead of being sent
    finally:
        srv.stop(0)
