"""A minimal TCP server loop for interop tests: socket bytes -> an h2 engine (the CPU oracle's H2Conn, or the device through
b2_h2_process_batch / b2_h2_pack_responses) -> the engine's control bytes and packed gRPC replies written back.
The peer is a REAL grpcio client (gRPC C-core's h2 stack): an independent implementation that rejects malformed SETTINGS /
HEADERS / DATA / trailers, so a completed call pins the framing the engine produced."""
import socket
import threading

import numpy as np


class OracleEngine:
    """One oracle H2Conn per TCP connection (tests only)."""
    def __init__(self, O):
        self.O = O
        self.conns = {}

    def open(self, cid):
        self.conns[cid] = self.O.H2Conn()

    def feed(self, cid, buf):
        """Returns (consumed, bytes to write, parse_error, n_requests)."""
        c = self.conns[cid]
        err, cons, msgs, ctrl, blob, _, _ = c.consume(buf)
        out = [ctrl]
        for m in msgs:
            ok = (m["flags"] & 3) == 3 and m["method_idx"] >= 0
            body = bytes(blob[m["msg_off"]:m["msg_off"] + m["msg_len"]]) if ok else b""
            out.append(c.pack_response(int(m["stream_id"]), body, grpc_status=0 if ok else 12,
                                       grpc_message=b"" if ok else b"unimplemented"))
        return cons, b"".join(out), err, len(msgs)


class DeviceEngine:
    """Connections share one b2 context; every feed is one b2_h2_process_batch + one b2_h2_pack_responses (echo by reference:
    the reply body is the request message still on the device)."""
    def __init__(self, ctx):
        import brpc_b200
        from brpc_b200.abi import H2_RESPONSE_DT
        self.ctx, self.b2, self.RDT = ctx, brpc_b200, H2_RESPONSE_DT
        self.lock = threading.Lock()

    def open(self, cid):
        with self.lock:
            self.ctx.h2_conn_reset(cid)

    def feed(self, cid, buf):
        with self.lock:
            data, runs = self.b2.make_runs([buf]); runs["socket_id"] = cid
            rs, msgs, out = self.ctx.h2_process_batch(data, runs, msg_cap=1024, out_cap=4 << 20)
            ctrl = bytes(out[int(rs["ctrl_off"][0]):int(rs["ctrl_off"][0]) + int(rs["ctrl_len"][0])])
            reply = b""
            if len(msgs):
                ct = b"application/grpc"; gm = b"unimplemented"
                r = np.zeros(len(msgs), self.RDT)
                ok = ((msgs["flags"] & 3) == 3) & (msgs["method_idx"] >= 0)
                r["conn"] = cid; r["stream_id"] = msgs["stream_id"]; r["status_code"] = 200
                r["flags"] = 1 | np.where(ok, np.where(msgs["flags"] & 16, 2, 4), 0)
                r["content_type_off"] = 0; r["content_type_len"] = len(ct)
                r["body_off"] = np.where(ok, msgs["msg_off"], 0); r["body_len"] = np.where(ok, msgs["msg_len"], 0)
                r["grpc_status"] = np.where(ok, 0, 12)
                r["grpc_message_off"] = len(ct); r["grpc_message_len"] = np.where(ok, 0, len(gm))
                frames = self.ctx.h2_pack_responses(np.frombuffer(ct + gm + bytes(16), np.uint8), r)
                reply = b"".join(frames)
            return int(rs["consumed"][0]), ctrl + reply, int(rs["parse_error"][0]), len(msgs)


class H2LoopServer:
    """accept() -> per-connection thread: recv, engine.feed(pending bytes), send what it returns.  `capture`: every
    connection's inbound byte chunks, for replay against another engine."""
    def __init__(self, engine):
        self.engine = engine
        self.sock = socket.socket(); self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.sock.bind(("127.0.0.1", 0)); self.sock.listen(64)
        self.port = self.sock.getsockname()[1]
        self.capture = {}; self.errors = []; self.n_requests = 0
        self._next = 0; self._stop = False
        self._t = threading.Thread(target=self._accept, daemon=True); self._t.start()

    def _accept(self):
        while not self._stop:
            try:
                s, _ = self.sock.accept()
            except OSError:
                return
            cid = self._next; self._next += 1
            self.capture[cid] = []
            self.engine.open(cid)
            threading.Thread(target=self._serve, args=(s, cid), daemon=True).start()

    def _serve(self, s, cid):
        pending = b""
        try:
            while True:
                chunk = s.recv(1 << 16)
                if not chunk:
                    return
                self.capture[cid].append(chunk)
                pending += chunk
                cons, out, err, n = self.engine.feed(cid, pending)
                pending = pending[cons:]
                self.n_requests += n
                if out:
                    s.sendall(out)
                if err != 2:                                    # anything but NOT_ENOUGH_DATA closes the socket (input_messenger.cpp:227-239)
                    self.errors.append((cid, err)); return
        except Exception as e:                                  # noqa: BLE001
            self.errors.append((cid, repr(e)))
        finally:
            s.close()

    def close(self):
        self._stop = True
        self.sock.close()
