"""h2 client-side traffic for the parser tests: a small HPACK encoder that tracks the dynamic table the way
RFC 7541 says a peer's decoder will (so indexed references stay valid), frame builders, and a seeded
generator of gRPC-ish connections with control frames and (optionally) protocol violations mixed in."""
import random

PREFACE = b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n"
STATIC = [(b":authority", b""), (b":method", b"GET"), (b":method", b"POST"), (b":path", b"/"), (b":path", b"/index.html"),
          (b":scheme", b"http"), (b":scheme", b"https"), (b":status", b"200"), (b":status", b"204"), (b":status", b"206"),
          (b":status", b"304"), (b":status", b"400"), (b":status", b"404"), (b":status", b"500"), (b"accept-charset", b""),
          (b"accept-encoding", b"gzip, deflate"), (b"accept-language", b""), (b"accept-ranges", b""), (b"accept", b""),
          (b"access-control-allow-origin", b""), (b"age", b""), (b"allow", b""), (b"authorization", b""), (b"cache-control", b""),
          (b"content-disposition", b""), (b"content-encoding", b""), (b"content-language", b""), (b"content-length", b""),
          (b"content-location", b""), (b"content-range", b""), (b"content-type", b""), (b"cookie", b""), (b"date", b""), (b"etag", b""),
          (b"expect", b""), (b"expires", b""), (b"from", b""), (b"host", b""), (b"if-match", b""), (b"if-modified-since", b""),
          (b"if-none-match", b""), (b"if-range", b""), (b"if-unmodified-since", b""), (b"last-modified", b""), (b"link", b""),
          (b"location", b""), (b"max-forwards", b""), (b"proxy-authenticate", b""), (b"proxy-authorization", b""), (b"range", b""),
          (b"referer", b""), (b"refresh", b""), (b"retry-after", b""), (b"server", b""), (b"set-cookie", b""),
          (b"strict-transport-security", b""), (b"transfer-encoding", b""), (b"user-agent", b""), (b"vary", b""), (b"via", b""),
          (b"www-authenticate", b"")]
assert len(STATIC) == 61


def hp_int(value, prefix, first):
    lim = (1 << prefix) - 1
    if value < lim:
        return bytes([first | value])
    out = bytearray([first | lim]); value -= lim
    while value >= 128:
        out.append((value & 0x7F) | 0x80); value >>= 7
    out.append(value)
    return bytes(out)


def hp_str(b):
    return hp_int(len(b), 7, 0) + bytes(b)


class HpackEncoder:
    def __init__(self, rng, max_size=4096):
        self.rng = rng; self.max_size = max_size; self.dyn = []      # newest first
        self.fixed_mode = None                                       # e.g. "auto": indexed when possible, else incremental

    def _size(self):
        return sum(len(n) + len(v) + 32 for n, v in self.dyn)

    def _add(self, n, v):
        es = len(n) + len(v) + 32
        while self.dyn and self._size() + es > self.max_size:
            self.dyn.pop()
        if es <= self.max_size:
            self.dyn.insert(0, (n, v))

    def _find(self, n, v):
        full = name = 0
        for i, (sn, sv) in enumerate(STATIC + self.dyn):
            if sn == n:
                if not name: name = i + 1
                if sv == v and not full: full = i + 1
        return full, name

    def field(self, n, v, mode=None):
        full, name = self._find(n, v)
        mode = mode or self.fixed_mode or self.rng.choice(["auto", "auto", "auto", "incr", "noidx", "never"])
        if full and mode == "auto":
            return hp_int(full, 7, 0x80)
        if mode in ("auto", "incr"):
            out = hp_int(name, 6, 0x40) + (b"" if name else hp_str(n)) + hp_str(v)
            if len(self.dyn) < 100:                                    # stay clear of the reference's 120-entry queue
                self._add(n, v)
                return out
            mode = "noidx"
        first = 0x10 if mode == "never" else 0x00
        return hp_int(name, 4, first) + (b"" if name else hp_str(n)) + hp_str(v)


def frame(ftype, flags, sid, payload=b""):
    return len(payload).to_bytes(3, "big") + bytes([ftype, flags]) + sid.to_bytes(4, "big") + bytes(payload)


def settings(pairs=(), ack=False):
    return frame(4, 1 if ack else 0, 0, b"".join(i.to_bytes(2, "big") + v.to_bytes(4, "big") for i, v in pairs))


def grpc_body(message, compressed=0):
    pb = b"\x0a" + _varint(len(message)) + message
    return bytes([compressed]) + len(pb).to_bytes(4, "big") + pb


def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80); v >>= 7
    out.append(v)
    return bytes(out)


def request_frames(rng, enc, sid, path=b"/example.EchoService/Echo", message=b"hello", content_type=b"application/grpc",
                   extra=(), split_headers=False, pad=False, priority=False, chunk=None, trailers=False):
    """HEADERS (+CONTINUATION) + DATA... for one unary call; returns the list of frames (bytes)."""
    fields = [enc.field(b":method", b"POST"), enc.field(b":scheme", b"http"), enc.field(b":path", path),
              enc.field(b":authority", b"127.0.0.1:8010"), enc.field(b"content-type", content_type), enc.field(b"te", b"trailers")]
    fields += [enc.field(n, v) for n, v in extra]
    out = []
    body = grpc_body(message) if message is not None else b""
    end_on_headers = message is None
    if split_headers and len(fields) > 2:
        k = rng.randrange(1, len(fields))
        first, rest = b"".join(fields[:k]), b"".join(fields[k:])
        out.append(_headers(rng, sid, first, end_headers=False, end_stream=end_on_headers, pad=pad, priority=priority))
        out.append(frame(9, 0x4, sid, rest))
    else:
        out.append(_headers(rng, sid, b"".join(fields), end_headers=True, end_stream=end_on_headers, pad=pad, priority=priority))
    if not end_on_headers:
        chunk = min(chunk or rng.choice([len(body) or 1, 64, 1000, 4000]), 16000)      # a frame (with padding) never exceeds the default max_frame_size
        pieces = [body[i:i + chunk] for i in range(0, len(body), chunk)] or [b""]
        for j, pc in enumerate(pieces):
            last = j == len(pieces) - 1
            fl = 0x1 if (last and not trailers) else 0
            if pad and rng.random() < 0.5:
                pl = rng.randrange(0, 20)
                out.append(frame(0, fl | 0x8, sid, bytes([pl]) + pc + bytes(pl)))
            else:
                out.append(frame(0, fl, sid, pc))
        if trailers:
            out.append(frame(1, 0x5, sid, b"\x00" + hp_str(b"x-trailer") + hp_str(b"done")))   # literal, no indexing: emitted after later HEADERS
    return out


def _headers(rng, sid, block, end_headers, end_stream, pad, priority):
    fl = (0x4 if end_headers else 0) | (0x1 if end_stream else 0)
    pre = b""; post = b""
    if pad:
        pl = rng.randrange(0, 16); fl |= 0x8; pre += bytes([pl]); post = bytes(pl)
    if priority:
        fl |= 0x20; pre += rng.randrange(1 << 31).to_bytes(4, "big") + bytes([rng.randrange(256)])
    return frame(1, fl, sid, pre + block + post)


def connection_script(rng, n_calls=12, violations=0.0, max_open=4):
    """A whole client connection as a list of frames (first element is the preface)."""
    enc = HpackEncoder(rng)
    out = [PREFACE, settings(rng.choice([(), ((3, 100), (4, 65535)), ((4, 1 << 20), (5, 32768)), ((1, 4096), (2, 0))]))]
    sid = 1
    pending = []                                   # frames of calls whose emission is interleaved
    rnd62 = b"abcdefghijklmnopqrstuvwxyz0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"
    for call in range(n_calls):
        msg = bytes(rng.choice(rnd62) for _ in range(rng.choice([0, 5, 100, 1024, 4096, 9000, 30000])))
        kind = rng.random()
        kw = dict(split_headers=rng.random() < 0.3, pad=rng.random() < 0.3, priority=rng.random() < 0.2, trailers=rng.random() < 0.15)
        if kind < 0.08:
            fr = request_frames(rng, enc, sid, message=None, **{k: v for k, v in kw.items() if k != "trailers"})
        elif kind < 0.16:
            fr = request_frames(rng, enc, sid, path=rng.choice([b"/EchoService/Echo", b"/example.EchoService/Nope", b"/nope.Svc/Echo?x=1", b"//example.EchoService//Echo#frag", b"/"]),
                                message=msg, **kw)
        elif kind < 0.24:
            fr = request_frames(rng, enc, sid, content_type=rng.choice([b"application/grpc+proto", b"application/grpc+json", b"application/json",
                                                                        b"application/grpcfoo", b"text/plain", b"application/grpc;charset=x", b"application/proto"]),
                                message=msg, **kw)
        else:
            fr = request_frames(rng, enc, sid, message=msg, extra=[(b"grpc-timeout", b"1S"), (b"x-req-%d" % (call % 3), b"v%d" % call)][:rng.randrange(3)], **kw)
        out.append(fr.pop(0))                          # stream ids must grow: a call's HEADERS (+CONTINUATION) go out at once,
        while fr and fr[0][3] == 9:                    # its DATA / trailers may interleave with later calls
            out.append(fr.pop(0))
        if fr: pending.append(fr)
        sid += 2
        # interleave: flush some frames of the open calls in round-robin order, HEADERS/CONTINUATION kept adjacent
        while len(pending) > rng.randrange(max_open):
            q = rng.randrange(len(pending))
            f = pending[q].pop(0); out.append(f)
            while pending[q] and pending[q][0][3] == 9:
                out.append(pending[q].pop(0))
            if not pending[q]: pending.pop(q)
        r = rng.random()
        if r < 0.10: out.append(frame(6, 0, 0, bytes(rng.randrange(256) for _ in range(8))))          # PING
        elif r < 0.16: out.append(frame(8, 0, rng.choice([0, max(1, sid - 2)]), rng.randrange(1, 1 << 20).to_bytes(4, "big")))
        elif r < 0.20: out.append(settings(ack=True))
        elif r < 0.24: out.append(settings([(4, rng.choice([65535, 1 << 16, 1 << 24]))]))
        elif r < 0.27 and not pending:                                                                # a call the client gives up: HEADERS, some DATA, RST_STREAM(CANCEL)
            fr = request_frames(rng, enc, sid, message=b"abandoned" * 50, chunk=100)
            out.extend(fr[:rng.randrange(1, len(fr))]); out.append(frame(3, 0, sid, (8).to_bytes(4, "big"))); sid += 2
        elif r < 0.29: out.append(frame(7, 0, 0, (sid).to_bytes(4, "big") + (0).to_bytes(4, "big") + b"bye"))
        if violations and rng.random() < violations:
            v = rng.randrange(13)
            if v == 0: out.append(frame(2, 0, 1, bytes(5)))                                        # PRIORITY -> GOAWAY, payload left unread
            elif v == 1: out.append(frame(0, 0x1, sid + 100, b"stray"))                            # DATA on an unknown stream
            elif v == 2: out.append(frame(6, 0x1, 0, bytes(8)))                                    # PING ack: payload left unread
            elif v == 3: out.append(frame(4, 0, 0, bytes(5)))                                      # SETTINGS with a bad length
            elif v == 4: out.append(frame(8, 0, 0, bytes(4)))                                      # WINDOW_UPDATE 0
            elif v == 5: out.append(frame(1, 0x5, 2, enc.field(b":method", b"POST")))              # even stream id
            elif v == 6: out.append(frame(1, 0x5, sid, b"\x40\x01:\x01x")); sid += 2              # unknown pseudo header ":" ...
            elif v == 7: out.append(frame(9, 0x4, sid + 50, b""))                                   # CONTINUATION without a stream
            elif v == 8: out.append(frame(1, 0x4, 0, b""))                                          # HEADERS on stream 0
            elif v == 9: out.append(frame(4, 0, 0, (2).to_bytes(2, "big") + (7).to_bytes(4, "big")))   # ENABLE_PUSH = 7
            elif v == 10: out.append(frame(5, 0, 1, b""))                                           # PUSH_PROMISE
            elif v == 11: out.append(frame(1, 0x5, sid, b"\x83\xbe")); sid += 2                    # index past the table
            elif v == 12: out.append(frame(0, 0x8, max(1, sid - 2), b""))                           # PADDED DATA with no room for the pad length
    for fr in pending:
        out.extend(fr)
    return out
