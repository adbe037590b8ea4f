"""Host-side mirror of brpc's InputMessenger for the GPU path.

`GpuInputMessenger` plays the role brpc::InputMessenger plays for a set of
Sockets (src/brpc/input_messenger.h:73-160): sockets append bytes to their read
buffer (`feed`, == Socket::DoRead into Socket::_read_buf), `poll()` hands all
pending bytes to the device in one batch (== one OnNewMessages round over every
readable socket, input_messenger.cpp:324-389) and then pops exactly the consumed
bytes off each read buffer (ParseResult contract, protocol.h:82-92), keeps the
partial frame tail and the preferred protocol index for the next round, and
fails the socket on any parse error other than NOT_ENOUGH_DATA
(input_messenger.cpp:227-239).  No parsing happens here; the cut loop, meta
decode, echo service and response packing all run in the CUDA kernels.
"""
import numpy as np

from .abi import RUN_DT, Context

PARSE_ERROR_NOT_ENOUGH_DATA = 2


def make_runs(chunks, align=16):
    """Concatenate per-socket byte strings into a batch buffer with 16-byte aligned runs."""
    offs, total = [], 0
    for c in chunks:
        offs.append(total)
        total += (len(c) + align - 1) // align * align
    data = np.zeros(max(total, 16), dtype=np.uint8)
    runs = np.zeros(len(chunks), dtype=RUN_DT)
    for i, c in enumerate(chunks):
        n = len(c)
        if n:
            data[offs[i]:offs[i] + n] = np.frombuffer(bytes(c), dtype=np.uint8) if not isinstance(c, np.ndarray) else c
        runs[i] = (i, offs[i], n, -1, 0)
    return data, runs


class Socket:
    __slots__ = ("id", "read_buf", "preferred_index", "failed", "error", "in_msgs", "in_bytes")

    def __init__(self, sid):
        self.id, self.read_buf, self.preferred_index = sid, bytearray(), -1
        self.failed, self.error, self.in_msgs, self.in_bytes = False, 0, 0, 0


class GpuInputMessenger:
    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.sockets = {}

    def add_socket(self, sid):
        self.sockets[sid] = Socket(sid)
        return self.sockets[sid]

    def feed(self, sid, data):
        s = self.sockets[sid]
        if not s.failed:
            s.read_buf += data
            s.in_bytes += len(data)

    def poll(self):
        """One batch over every socket with pending bytes.
        Returns a list of (socket_id, msg_desc_record, response_bytes) in per-socket order."""
        live = [s for s in self.sockets.values() if s.read_buf and not s.failed]
        if not live:
            return []
        data, runs = make_runs([s.read_buf for s in live])
        for i, s in enumerate(live):
            runs[i]["socket_id"] = s.id
            runs[i]["preferred_proto"] = s.preferred_index
        rs, msgs, resp, _ = self.ctx.process_batch(data, runs)
        out = []
        for i, s in enumerate(live):
            st = rs[i]
            del s.read_buf[:int(st["consumed"])]
            s.preferred_index = int(st["preferred_proto"])
            s.in_msgs += int(st["n_msgs"])
            if st["parse_error"] != PARSE_ERROR_NOT_ENOUGH_DATA:
                s.failed, s.error = True, int(st["parse_error"])   # Socket::SetFailed(EINVAL, ...)
            for m in msgs[int(st["first_msg"]):int(st["first_msg"]) + int(st["n_msgs"])]:
                body = bytes(resp[int(m["resp_off"]):int(m["resp_off"]) + int(m["resp_len"])])
                out.append((s.id, m.copy(), body))
        return out
