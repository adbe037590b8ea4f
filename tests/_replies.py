"""Building b2_reply batches (REPLY_DT records + the byte buffer they point into) for tests of b2_pack_responses."""
import struct

import numpy as np

from brpc_b200.abi import REPLY_DT


class ReplyBatch:
    def __init__(self):
        self.buf = bytearray(); self.recs = []

    def _put(self, b, align=1):
        while len(self.buf) % align:
            self.buf.append(0)
        off = len(self.buf); self.buf += b
        return off, len(b)

    def add(self, error_code=0, error_text=b"", body=b"", attachment=b"", compress_type=0, checksum_type=0, content_type=0, correlation_id=0,
            request_checksum=b"", stream=None, user_fields=()):
        r = np.zeros(1, REPLY_DT)[0]
        r["error_code"] = error_code; r["correlation_id"] = correlation_id
        r["compress_type"] = compress_type; r["checksum_type"] = checksum_type; r["content_type"] = content_type
        r["error_text_off"], r["error_text_len"] = self._put(bytes(error_text))
        r["body_off"], r["body_len"] = self._put(bytes(body))
        r["attachment_off"], r["attachment_len"] = self._put(bytes(attachment))
        r["checksum_value_off"], r["checksum_value_len"] = self._put(bytes(request_checksum))
        flags = 0
        if stream is not None:
            flags |= 1 | (2 if stream["need_feedback"] else 0) | (4 if stream["writable"] else 0)
            r["stream_id"] = stream["stream_id"]
            off, _ = self._put(b"".join(struct.pack("<q", x) for x in stream["extra"]), align=8)
            r["extra_streams_off"] = off; r["n_extra_streams"] = len(stream["extra"])
        uf = b"".join(struct.pack("<II", len(k), len(v)) + bytes(k) + bytes(v) for k, v in user_fields)
        r["user_fields_off"], _ = self._put(uf); r["n_user_fields"] = len(user_fields)
        r["flags"] = flags
        self.recs.append(r)

    def arrays(self):
        return np.frombuffer(bytes(self.buf) + b"\0" * 16, np.uint8), np.array(self.recs, dtype=REPLY_DT)
