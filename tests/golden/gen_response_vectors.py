#!/usr/bin/env python
"""Generate tests/golden/response_vectors.json: ENCODE-side golden vectors for replies the HOST produced — SendRpcResponse
(src/brpc/policy/baidu_rpc_protocol.cpp:273-460) with every RpcMeta field it can set (:339-380): response{error_code,[error_text]},
compress_type, correlation_id, [attachment_size], [stream_settings{stream_id, need_feedback, writable, extra_stream_ids}]
(Stream::FillSettings, stream.cpp:678-682), [user_fields], content_type, checksum_type, checksum_value — serialized HERE by
python-protobuf (upb) from the reference's .proto files, body compressed by the reference's vendored snappy and checksummed by its
crc32c.cc (oracle/_ref).  They pin oracle/b2_oracle.c:orc_pack_response, the checker of b2_pack_responses.  Run in the authoring
container; the JSON is committed.  (user_fields: a protobuf map has no defined wire order, so the vectors hold at most one entry.)"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as G  # noqa: E402
import gen_encode_vectors as E  # noqa: E402


def main():
    rng = random.Random(20260921)
    out = []
    for i in range(400):
        err = rng.choice([0, 0, 0, 0, -1, 1003, 2001, 2002, 1, 2147483647])
        text = E.rand_text(rng, rng.choice([0, 5, 40, 300])) if err else ""
        body = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 20, 300, 2000]))) if rng.random() < 0.5 else (b"abcdefgh" * rng.choice([1, 40, 300]))
        att = bytes(rng.randrange(256) for _ in range(rng.choice([0, 0, 7, 200])))
        compress = rng.choice([0, 0, 1]); checksum = rng.choice([0, 0, 1]); content = rng.choice([0, 0, 0, 1, 3])
        cid = G.rand_i64(rng)
        req_cks = bytes(rng.randrange(256) for _ in range(rng.choice([0, 0, 4, 9])))
        stream = None
        if rng.random() < 0.35:
            stream = dict(stream_id=G.rand_i64(rng), need_feedback=rng.random() < 0.5, writable=rng.random() < 0.5,
                          extra=[G.rand_i64(rng) for _ in range(rng.choice([0, 0, 1, 3]))])
        user = None
        if rng.random() < 0.3:
            user = [E.rand_text(rng, rng.choice([0, 1, 8, 130])), E.rand_text(rng, rng.choice([0, 3, 200]))]
        eff = 2001 if err == -1 else err
        m = G.RpcMeta()
        m.response.error_code = eff
        if text: m.response.error_text = text
        m.correlation_id = cid; m.compress_type = compress; m.content_type = content; m.checksum_type = checksum
        wire_body = b""; cks = req_cks
        if eff == 0:
            wire_body = E.snappy(body) if compress == 1 else body
            if checksum == 1: cks = E.masked_crc_be(wire_body)
            if att: m.attachment_size = len(att)
        m.checksum_value = cks
        if stream:
            m.stream_settings.stream_id = stream["stream_id"]; m.stream_settings.need_feedback = stream["need_feedback"]; m.stream_settings.writable = stream["writable"]
            for x in stream["extra"]: m.stream_settings.extra_stream_ids.append(x)
        if user: m.user_fields[user[0]] = user[1]
        frame = E.frame(b"PRPC", m.SerializeToString(), (wire_body + att) if eff == 0 else b"")
        out.append(dict(error_code=err, error_text=text, body=body.hex(), attachment=att.hex(), compress_type=compress, checksum_type=checksum,
                        content_type=content, correlation_id=cid, request_checksum=req_cks.hex(), stream=stream, user_field=user, frame=frame.hex()))
    path = os.path.join(HERE, "response_vectors.json")
    json.dump(out, open(path, "w"), separators=(",", ":"))
    print("wrote", path, len(out), os.path.getsize(path))


if __name__ == "__main__":
    main()
