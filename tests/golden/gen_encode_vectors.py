#!/usr/bin/env python
"""Generate tests/golden/encode_vectors.json: ENCODE-side golden vectors that pin the oracle's
encoders (orc_pack_echo_request, send_rpc_response inside orc_process_batch, orc_pack_stream_frame)
to an independent protobuf implementation (python-protobuf / upb) and to the reference's own
crc32c.cc + vendored snappy (oracle/_ref).  Run in the authoring container; the JSON is committed.

Every vector holds complete frame bytes built HERE, not by the oracle:
  request frame  = "PRPC" BE32(meta+body+att) BE32(meta) | RpcMeta.SerializeToString() | body | att
      with RpcMeta carrying exactly the fields PackRpcRequest sets
      (src/brpc/policy/baidu_rpc_protocol.cpp:1045-1133): request{service_name, method_name, [log_id],
      [trace_id, span_id, parent_span_id], [request_id], [timeout_ms]}, compress_type, checksum_type,
      checksum_value (always, possibly empty), correlation_id, [attachment_size], content_type;
      body = SerializeRpcRequest (:1015-1043): EchoRequest bytes, snappy-compressed by the vendored
      snappy when compress_type = SNAPPY, CRC32C over the (compressed) body.
  response frame = what SendRpcResponse (:273-460) writes for that request under a given server
      configuration: RpcMeta{response{error_code, [error_text]}, correlation_id, compress_type,
      content_type, checksum_type, checksum_value, [attachment_size]} (:339-352), body, attachment;
      error texts follow Controller::SetFailed (src/brpc/controller.cpp:468-495) and the format
      strings at baidu_rpc_protocol.cpp:700-707, :741-757, :819-829.
  stream frame   = "STRM" ... StreamFrameMeta.SerializeToString() (policy/streaming_rpc_protocol.cpp:42-58).
"""
import ctypes
import json
import os
import random
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as G  # noqa: E402

ROOT = G.ROOT
ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_leaf.so"))
ref.ref_crc32c_extend.restype = ctypes.c_uint32
ref.ref_crc32c_extend.argtypes = [ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t]
ref.ref_crc32c_mask.restype = ctypes.c_uint32
ref.ref_crc32c_mask.argtypes = [ctypes.c_uint32]
ref.ref_snappy_max_compressed_length.restype = ctypes.c_size_t
ref.ref_snappy_max_compressed_length.argtypes = [ctypes.c_size_t]
ref.ref_snappy_compress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)]


def snappy(raw):
    cap = ref.ref_snappy_max_compressed_length(len(raw))
    out = ctypes.create_string_buffer(cap)
    n = ctypes.c_size_t(cap)
    ref.ref_snappy_compress(raw, len(raw), out, ctypes.byref(n))
    return out.raw[:n.value]


def masked_crc_be(b):
    return struct.pack(">I", ref.ref_crc32c_mask(ref.ref_crc32c_extend(0, b, len(b))))


def frame(magic, meta, payload):
    return magic + struct.pack(">II", len(meta) + len(payload), len(meta)) + meta + payload


def echo_body(message, compress, checksum, cls):
    m = cls(); m.message = message
    body = m.SerializeToString()
    if compress == 1:
        body = snappy(body)
    cks = masked_crc_be(body) if checksum == 1 else b""
    return body, cks


ASCII = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 _-"


def rand_text(rng, n):
    return "".join(rng.choice(ASCII) for _ in range(n))


def gen_rpc(rng, n):
    out = []
    for i in range(n):
        kind = rng.random()
        service, method = "example.EchoService", "Echo"
        if kind < 0.08: service = "EchoService"                      # jprotobuf short name
        elif kind < 0.14: service = "nosuch" + rand_text(rng, rng.randrange(0, 8)).replace(" ", "")      # no '.': ENOSERVICE
        elif kind < 0.20: service = "example.Other" + str(rng.randrange(100))                            # ENOMETHOD
        elif kind < 0.26: method = "Echo" + rand_text(rng, rng.randrange(1, 6)).replace(" ", "")         # ENOMETHOD
        big = i % 40 == 7                                                # a few large bodies (varint widths, snappy fragments)
        mlen = rng.choice([1024, 4096, 16383, 16384, 70000]) if big else rng.choice([0, 1, 5, 11, 16, 63, 64, 127, 128, 129, 300, rng.randrange(0, 600)])
        message = rand_text(rng, mlen) if rng.random() < 0.7 and not big else "r" * mlen
        if big and rng.random() < 0.5: message = (rand_text(rng, 97) * (mlen // 97 + 1))[:mlen]
        att = bytes(rng.randrange(256) for _ in range(rng.choice([0, 0, 0, 1, 17, 200, 900])))
        compress = 1 if rng.random() < 0.25 else 0
        checksum = 1 if rng.random() < 0.35 else 0
        cid = G.rand_i64(rng)
        log_id = G.rand_i64(rng) if rng.random() < 0.7 else None
        trace = (G.rand_i64(rng), G.rand_i64(rng), G.rand_i64(rng)) if rng.random() < 0.15 else None
        request_id = rand_text(rng, rng.randrange(1, 20)).replace(" ", "") if rng.random() < 0.15 else None
        timeout_ms = rng.choice([1, 500, 30000, 2147483647]) if rng.random() < 0.2 else 0
        body, cks = echo_body(message, compress, checksum, G.EchoRequest)
        m = G.RpcMeta()
        m.request.service_name = service; m.request.method_name = method
        if log_id is not None: m.request.log_id = log_id
        if trace: m.request.trace_id, m.request.span_id, m.request.parent_span_id = trace
        if request_id: m.request.request_id = request_id
        if timeout_ms > 0: m.request.timeout_ms = timeout_ms
        m.compress_type = compress; m.checksum_type = checksum; m.checksum_value = cks
        m.correlation_id = cid
        if att: m.attachment_size = len(att)
        m.content_type = 0
        req = frame(b"PRPC", m.SerializeToString(), body + att)
        # a corrupted body (CRC mismatch / bad snappy / bad pb) exercises the EREQUEST text
        corrupt = rng.random() < 0.08 and len(body) > 3 and (checksum or compress)
        if corrupt:
            bb = bytearray(body); bb[rng.randrange(len(bb))] ^= 0x5a
            wire_req = frame(b"PRPC", m.SerializeToString(), bytes(bb) + att)
        else:
            wire_req = req
        # server configuration
        echo_att = rng.random() < 0.7
        r_cks = 1 if rng.random() < 0.3 else 0
        r_cmp = 1 if rng.random() < 0.2 else 0
        identity = rng.choice([None, None, "10.0.0.1:8000", "192.168.100.200:65535"])
        err, text = 0, ""
        if service.startswith("nosuch"):
            err, text = 1001, "Fail to find service=%s" % service
        elif service.startswith("example.Other") or method != "Echo":
            err, text = 1002, "Fail to find method=%s/%s" % (service, method)
        elif corrupt:
            # a flipped byte always breaks the CRC; with snappy-only it may still decode to a valid pb: skip those
            if not checksum:
                continue
            err = 1003
            text = ("Fail to parse request=example.EchoRequest, ContentType=pb, CompressType=%s, ChecksumType=%s, request_size=%d"
                    % ("snappy" if compress else "none", "crc32c" if checksum else "none", len(body) + len(att)))
        r = G.RpcMeta()
        r.response.error_code = err
        if err:
            r.response.error_text = ("[%s]" % identity if identity else "") + "[E%d]" % err + text
            r.compress_type = 0; r.checksum_type = 0; r.checksum_value = cks      # request's checksum_value travels back (:608 + :349)
            r.correlation_id = cid; r.content_type = 0
            resp = frame(b"PRPC", r.SerializeToString(), b"")
        else:
            rbody, rcks = echo_body(message, r_cmp, r_cks, G.EchoResponse)
            r.compress_type = r_cmp; r.checksum_type = r_cks
            r.checksum_value = rcks if r_cks else cks
            r.correlation_id = cid; r.content_type = 0
            ratt = att if echo_att else b""
            if ratt: r.attachment_size = len(ratt)
            resp = frame(b"PRPC", r.SerializeToString(), rbody + ratt)
        rec = {"service": service, "method": method, "message_hex": message.encode().hex(), "attachment_hex": att.hex(),
               "compress": compress, "checksum": checksum, "correlation_id": cid, "log_id": log_id,
               "trace": trace, "request_id": request_id, "timeout_ms": timeout_ms,
               "request_hex": req.hex(), "corrupt": bool(corrupt),
               "server": {"echo_attachment": int(echo_att), "response_checksum": r_cks, "response_compress": r_cmp, "identity": identity},
               "error_code": err, "response_hex": resp.hex()}
        if corrupt:
            rec["wire_request_hex"] = wire_req.hex()
        out.append(rec)
    return out


def gen_stream(rng, n):
    out = []
    for i in range(n):
        m = G.StreamFrameMeta()
        sid = rng.randrange(0, 1 << 62); m.stream_id = sid
        src = rng.randrange(0, 1 << 62) if rng.random() < 0.8 else None
        if src is not None: m.source_stream_id = src
        ft = rng.randrange(0, 5); m.frame_type = ft
        hc = rng.choice([None, True, False])
        if hc is not None: m.has_continuation = hc
        data = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 100, 5000])))
        out.append({"stream_id": sid, "source_stream_id": src, "frame_type": ft, "has_continuation": hc,
                    "data_hex": data.hex(), "frame_hex": frame(b"STRM", m.SerializeToString(), data).hex()})
    return out


def main():
    rng = random.Random(G.SEED + 2)
    rpc = gen_rpc(rng, 420)
    stream = gen_stream(rng, 80)
    with open(os.path.join(HERE, "encode_vectors.json"), "w") as f:
        json.dump({"seed": G.SEED + 2, "rpc": rpc, "stream": stream}, f, indent=0)
    print("rpc %d (errors %d, corrupt %d)  stream %d" % (len(rpc), sum(1 for r in rpc if r["error_code"]), sum(1 for r in rpc if r["corrupt"]), len(stream)))


if __name__ == "__main__":
    sys.exit(main())
