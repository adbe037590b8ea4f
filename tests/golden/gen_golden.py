#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ (run in the AUTHORING container,
where /root/reference and python-protobuf exist; the fixtures are committed and
travel to the GPU box, this script's inputs do not).

What pins what:
  * rpc_meta_vectors.json   - RpcMeta / StreamFrameMeta / EchoRequest wire bytes
    produced and parsed by python-protobuf (upb) from descriptors transcribed
    from the reference's .proto files:
      src/brpc/policy/baidu_rpc_meta.proto:26-55, src/brpc/streaming_rpc_meta.proto:24-53,
      src/brpc/options.proto:69-93, example/echo_c++/echo.proto:23-29.
    There is no protoc here, so the descriptors are built with descriptor_pb2.
  * pb_fuzz_vectors.json    - mutated metas with upb's accept/reject verdict and,
    when accepted, the decoded fields.
  * crc32c_kat.json         - RFC 3720 B.4 known answers as asserted by the
    reference's test/crc32c_unittest.cc:18-71, re-derived through the reference's
    own crc32c.cc (oracle/_ref) so the numbers are the reference's, not ours.
  * snappy_vectors.json     - compressed bytes produced by the reference's vendored
    snappy (oracle/_ref) for the generator patterns of
    test/brpc_snappy_compress_unittest.cpp:80-255.
"""
import ctypes
import json
import os
import random
import sys

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SEED = 20260921

F = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=F.LABEL_OPTIONAL, type_name=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name
    return f


def build_pool():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "b2_golden.proto"
    fd.package = "brpc.policy"
    fd.syntax = "proto2"

    ct = fd.enum_type.add()
    ct.name = "ContentType"
    for i, n in enumerate(["CONTENT_TYPE_PB", "CONTENT_TYPE_JSON", "CONTENT_TYPE_PROTO_JSON", "CONTENT_TYPE_PROTO_TEXT"]):
        v = ct.value.add(); v.name, v.number = n, i
    ft = fd.enum_type.add()
    ft.name = "FrameType"
    for i, n in enumerate(["FRAME_TYPE_UNKNOWN", "FRAME_TYPE_RST", "FRAME_TYPE_CLOSE", "FRAME_TYPE_DATA", "FRAME_TYPE_FEEDBACK"]):
        v = ft.value.add(); v.name, v.number = n, i

    m = fd.message_type.add(); m.name = "ChunkInfo"
    _field(m, "stream_id", 1, F.TYPE_INT64, F.LABEL_REQUIRED)
    _field(m, "chunk_id", 2, F.TYPE_INT64, F.LABEL_REQUIRED)

    m = fd.message_type.add(); m.name = "StreamSettings"
    _field(m, "stream_id", 1, F.TYPE_INT64, F.LABEL_REQUIRED)
    _field(m, "need_feedback", 2, F.TYPE_BOOL)
    _field(m, "writable", 3, F.TYPE_BOOL)
    _field(m, "extra_stream_ids", 4, F.TYPE_INT64, F.LABEL_REPEATED)

    m = fd.message_type.add(); m.name = "Feedback"
    _field(m, "consumed_size", 1, F.TYPE_INT64)

    m = fd.message_type.add(); m.name = "StreamFrameMeta"
    _field(m, "stream_id", 1, F.TYPE_INT64, F.LABEL_REQUIRED)
    _field(m, "source_stream_id", 2, F.TYPE_INT64)
    _field(m, "frame_type", 3, F.TYPE_ENUM, type_name=".brpc.policy.FrameType")
    _field(m, "has_continuation", 4, F.TYPE_BOOL)
    _field(m, "feedback", 5, F.TYPE_MESSAGE, type_name=".brpc.policy.Feedback")

    m = fd.message_type.add(); m.name = "RpcRequestMeta"
    _field(m, "service_name", 1, F.TYPE_STRING, F.LABEL_REQUIRED)
    _field(m, "method_name", 2, F.TYPE_STRING, F.LABEL_REQUIRED)
    _field(m, "log_id", 3, F.TYPE_INT64)
    _field(m, "trace_id", 4, F.TYPE_INT64)
    _field(m, "span_id", 5, F.TYPE_INT64)
    _field(m, "parent_span_id", 6, F.TYPE_INT64)
    _field(m, "request_id", 7, F.TYPE_STRING)
    _field(m, "timeout_ms", 8, F.TYPE_INT32)

    m = fd.message_type.add(); m.name = "RpcResponseMeta"
    _field(m, "error_code", 1, F.TYPE_INT32)
    _field(m, "error_text", 2, F.TYPE_STRING)

    m = fd.message_type.add(); m.name = "RpcMeta"
    _field(m, "request", 1, F.TYPE_MESSAGE, type_name=".brpc.policy.RpcRequestMeta")
    _field(m, "response", 2, F.TYPE_MESSAGE, type_name=".brpc.policy.RpcResponseMeta")
    _field(m, "compress_type", 3, F.TYPE_INT32)
    _field(m, "correlation_id", 4, F.TYPE_INT64)
    _field(m, "attachment_size", 5, F.TYPE_INT32)
    _field(m, "chunk_info", 6, F.TYPE_MESSAGE, type_name=".brpc.policy.ChunkInfo")
    _field(m, "authentication_data", 7, F.TYPE_BYTES)
    _field(m, "stream_settings", 8, F.TYPE_MESSAGE, type_name=".brpc.policy.StreamSettings")
    e = m.nested_type.add(); e.name = "UserFieldsEntry"; e.options.map_entry = True
    _field(e, "key", 1, F.TYPE_STRING)
    _field(e, "value", 2, F.TYPE_STRING)
    _field(m, "user_fields", 9, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".brpc.policy.RpcMeta.UserFieldsEntry")
    _field(m, "content_type", 10, F.TYPE_ENUM, type_name=".brpc.policy.ContentType")
    _field(m, "checksum_type", 11, F.TYPE_INT32)
    _field(m, "checksum_value", 12, F.TYPE_BYTES)

    m = fd.message_type.add(); m.name = "EchoRequest"
    _field(m, "message", 1, F.TYPE_STRING, F.LABEL_REQUIRED)
    m = fd.message_type.add(); m.name = "EchoResponse"
    _field(m, "message", 1, F.TYPE_STRING, F.LABEL_REQUIRED)

    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return pool


POOL = build_pool()


def cls(name):
    return message_factory.GetMessageClass(POOL.FindMessageTypeByName("brpc.policy." + name))


RpcMeta, StreamFrameMeta, EchoRequest, EchoResponse = cls("RpcMeta"), cls("StreamFrameMeta"), cls("EchoRequest"), cls("EchoResponse")


def _bytes_of(s):
    # proto2 strings are not UTF-8 checked on parse; surface them as latin-1 safe hex
    return s.encode("utf-8", "surrogateescape").hex() if isinstance(s, str) else bytes(s).hex()


def meta_fields(m):
    """Decoded view of an RpcMeta, in the vocabulary of oracle/b2_oracle.h."""
    d = {}
    d["has_request"] = m.HasField("request")
    if d["has_request"]:
        r = m.request
        d["service_name"] = _bytes_of(r.service_name)
        d["method_name"] = _bytes_of(r.method_name)
        for k in ("log_id", "trace_id", "span_id", "parent_span_id", "timeout_ms"):
            d["has_" + k] = r.HasField(k)
            d[k] = getattr(r, k)
        d["has_request_id"] = r.HasField("request_id")
        d["request_id"] = _bytes_of(r.request_id)
    d["has_response"] = m.HasField("response")
    if d["has_response"]:
        d["has_error_code"] = m.response.HasField("error_code")
        d["error_code"] = m.response.error_code
        d["has_error_text"] = m.response.HasField("error_text")
        d["error_text"] = _bytes_of(m.response.error_text)
    for k in ("compress_type", "correlation_id", "attachment_size", "content_type", "checksum_type"):
        d["has_" + k] = m.HasField(k)
        d[k] = getattr(m, k)
    d["has_checksum_value"] = m.HasField("checksum_value")
    d["checksum_value"] = bytes(m.checksum_value).hex()
    d["has_authentication_data"] = m.HasField("authentication_data")
    d["authentication_data"] = bytes(m.authentication_data).hex()
    d["has_chunk_info"] = m.HasField("chunk_info")
    d["has_stream_settings"] = m.HasField("stream_settings")
    if d["has_stream_settings"]:
        d["ss_stream_id"] = m.stream_settings.stream_id
        d["ss_need_feedback"] = m.stream_settings.need_feedback
        d["ss_writable"] = m.stream_settings.writable
        d["ss_n_extra"] = len(m.stream_settings.extra_stream_ids)
    d["n_user_fields_distinct"] = len(m.user_fields)
    return d


def stream_fields(m):
    d = {}
    for k in ("stream_id", "source_stream_id", "frame_type", "has_continuation"):
        d["has_" + k] = m.HasField(k)
        d[k] = int(getattr(m, k))
    d["has_feedback"] = m.HasField("feedback")
    d["feedback_has_consumed_size"] = m.feedback.HasField("consumed_size") if d["has_feedback"] else False
    d["consumed_size"] = m.feedback.consumed_size if d["has_feedback"] else 0
    return d


def rand_i64(rng):
    c = rng.random()
    if c < 0.3:
        return rng.randrange(0, 128)
    if c < 0.6:
        return rng.randrange(0, 1 << 31)
    if c < 0.8:
        return rng.randrange(-(1 << 63), 1 << 63)
    return -rng.randrange(1, 1000)


def rand_i32(rng):
    c = rng.random()
    if c < 0.5:
        return rng.randrange(0, 4)
    if c < 0.8:
        return rng.randrange(0, 1 << 31)
    return -rng.randrange(1, 1 << 31)


def rand_name(rng, n=None):
    n = rng.randrange(0, 40) if n is None else n
    return "".join(rng.choice("abcdefghijklmnopqrstuvwxyzABCXYZ0123456789._") for _ in range(n))


def gen_valid_metas(rng):
    out = []

    def add(m, note):
        b = m.SerializeToString()
        p = RpcMeta(); p.ParseFromString(b)
        out.append({"note": note, "hex": b.hex(), "fields": meta_fields(p)})

    # the BASELINE echo request meta (SURVEY §8 preamble): 46 bytes
    m = RpcMeta()
    m.request.service_name = "example.EchoService"; m.request.method_name = "Echo"; m.request.log_id = 12345
    m.compress_type = 0; m.correlation_id = (7 << 32) | 123456
    m.content_type = 0; m.checksum_type = 0; m.checksum_value = b""
    add(m, "baseline echo request meta (46 B)")
    # the BASELINE echo response meta: 18 bytes
    m = RpcMeta()
    m.response.error_code = 0; m.compress_type = 0; m.correlation_id = (7 << 32) | 123456
    m.content_type = 0; m.checksum_type = 0; m.checksum_value = b""
    add(m, "baseline echo response meta (18 B)")
    for i in range(300):
        m = RpcMeta()
        if rng.random() < 0.85:
            m.request.service_name = rng.choice(["example.EchoService", "EchoService", rand_name(rng)])
            m.request.method_name = rng.choice(["Echo", rand_name(rng, rng.randrange(0, 12))])
            if rng.random() < 0.7: m.request.log_id = rand_i64(rng)
            if rng.random() < 0.3:
                m.request.trace_id = rand_i64(rng); m.request.span_id = rand_i64(rng); m.request.parent_span_id = rand_i64(rng)
            if rng.random() < 0.2: m.request.request_id = rand_name(rng)
            if rng.random() < 0.2: m.request.timeout_ms = rand_i32(rng)
        if rng.random() < 0.2:
            m.response.error_code = rand_i32(rng)
            if rng.random() < 0.5: m.response.error_text = rand_name(rng, rng.randrange(0, 300))
        if rng.random() < 0.9: m.compress_type = rand_i32(rng)
        if rng.random() < 0.95: m.correlation_id = rand_i64(rng)
        if rng.random() < 0.3: m.attachment_size = rand_i32(rng)
        if rng.random() < 0.1: m.chunk_info.stream_id = rand_i64(rng); m.chunk_info.chunk_id = rand_i64(rng)
        if rng.random() < 0.15: m.authentication_data = bytes(rng.randrange(256) for _ in range(rng.randrange(0, 40)))
        if rng.random() < 0.15:
            m.stream_settings.stream_id = rand_i64(rng)
            if rng.random() < 0.5: m.stream_settings.need_feedback = rng.random() < 0.5
            if rng.random() < 0.5: m.stream_settings.writable = rng.random() < 0.5
            for _ in range(rng.randrange(0, 4)): m.stream_settings.extra_stream_ids.append(rand_i64(rng))
        if rng.random() < 0.15:
            for _ in range(rng.randrange(1, 4)): m.user_fields[rand_name(rng, 5)] = rand_name(rng, 9)
        if rng.random() < 0.9: m.content_type = rng.randrange(0, 4)
        if rng.random() < 0.9: m.checksum_type = rng.choice([0, 1, 1, 7, -1])
        if rng.random() < 0.9: m.checksum_value = bytes(rng.randrange(256) for _ in range(rng.choice([0, 4, 4, 3, 9])))
        add(m, "random #%d" % i)
    return out


def gen_stream_metas(rng):
    out = []
    for i in range(120):
        m = StreamFrameMeta()
        m.stream_id = rand_i64(rng)
        if rng.random() < 0.7: m.source_stream_id = rand_i64(rng)
        if rng.random() < 0.9: m.frame_type = rng.randrange(0, 5)
        if rng.random() < 0.6: m.has_continuation = rng.random() < 0.5
        if rng.random() < 0.3:
            m.feedback.SetInParent()
            if rng.random() < 0.8: m.feedback.consumed_size = rand_i64(rng)
        b = m.SerializeToString()
        p = StreamFrameMeta(); p.ParseFromString(b)
        out.append({"hex": b.hex(), "fields": stream_fields(p)})
    return out


def mutate(rng, b):
    b = bytearray(b)
    op = rng.randrange(7)
    if op == 0 and b:
        b[rng.randrange(len(b))] = rng.randrange(256)
    elif op == 1 and b:
        del b[rng.randrange(len(b)):]
    elif op == 2:
        pos = rng.randrange(len(b) + 1)
        b[pos:pos] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 6)))
    elif op == 3 and b:
        pos = rng.randrange(len(b)); del b[pos:pos + rng.randrange(1, 4)]
    elif op == 4:
        # append a random well-formed unknown field (incl. groups / fixed widths)
        fn = rng.choice([13, 14, 15, 16, 100, 2047, 300000])
        wt = rng.choice([0, 1, 2, 5, 3])
        def varint(v):
            o = bytearray()
            while v >= 0x80: o.append((v & 0x7f) | 0x80); v >>= 7
            o.append(v); return bytes(o)
        tag = varint((fn << 3) | wt)
        if wt == 0: body = varint(rng.randrange(1 << 64))
        elif wt == 1: body = bytes(rng.randrange(256) for _ in range(8))
        elif wt == 5: body = bytes(rng.randrange(256) for _ in range(4))
        elif wt == 2: n = rng.randrange(0, 20); body = varint(n) + bytes(rng.randrange(256) for _ in range(n))
        else: body = varint((7 << 3) | 0) + varint(5) + varint((fn << 3) | 4)
        pos = rng.choice([0, len(b)])
        b[pos:pos] = tag + body
    elif op == 5 and len(b) > 2:
        # duplicate a slice (repeated field occurrences / merge semantics)
        i = rng.randrange(len(b)); j = rng.randrange(i, len(b))
        b[j:j] = b[i:j]
    else:
        # over-long varint encodings of a small value appended as field 3
        b += bytes([0x18]) + bytes([0x80 | rng.randrange(128) for _ in range(rng.randrange(1, 11))]) + bytes([rng.randrange(2)])
    return bytes(b)


def gen_fuzz(rng, seeds, Cls, fields_fn, n):
    out = []
    for i in range(n):
        b = bytes.fromhex(rng.choice(seeds)["hex"])
        for _ in range(rng.randrange(1, 4)):
            b = mutate(rng, b)
        m = Cls()
        try:
            m.ParseFromString(b)
            # python's ParseFromString skips the required-field check that C++
            # ParseFromCodedStream (MessageLite::ParseFrom<kParse>) performs
            ok = m.IsInitialized()
        except Exception:
            ok = False
        rec = {"hex": b.hex(), "ok": ok}
        if ok:
            try:
                rec["fields"] = fields_fn(m)
            except Exception:
                continue   # e.g. invalid UTF-8 surfaced by the python accessor; skip
        out.append(rec)
    return out


def gen_echo(rng):
    out = []
    for n in [0, 1, 11, 16, 64, 127, 128, 1024, 4096, 16383, 16384, 70000]:
        m = EchoRequest(); m.message = "r" * n
        out.append({"hex": m.SerializeToString().hex(), "ok": True, "message_len": n})
    seeds = [{"hex": o["hex"]} for o in out[:8]]
    for i in range(200):
        b = bytes.fromhex(rng.choice(seeds)["hex"])
        for _ in range(rng.randrange(1, 3)):
            b = mutate(rng, b)
        m = EchoRequest()
        try:
            m.ParseFromString(b); ok = m.IsInitialized()
        except Exception:
            ok = False
        rec = {"hex": b.hex(), "ok": ok}
        if ok:
            try:
                rec["message_hex"] = _bytes_of(m.message)
            except Exception:
                continue
        out.append(rec)
    return out


def gen_ref_leaf():
    ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_leaf.so"))
    ref.ref_crc32c_extend.restype = ctypes.c_uint32
    ref.ref_crc32c_extend.argtypes = [ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t]
    ref.ref_crc32c_mask.restype = ctypes.c_uint32; ref.ref_crc32c_mask.argtypes = [ctypes.c_uint32]
    ref.ref_snappy_max_compressed_length.restype = ctypes.c_size_t
    ref.ref_snappy_max_compressed_length.argtypes = [ctypes.c_size_t]
    ref.ref_snappy_compress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)]

    def crc(b, init=0):
        return ref.ref_crc32c_extend(init, b, len(b))

    kat = []
    # test/crc32c_unittest.cc:18-54 (RFC 3720 B.4)
    buf = bytes(32); kat.append({"name": "32 zeros", "hex": buf.hex(), "crc": crc(buf), "expect": 0x8a9136aa})
    buf = bytes([0xff] * 32); kat.append({"name": "32 0xff", "hex": buf.hex(), "crc": crc(buf), "expect": 0x62a8ab43})
    buf = bytes(range(32)); kat.append({"name": "0..31", "hex": buf.hex(), "crc": crc(buf), "expect": 0x46dd794e})
    buf = bytes(31 - i for i in range(32)); kat.append({"name": "31..0", "hex": buf.hex(), "crc": crc(buf), "expect": 0x113fdb5c})
    data = bytes([0x01, 0xc0, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00,
                  0x14, 0x00, 0x00, 0x00, 0x00, 0x00, 0x04, 0x00, 0x00, 0x00, 0x00, 0x14, 0x00, 0x00, 0x00, 0x18,
                  0x28, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x02, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00])
    kat.append({"name": "iscsi read pdu", "hex": data.hex(), "crc": crc(data), "expect": 0xd9963a56})
    kat.append({"name": "123456789", "hex": b"123456789".hex(), "crc": crc(b"123456789"), "expect": 0xe3069283})
    for k in kat:
        assert k["crc"] == k["expect"], k
    rng = random.Random(SEED)
    for n in [0, 1, 3, 7, 8, 9, 15, 16, 17, 63, 64, 65, 1027, 4099, 65536 + 3]:
        b = bytes(rng.randrange(256) for _ in range(n))
        kat.append({"name": "random %d" % n, "hex": b.hex(), "crc": crc(b), "masked": ref.ref_crc32c_mask(crc(b))})
    # Extend associativity (crc32c_unittest.cc:60-63)
    kat.append({"name": "extend hello+world", "hex": b"hello world".hex(), "crc": crc(b"world", crc(b"hello ")),
                "expect": crc(b"hello world")})

    snap = []
    table62 = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"
    def pattern(n):      # brpc_snappy_compress_unittest.cpp: repeating a-z0-9 text
        t = "abcdefghijklmnopqrstuvwxyz0123456789"
        return (t * (n // len(t) + 1))[:n].encode()
    def rnd62(n, rng):
        return "".join(rng.choice(table62) for _ in range(n)).encode()
    cases = [("empty", b"")]
    for n in [1, 5, 16, 128, 200, 782, 1024, 4096, 12435, 16384, 32768, 65536, 65537, 100000, 262144]:
        cases.append(("pattern %d" % n, pattern(n)))
    for n in [15, 64, 1027, 4099, 70000]:
        cases.append(("random62 %d" % n, rnd62(n, rng)))
    cases.append(("zeros 100000", bytes(100000)))
    mixed = bytearray()
    for i in range(200):
        mixed += rnd62(rng.randrange(1, 200), rng) if rng.random() < 0.5 else bytes(mixed[max(0, len(mixed) - rng.randrange(1, 3000)):][:rng.randrange(1, 400)])
    cases.append(("mixed backrefs", bytes(mixed)))
    for name, raw in cases:
        cap = ref.ref_snappy_max_compressed_length(len(raw))
        out = ctypes.create_string_buffer(cap)
        olen = ctypes.c_size_t(cap)
        ref.ref_snappy_compress(raw, len(raw), out, ctypes.byref(olen))
        comp = out.raw[:olen.value]
        rec = {"name": name, "raw_len": len(raw), "comp_hex": comp.hex()}
        # raw is reproducible from the name for the big deterministic cases; store it when small
        if len(raw) <= 20000 or not name.startswith("pattern"):
            rec["raw_hex"] = raw.hex()
        snap.append(rec)
    return kat, snap


def main():
    rng = random.Random(SEED)
    metas = gen_valid_metas(rng)
    streams = gen_stream_metas(rng)
    with open(os.path.join(HERE, "rpc_meta_vectors.json"), "w") as f:
        json.dump({"seed": SEED, "rpc_meta": metas, "stream_frame_meta": streams, "echo_request": gen_echo(rng)}, f, indent=0)
    fuzz = {"seed": SEED,
            "rpc_meta": gen_fuzz(rng, metas, RpcMeta, meta_fields, 3000),
            "stream_frame_meta": gen_fuzz(rng, streams, StreamFrameMeta, stream_fields, 800)}
    with open(os.path.join(HERE, "pb_fuzz_vectors.json"), "w") as f:
        json.dump(fuzz, f, indent=0)
    kat, snap = gen_ref_leaf()
    with open(os.path.join(HERE, "crc32c_kat.json"), "w") as f:
        json.dump(kat, f, indent=0)
    with open(os.path.join(HERE, "snappy_vectors.json"), "w") as f:
        json.dump(snap, f, indent=0)
    print("rpc_meta %d  stream %d  fuzz %d/%d  crc %d  snappy %d" % (
        len(metas), len(streams), len(fuzz["rpc_meta"]), len(fuzz["stream_frame_meta"]), len(kat), len(snap)))
    print("fuzz accept rate rpc_meta: %.2f" % (sum(r["ok"] for r in fuzz["rpc_meta"]) / len(fuzz["rpc_meta"])))


if __name__ == "__main__":
    sys.exit(main())
