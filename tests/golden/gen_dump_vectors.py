#!/usr/bin/env python
"""tests/golden/dump_vectors.json: rpc_dump records (RpcDumpContext::Serialize, src/brpc/rpc_dump.cpp:237-258) built with
python-protobuf from a descriptor transcribed from src/brpc/rpc_dump.proto:23-48, and — for the baidu_std samples — the request
frame rpc_replay sends for them: PackRpcRequest's replay branch (baidu_rpc_protocol.cpp:1067-1075,1080,1106-1120) serialised by
python-protobuf's RpcMeta (gen_golden.py).  Pins oracle/b2_oracle.c::process_dump_record and, through it, the device."""
import json
import os
import random
import struct
import sys

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as G  # noqa: E402

F = descriptor_pb2.FieldDescriptorProto


def build():
    fd = descriptor_pb2.FileDescriptorProto(); fd.name = "b2_dump.proto"; fd.package = "brpc.dump"; fd.syntax = "proto2"
    ct = fd.enum_type.add(); ct.name = "CompressType"
    for i, n in enumerate(["COMPRESS_TYPE_NONE", "COMPRESS_TYPE_SNAPPY", "COMPRESS_TYPE_GZIP", "COMPRESS_TYPE_ZLIB", "COMPRESS_TYPE_LZ4"]):
        v = ct.value.add(); v.name, v.number = n, i
    pt = fd.enum_type.add(); pt.name = "ProtocolType"
    for i in range(28):
        v = pt.value.add(); v.name, v.number = "PROTOCOL_%d" % i, i
    m = fd.message_type.add(); m.name = "RpcDumpMeta"
    G._field(m, "service_name", 1, F.TYPE_STRING); G._field(m, "method_name", 2, F.TYPE_STRING); G._field(m, "method_index", 3, F.TYPE_INT32)
    G._field(m, "compress_type", 4, F.TYPE_ENUM, type_name=".brpc.dump.CompressType"); G._field(m, "protocol_type", 5, F.TYPE_ENUM, type_name=".brpc.dump.ProtocolType")
    G._field(m, "attachment_size", 6, F.TYPE_INT32); G._field(m, "authentication_data", 7, F.TYPE_BYTES); G._field(m, "user_data", 8, F.TYPE_BYTES); G._field(m, "nshead", 9, F.TYPE_BYTES)
    pool = descriptor_pool.DescriptorPool(); pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("brpc.dump.RpcDumpMeta"))


RpcDumpMeta = build()


def main():
    rng = random.Random(G.SEED + 9)
    files = []
    for fi in range(12):
        base_cid = rng.randrange(1, 1 << 50)
        records = []; blob = b""
        for k in range(rng.randrange(1, 40)):
            m = RpcDumpMeta()
            proto = rng.choice([1, 1, 1, 1, 3, 7, None])
            if rng.random() < 0.95: m.service_name = rng.choice(["example.EchoService", "EchoService", "", G.rand_name(rng, 30)])
            if rng.random() < 0.95: m.method_name = rng.choice(["Echo", "", G.rand_name(rng, 9)])
            if rng.random() < 0.2: m.method_index = rng.randrange(0, 9)
            if rng.random() < 0.6: m.compress_type = rng.randrange(0, 5)
            if proto is not None: m.protocol_type = proto
            body = bytes(rng.randrange(256) for _ in range(rng.choice([0, 3, 100, 1027, 5000])))
            att = 0
            if rng.random() < 0.3 and len(body) > 4:
                att = rng.randrange(1, len(body)); m.attachment_size = att
            if rng.random() < 0.1: m.authentication_data = b"cred"
            if rng.random() < 0.1: m.user_data = b"\x01\x02"
            meta = m.SerializeToString()
            rec = b"PRPC" + struct.pack(">II", len(meta) + len(body), len(meta)) + meta + body
            cid = base_cid + k
            frame = None
            if proto == 1:
                r = G.RpcMeta()
                r.request.service_name = m.service_name; r.request.method_name = m.method_name
                r.compress_type = m.compress_type if m.HasField("compress_type") else 0
                r.correlation_id = cid
                if att: r.attachment_size = att
                r.content_type = 0
                rm = r.SerializeToString()
                frame = (b"PRPC" + struct.pack(">II", len(rm) + len(body), len(rm)) + rm + body).hex()
            records.append({"status": 10 if proto == 1 else 6, "protocol": proto or 0, "frame_hex": frame})
            blob += rec
        tail = rng.choice([b"", b"PRPC\x00\x00", b"PRPC" + struct.pack(">II", 50, 10) + bytes(20)])      # a record still being written
        files.append({"base_cid": base_cid, "file_hex": (blob + tail).hex(), "consumed": len(blob), "records": records})
    # format errors end the file (SampleIterator::Pop): wrong magic, meta_size > body_size
    good = bytes.fromhex(files[0]["file_hex"])[:files[0]["consumed"]]
    files.append({"base_cid": 5, "file_hex": (good + b"XRPC" + bytes(40)).hex(), "consumed": len(good), "records": files[0]["records"], "error": 5})
    files.append({"base_cid": 5, "file_hex": (good + b"PRPC" + struct.pack(">II", 4, 9) + bytes(40)).hex(), "consumed": len(good), "records": files[0]["records"], "error": 5})
    # the cid of the error files starts at 5: rebuild their expected frames by patching? simpler: mark them as cut-only checks
    for f in files[-2:]:
        f["records"] = [{"status": r["status"], "protocol": r["protocol"], "frame_hex": None} for r in f["records"]]
    with open(os.path.join(HERE, "dump_vectors.json"), "w") as fo:
        json.dump({"seed": G.SEED + 9, "files": files}, fo)
    print("files", len(files), "records", sum(len(f["records"]) for f in files))


if __name__ == "__main__":
    main()
