"""CPU tests: the oracle's ENCODERS against tests/golden/encode_vectors.json — complete request / response /
stream frames built by python-protobuf (an independent protobuf implementation) plus the reference's own
crc32c.cc and vendored snappy (tests/golden/gen_encode_vectors.py).  This pins
  orc_pack_echo_request  <- PackRpcRequest + SerializeRpcRequest   (baidu_rpc_protocol.cpp:1015-1133)
  send_rpc_response      <- SendRpcResponse + SerializeRpcHeaderAndMeta (:273-460, :83-103), through orc_process_batch
  orc_pack_stream_frame  <- PackStreamMessage (policy/streaming_rpc_protocol.cpp:42-58)
byte for byte; the decoders were already pinned by test_oracle_golden.py."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    with open(os.path.join(HERE, "golden", "encode_vectors.json")) as f:
        return json.load(f)


def server_config(O, sv):
    m = dict(O.ECHO_METHOD); m["echo_attachment"] = sv["echo_attachment"]
    m["response_checksum_type"] = sv["response_checksum"]; m["response_compress_type"] = sv["response_compress"]
    return O.make_config([m], server_identity=sv["identity"].encode() if sv["identity"] else None)


def test_pack_echo_request_matches_protobuf(oracle):
    n = 0
    for v in load()["rpc"]:
        got = oracle.pack_echo_request(service=v["service"].encode(), method=v["method"].encode(), log_id=v["log_id"],
                                       correlation_id=v["correlation_id"], compress_type=v["compress"], checksum_type=v["checksum"],
                                       message=bytes.fromhex(v["message_hex"]), attachment=bytes.fromhex(v["attachment_hex"]),
                                       trace=v["trace"], request_id=v["request_id"].encode() if v["request_id"] else None,
                                       timeout_ms=v["timeout_ms"])
        assert got.hex() == v["request_hex"], "request frame %d differs" % n
        n += 1
    assert n > 300


def test_send_rpc_response_matches_protobuf(oracle):
    n_err = 0
    for i, v in enumerate(load()["rpc"]):
        wire = bytes.fromhex(v.get("wire_request_hex", v["request_hex"]))
        data = np.frombuffer(wire + bytes(64), np.uint8)
        runs = np.zeros(1, oracle.RUN_DT); runs[0] = (1, 0, len(wire), -1, 0)
        rs, msgs, resp = oracle.process_batch(server_config(oracle, v["server"]), data, runs)
        assert len(msgs) == 1 and int(rs["consumed"][0]) == len(wire)
        got = bytes(resp[int(msgs["resp_off"][0]):int(msgs["resp_off"][0]) + int(msgs["resp_len"][0])])
        assert int(msgs["error_code"][0]) == v["error_code"], i
        assert got.hex() == v["response_hex"], "response frame %d differs (error_code %d)" % (i, v["error_code"])
        n_err += v["error_code"] != 0
    assert n_err > 40


def test_pack_stream_frame_matches_protobuf(oracle):
    for v in load()["stream"]:
        got = oracle.pack_stream_frame(v["stream_id"], -1 if v["source_stream_id"] is None else v["source_stream_id"], v["frame_type"],
                                       v["has_continuation"], bytes.fromhex(v["data_hex"]))
        assert got.hex() == v["frame_hex"]
