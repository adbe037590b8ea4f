"""CPU: the rpc_dump replay source of the oracle (SampleIterator::Pop + PackRpcRequest's replay branch) against python-protobuf
golden records / request frames (tests/golden/gen_dump_vectors.py)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    with open(os.path.join(HERE, "golden", "dump_vectors.json")) as f:
        return json.load(f)["files"]


def check(rs, msgs, resp, f):
    assert int(rs["consumed"][0]) == f["consumed"] and int(rs["parse_error"][0]) == f.get("error", 2)
    assert len(msgs) == len(f["records"])
    for k, r in enumerate(f["records"]):
        assert int(msgs["status"][k]) == r["status"] and int(msgs["protocol"][k]) == r["protocol"], k
        assert int(msgs["correlation_id"][k]) == f["base_cid"] + k
        if r["frame_hex"] is not None:
            got = bytes(resp[int(msgs["resp_off"][k]):int(msgs["resp_off"][k]) + int(msgs["resp_len"][k])])
            assert got.hex() == r["frame_hex"], "replayed request frame %d differs" % k
        elif r["status"] != 10:
            assert int(msgs["resp_len"][k]) == 0


def test_oracle_replays_dump_files_like_rpc_replay(oracle):
    for f in load():
        data = np.frombuffer(bytes.fromhex(f["file_hex"]) + bytes(64), np.uint8)
        runs = np.zeros(1, oracle.RUN_DT); runs[0] = (f["base_cid"], 0, len(data) - 64, -1, 4)        # B2_RUN_RPC_DUMP
        rs, msgs, resp = oracle.process_batch(oracle.make_config(), data, runs)
        check(rs, msgs, resp, f)
