"""GPU: b2_pack_responses (SendRpcResponse for replies the host produced) against the python-protobuf golden frames and, on random
batches incl. several user fields, big bodies and snappy + crc32c, against the oracle."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import _oracle as O  # noqa: E402
from _replies import ReplyBatch  # noqa: E402
from _traffic import SEED, rnd62  # noqa: E402
from test_oracle_replies import load_vectors  # noqa: E402


def test_device_reply_frames_equal_golden_and_oracle():
    import brpc_b200
    ctx = brpc_b200.Context(device=0, max_batch_bytes=64 << 20, max_msgs=1 << 16, max_runs=16, max_resp_bytes=128 << 20)
    vec, (data, recs) = load_vectors()
    got = ctx.pack_responses(data, recs)
    for k, (v, g) in enumerate(zip(vec, got)):
        assert g.hex() == v["frame"], (k, v["error_code"], v["compress_type"], v["checksum_type"])
    rng = random.Random(SEED + 91)
    b = ReplyBatch()
    for i in range(1500):
        err = rng.choice([0, 0, 0, -1, 1003, 2002])
        n = rng.choice([0, 1, 16, 100, 1000, 5000, 70000, 300000]) if i % 50 == 0 else rng.choice([0, 1, 16, 100, 1000, 5000])
        body = rng.choice([rnd62(rng, n), b"r" * n, (rnd62(rng, 50) * (n // 50 + 1))[:n]])
        stream = None
        if rng.random() < 0.3:
            stream = dict(stream_id=rng.getrandbits(62), need_feedback=rng.random() < 0.5, writable=rng.random() < 0.5,
                          extra=[rng.getrandbits(rng.choice([3, 30, 62])) for _ in range(rng.choice([0, 1, 5]))])
        b.add(error_code=err, error_text=rnd62(rng, rng.choice([0, 3, 90])) if err else b"", body=body, attachment=rnd62(rng, rng.choice([0, 0, 33, 4000])),
              compress_type=rng.choice([0, 0, 1, 1, 2]), checksum_type=rng.choice([0, 1]), content_type=rng.choice([0, 0, 1]),
              correlation_id=rng.getrandbits(63) * rng.choice([1, -1]), request_checksum=rnd62(rng, rng.choice([0, 4])), stream=stream,
              user_fields=[(rnd62(rng, rng.choice([1, 9])), rnd62(rng, rng.choice([0, 40, 300]))) for _ in range(rng.choice([0, 0, 1, 3]))])
    data, recs = b.arrays()
    got = ctx.pack_responses(data, recs)
    empty = 0
    for k, (r, g) in enumerate(zip(recs, got)):
        w = O.pack_response(r, data)
        assert g == w, (k, int(r["error_code"]), int(r["compress_type"]), int(r["checksum_type"]), len(g), len(w))
        empty += len(w) == 0
    assert 50 < empty < 400                                      # gzip replies of OK calls are the only ones not packed
