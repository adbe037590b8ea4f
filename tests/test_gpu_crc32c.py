"""GPU: the warp-level CRC-32C primitive (b2_crc32c_batch) vs the RFC 3720 known answers the
reference asserts (test/crc32c_unittest.cc:18-71) and vs the oracle at every length 0..300 x
every 16 alignments, plus large slices."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_crc32c_kat_and_sweep(oracle):
    import brpc_b200
    ctx = brpc_b200.Context(device=0, max_batch_bytes=16 << 20, max_msgs=1 << 16, max_runs=16)
    kat = json.load(open(os.path.join(HERE, "golden", "crc32c_kat.json")))
    blobs = [bytes.fromhex(k["hex"]) for k in kat if not k["name"].startswith("extend")]
    want = [k["crc"] for k in kat if not k["name"].startswith("extend")]
    buf = bytearray(); offs = []; lens = []
    for b in blobs:
        pad = random.Random(len(buf)).randrange(0, 7)
        buf += b"\xee" * pad; offs.append(len(buf)); lens.append(len(b)); buf += b
    got = ctx.crc32c_batch(np.frombuffer(bytes(buf), np.uint8), offs, lens)
    assert got.tolist() == want

    rng = np.random.default_rng(20260921)
    data = rng.integers(0, 256, size=1 << 20, dtype=np.uint8)
    offs, lens = [], []
    for n in range(0, 301):
        for a in range(16):
            offs.append(1000 + 17 * n + a); lens.append(n)
    for n in [1023, 1024, 1027, 4096, 4099, 65536, 65539, 300000, 1 << 19]:
        for a in [0, 1, 2, 3, 5, 13]:
            offs.append(a + 7); lens.append(n)
    got = ctx.crc32c_batch(data, offs, lens)
    raw = data.tobytes()
    exp = [oracle.crc32c(raw[o:o + n]) for o, n in zip(offs, lens)]
    bad = [(o, n, g, e) for o, n, g, e in zip(offs, lens, got.tolist(), exp) if g != e]
    assert not bad, bad[:5]
