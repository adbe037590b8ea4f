"""GPU: client-side replies whose EchoResponse is empty under a CRC32C checksum (found by tools/fuzz_small_host.py on the CPU-emulated k_small:
the pack stage skipped messages with nothing to hand over and with them their Crc32cVerify).  Every path that serves client runs: k_small,
the slot-scan pipeline, k_fused + k_pack_slow.  (Named to sort last: written after the round's last GPU run.)"""
import os

import numpy as np
import pytest

import _oracle as O
from _compare import assert_same
from _traffic import echo_frame, rnd62
from test_device_small_host import empty_reply_frames

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("small,fused", [("on", "on"), ("off", "on"), ("off", "off")])
def test_empty_replies_have_their_checksum_verified(small, fused):
    import random
    import brpc_b200
    os.environ["B2_SMALL"] = small; os.environ["B2_FUSED"] = fused
    try:
        ctx = brpc_b200.Context(device=0, max_batch_bytes=16 << 20, max_msgs=1 << 16, max_runs=256)
    finally:
        os.environ.pop("B2_SMALL"); os.environ.pop("B2_FUSED")
    rng = random.Random(20260921)
    fr = empty_reply_frames()
    # ordinary replies around them so that the tile pipeline has tiles to cut (and k_fused fast messages next to the parked ones)
    cfg = O.make_config()
    per = 10 if small == "on" else 40                                  # <= 128 KB / 1024 messages goes down the one-launch path
    req = [b"".join(echo_frame(rng, 100 * s + j, rnd62(rng, 1024)) for j in range(per)) for s in range(8)]
    d0, r0 = brpc_b200.make_runs(req)
    o = O.process_batch(cfg, d0, r0)
    replies = [bytes(o[2][int(m["resp_off"]):int(m["resp_off"]) + int(m["resp_len"])]) for m in o[1]]
    chunks = []
    for s in range(8):
        part = replies[per * s:per * s + per]
        for j, f in enumerate(fr[s::8]):
            part.insert(min(len(part), 2 + 4 * j), f)
        chunks.append(b"".join(part))
    data, runs = brpc_b200.make_runs(chunks)
    runs["flags"] = 1                                                    # B2_RUN_CLIENT
    for _ in range(2):                                                   # (the second batch runs with the adapted tile size)
        dev = ctx.process_batch(data, runs)
        assert_same(dev, O.process_batch(cfg, data, runs), "small=%s fused=%s" % (small, fused))
    e = dev[1]["error_code"]
    assert int((e == 1003).sum()) == 14 and int((e == 0).sum()) == 8 * per + 1
