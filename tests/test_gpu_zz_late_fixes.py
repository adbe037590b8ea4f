"""GPU: the two parity bugs found on the CPU after the round's last GPU run (profiles/r2_emulator.md), on every pipeline that serves client runs
(k_small, the slot-scan pipeline, k_fused + k_pack_slow):
  * client-side replies whose EchoResponse is empty under a CRC32C checksum (tools/fuzz_small_host.py: the pack stage skipped messages with
    nothing to hand over and with them their Crc32cVerify);
  * client-side sockets with more than baidu_std / streaming_rpc enabled (tools/fuzz_emul.py: the channel's protocol is fixed, so what a tile
    holds depends on the message before it; the speculative tile walk assumed it does not).
(Named to sort last: these tests have only run on the emulated library, tests/test_emulated_library.py.)"""
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


def _hulu(meta, payload):
    import struct
    return b"HULU" + struct.pack("<II", len(meta) + len(payload), len(meta)) + meta + payload


def _sofa(meta, payload):
    import struct
    return b"SOFA" + struct.pack("<IQQ", len(meta), len(payload), len(meta) + len(payload)) + meta + payload


@pytest.mark.parametrize("small,fused,tile", [("off", "on", 1024), ("off", "off", 1024), ("off", "on", 0), ("on", "on", 0)])
def test_client_channel_protocol_is_fixed_also_across_tiles(small, fused, tile):
    """On a client-side socket the first message fixes the protocol (input_messenger.cpp:122-138): a baidu_std reply behind a hulu message is an
    error there, not a message — also when the two fall into different tiles (found by tools/fuzz_emul.py: the speculative tile walk does not
    depend on the preferred index on server sockets, on these it does; k_resolve now takes the exact chain for them)."""
    import random
    import brpc_b200
    os.environ["B2_SMALL"] = small; os.environ["B2_FUSED"] = fused
    try:
        ctx = brpc_b200.Context(device=0, max_batch_bytes=16 << 20, max_msgs=1 << 16, max_runs=256, tile_bytes=tile)
    finally:
        os.environ.pop("B2_SMALL"); os.environ.pop("B2_FUSED")
    mask = (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4)
    ctx.set_protocols(mask)
    cfg = O.make_config(protocols=mask)
    rng = random.Random(20260922)
    d0, r0 = brpc_b200.make_runs([b"".join(echo_frame(rng, j, rnd62(rng, rng.choice([10, 300, 1024]))) for j in range(24))])
    o = O.process_batch(O.make_config(), d0, r0)
    replies = [bytes(o[2][int(m["resp_off"]):int(m["resp_off"]) + int(m["resp_len"])]) for m in o[1]]
    hulus = [_hulu(rnd62(rng, 5), rnd62(rng, rng.choice([20, 400, 900]))) for _ in range(24)]
    sofas = [_sofa(rnd62(rng, 7), rnd62(rng, rng.choice([20, 400, 900]))) for _ in range(24)]
    strm = O.pack_stream_frame(77, 78, 3, False, b"stream-data")
    chunks = [b"".join(hulus) + b"".join(replies),                    # the channel speaks hulu: the first PRPC frame ends it
              b"".join(replies) + b"".join(hulus),                    # ... baidu_std: the first HULU frame ends it
              b"".join(replies[:12]) + strm + b"".join(replies[12:]) + b"".join(sofas),      # baidu_std <-> streaming_rpc is the one allowed pair
              b"".join(sofas[:3]) + strm + b"".join(sofas),
              b"".join(hulus[:7]) + sofas[0] + b"".join(hulus),
              b"".join(replies), b"".join(hulus), b"".join(sofas)]
    if small == "on":
        chunks = [c[:9000] for c in chunks]                                 # <= 128 KB: the one-launch path
    data, runs = brpc_b200.make_runs(chunks)
    runs["flags"] = 1                                                    # B2_RUN_CLIENT
    for pref in (-1, 1, 3, 4):
        runs["preferred_proto"] = pref
        for _ in range(2):
            dev = ctx.process_batch(data, runs)
            assert_same(dev, O.process_batch(cfg, data, runs), "client channel pref=%d small=%s fused=%s tile=%d" % (pref, small, fused, tile))
    runs["preferred_proto"] = -1
    dev = ctx.process_batch(data, runs)
    n = dev[0]["n_msgs"]
    if small == "off":                                                  # (whole runs: what each channel accepted before the foreign frame)
        assert list(n) == [24, 24, 25, 3, 7, 24, 24, 24]


def test_by_reference_entries_of_client_side_messages_are_written():
    """B2_RESP_BY_REF: refs[i] is part of the output for EVERY message; the client-side branch of the decoder left it as the memory came
    (tools/fuzz_emul.py on the emulated library, whose cudaMalloc memory arrives filled with 0xa5)."""
    import random
    import brpc_b200
    from brpc_b200.abi import PinnedBuffer
    rng = random.Random(20260923)
    m = dict(brpc_b200.abi.ECHO_METHOD, response_compress_type=1)
    d0, r0 = brpc_b200.make_runs([b"".join(echo_frame(rng, 100 * s + j, rnd62(rng, rng.choice([0, 40, 1024]))) for j in range(30)) for s in range(6)])
    o = O.process_batch(O.make_config(methods=[m]), d0, r0)
    replies = [bytes(o[2][int(x["resp_off"]):int(x["resp_off"]) + int(x["resp_len"])]) for x in o[1]]           # snappy-compressed EchoResponses
    data, runs = brpc_b200.make_runs([b"".join(replies[30 * s:30 * s + 30]) for s in range(6)])
    runs["flags"] = 1
    pin = PinnedBuffer(len(data)); pin.array[:] = data
    for small in ("on", "off"):
        os.environ["B2_SMALL"] = small
        try:
            ctx = brpc_b200.Context(device=0, max_batch_bytes=16 << 20, max_msgs=1 << 16, max_runs=256)
        finally:
            os.environ.pop("B2_SMALL")
        for im in (0, 1):
            ctx.set_modes(im, 1)
            rs, msgs, resp, info = ctx.process_batch_ptr(pin.ptr, len(data), runs)
            assert len(msgs) == 180 and set(msgs["status"].tolist()) <= {7, 8} and int((msgs["status"] == 8).sum()) > 100
            refs = info["refs"]
            assert refs is not None and not refs["src_len"].any() and not refs["prefix_len"].any() and not refs["src_off"].any()
        ctx.close()
    pin.free()
