"""GPU: the other length-prefixed protocols in the cut loop — hulu_pbrpc, sofa_pbrpc, nshead — against the oracle's
restatement of ParseHuluMessage / ParseSofaMessage / ParseNsheadMessage inside the same CutInputMessage loop.  The fixture is the
reference's own stream builder (test/brpc_input_messenger_unittest.cpp:95-100: 1024 hulu frames of 32 bytes, re-sent with
arbitrary split points), then mixed-protocol connections with preferred-index switching and corruptions."""
import random
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import _oracle as O  # noqa: E402
from _compare import assert_same  # noqa: E402
from _traffic import SEED, echo_frame, rnd62  # noqa: E402

ALL = (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4) | (1 << 12)


def hulu(meta, payload):
    return b"HULU" + struct.pack("<II", len(meta) + len(payload), len(meta)) + meta + payload


def sofa(meta, payload):
    return b"SOFA" + struct.pack("<IQQ", len(meta), len(payload), len(meta) + len(payload)) + meta + payload


def nshead(body, log_id=7):
    return struct.pack("<HHI16sIII", 1, 2, log_id, b"b2-test", 0xfb709394, 0, len(body)) + body


def ctx_and_cfg(b2, mask=ALL, **kw):
    ctx = b2.Context(device=0, max_batch_bytes=32 << 20, max_msgs=1 << 18, max_runs=4096, **kw)
    ctx.set_protocols(mask)
    return ctx, O.make_config(protocols=mask)


def both(b2, ctx, cfg, chunks, preferred=None, what=""):
    data, runs = b2.make_runs(chunks)
    if preferred is not None:
        runs["preferred_proto"] = preferred
    dev = ctx.process_batch(data, runs)
    orc = O.process_batch(cfg, data, runs)
    assert_same(dev, orc, what)
    return dev


def test_reference_hulu_fixture_every_split():
    """brpc_input_messenger_unittest.cpp:95-100: 1024 messages of 32 bytes (12-byte hulu header, meta 0), cut anywhere."""
    import brpc_b200 as b2
    ctx, cfg = ctx_and_cfg(b2)
    msg = hulu(b"", bytes(range(20)))
    assert len(msg) == 32
    stream = msg * 1024
    rng = random.Random(SEED)
    cuts = sorted(set([0, 1, 3, 4, 11, 12, 13, 31, 32, 33, 63, 64, 65] + [rng.randrange(len(stream) + 1) for _ in range(300)] + [len(stream)]))
    for pref in (-1, 3, 1):
        dev = both(b2, ctx, cfg, [stream[:c] for c in cuts], preferred=pref, what="hulu fixture pref=%d" % pref)
        rs, msgs = dev[0], dev[1]
        assert np.all(rs["n_msgs"] == np.array(cuts) // 32) and np.all(msgs["status"] == 9) and np.all(msgs["protocol"] == 3)
        assert np.all(msgs["body_size"] == 20) and np.all(msgs["meta_size"] == 0)
    # a big batch of it goes through the tile pipeline
    dev = both(b2, ctx, cfg, [stream * 8 for _ in range(32)], what="hulu fixture x8")
    assert len(dev[1]) == 32 * 8192


def mixed_stream(rng, n):
    out = []
    for i in range(n):
        c = rng.random()
        body = rnd62(rng, rng.choice([0, 1, 20, 100, 1000, 5000]))
        meta = rnd62(rng, rng.choice([0, 5, 40]))
        if c < 0.25: out.append(hulu(meta, body))
        elif c < 0.45: out.append(sofa(meta, body))
        elif c < 0.65: out.append(nshead(body, log_id=i))
        elif c < 0.90: out.append(echo_frame(rng, i, body))
        elif c < 0.93: out.append(b"HULU" + struct.pack("<II", 10, 50) + bytes(10))          # meta_size > body_size: popped, TRY_OTHERS
        elif c < 0.95: out.append(O.pack_stream_frame(rng.randrange(1 << 30), -1, 3, None, body))
        else: out.append(sofa(meta, body)[:24 - 8] + struct.pack("<Q", 12345) + meta + body)   # msg_size mismatch: TRY_OTHERS, nothing popped -> connection dies
    return out


def test_mixed_protocol_connections_and_preferred_index():
    import brpc_b200 as b2
    rng = random.Random(SEED + 5)
    for mask in (ALL, (1 << 1) | (1 << 3), (1 << 3) | (1 << 4) | (1 << 12)):
        ctx, cfg = ctx_and_cfg(b2, mask)
        streams = [b"".join(mixed_stream(rng, rng.randrange(1, 60))) for _ in range(80)]
        chunks = [s[:rng.randrange(len(s) + 1)] if rng.random() < 0.5 else s for s in streams]
        for pref in (-1, 1, 3, 4, 12):
            both(b2, ctx, cfg, chunks, preferred=pref, what="mixed mask=%x pref=%d" % (mask, pref))
    # short prefixes of every header, every preferred index (nshead answers NOT_ENOUGH_DATA below 28 bytes whatever the bytes are)
    ctx, cfg = ctx_and_cfg(b2)
    frames = [hulu(b"m", b"payload"), sofa(b"me", b"payload!"), nshead(b"0123456789"), echo_frame(rng, 1, b"x")]
    chunks = [f[:n] for f in frames for n in range(0, 40)] + [b"HU", b"SOF", b"SOFB", b"XXXX" * 9, b"\x00" * 27, b"\x00" * 28]
    for pref in (-1, 1, 2, 3, 4, 12):
        both(b2, ctx, cfg, chunks, preferred=pref, what="prefixes pref=%d" % pref)


def test_large_mixed_batch_through_the_tile_pipeline(monkeypatch):
    import brpc_b200 as b2
    monkeypatch.setenv("B2_SMALL", "off")
    rng = random.Random(SEED + 6)
    ctx, cfg = ctx_and_cfg(b2)
    good = [f for f in mixed_stream(rng, 3000) if not (f[:4] == b"SOFA" and struct.unpack("<Q", f[16:24])[0] == 12345)]
    streams = [b"".join(rng.sample(good, 400)) for _ in range(48)]
    both(b2, ctx, cfg, streams, what="large mixed")
    ctx.set_modes(1, 1)
    from brpc_b200.abi import PinnedBuffer
    data, runs = b2.make_runs(streams)
    pin = PinnedBuffer(len(data)); pin.array[:] = data
    rs, msgs, resp, info = ctx.process_batch_ptr(pin.ptr, len(data), runs)
    o_rs, o_msgs, o_resp = O.process_batch(cfg, data, runs)
    assert np.array_equal(msgs["status"], o_msgs["status"]) and np.array_equal(msgs["frame_off"], o_msgs["frame_off"]) and np.array_equal(rs["consumed"], o_rs["consumed"])
