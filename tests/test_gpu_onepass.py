"""GPU: k_onepass (search + walk + decode + echo + pack of a tile from one load of its bytes, on speculation) with k_resolve as the
exactness gate.  (1) bench-shaped traffic is served by it and equals the oracle; (2) traffic it cannot serve on speculation — payloads that
hold byte strings looking like frames, frames larger than a tile, errors inside a run — is flagged by k_resolve and re-run on the staged
pipeline with the same (oracle-identical) results; (3) both pipelines agree on mixed traffic."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import _oracle as O  # noqa: E402
from _compare import assert_same  # noqa: E402
from _traffic import SEED, echo_frame, mixed_frames, rnd62, split_runs  # noqa: E402


def _ctx(b2, **kw):
    return b2.Context(device=0, max_batch_bytes=64 << 20, max_msgs=1 << 17, max_runs=256, **kw)


def test_onepass_serves_pipelined_echo_traffic():
    import brpc_b200 as b2
    rng = random.Random(SEED + 501)
    ctx = _ctx(b2)
    cfg = O.make_config()
    for trial in range(3):
        streams = []
        for s in range(48):
            n = rng.choice([1000, 1024, 1100]) if trial < 2 else rng.choice([600, 900, 1024, 1300])
            streams.append(b"".join(echo_frame(rng, s * 1000 + i, rnd62(rng, n) if i % 3 else b"r" * n,
                                               attachment=rnd62(rng, 17) if (trial == 1 and i % 5 == 0) else b"") for i in range(rng.randrange(150, 260))))
        streams = [st[:len(st) - rng.randrange(0, 700)] for st in streams]             # runs end inside a frame, as a read would
        data, runs = b2.make_runs(streams)
        dev = ctx.process_batch(data, runs)
        info = ctx.batch_info()
        assert_same(dev, O.process_batch(cfg, data, runs), "trial %d" % trial)
        assert info["onepass"], info
    assert ctx.batch_info()["tile_bytes"] == 8192


def test_onepass_is_turned_down_and_the_staged_pipeline_answers():
    import brpc_b200 as b2
    rng = random.Random(SEED + 502)
    ctx = _ctx(b2)
    cfg = O.make_config()
    fake = echo_frame(rng, 7, b"x" * 40)                                  # a whole valid frame, to be carried inside payloads
    cases = []
    # payloads that hold frame look-alikes right where a tile starts its search
    cases.append([b"".join(echo_frame(rng, s * 1000 + i, (fake * 30)[:1024]) for i in range(200)) for s in range(16)])
    # frames larger than a tile between small ones
    cases.append([b"".join(echo_frame(rng, s * 1000 + i, rnd62(rng, 20000 if i % 7 == 3 else 1024)) for i in range(120)) for s in range(16)])
    # garbage in the middle of a run (the socket would be closed there)
    cases.append([b"".join(echo_frame(rng, s * 1000 + i, rnd62(rng, 1024)) for i in range(90)) + (b"\x00garbage!" * 3 if s % 2 else b"") +
                  b"".join(echo_frame(rng, s * 1000 + 500 + i, rnd62(rng, 1024)) for i in range(90)) for s in range(16)])
    for k, streams in enumerate(cases):
        ctx2 = _ctx(b2)                                                    # (a fresh context: no skip window from the previous case)
        warm = [b"".join(echo_frame(rng, i, rnd62(rng, 1024)) for i in range(200)) for _ in range(16)]
        d0, r0 = b2.make_runs(warm)
        ctx2.process_batch(d0, r0); ctx2.process_batch(d0, r0)
        assert ctx2.batch_info()["onepass"]
        data, runs = b2.make_runs(streams)
        dev = ctx2.process_batch(data, runs)
        assert_same(dev, O.process_batch(cfg, data, runs), "case %d" % k)
        assert not ctx2.batch_info()["onepass"], "case %d should have been turned down" % k
        dev = ctx2.process_batch(d0, r0)                                   # the skip window: staged for a while, still right
        assert_same(dev, O.process_batch(cfg, d0, r0), "after case %d" % k)
        assert not ctx2.batch_info()["onepass"]


def test_both_pipelines_agree_on_mixed_traffic(monkeypatch):
    import brpc_b200 as b2
    rng = random.Random(SEED + 503)
    cfg = O.make_config()
    streams = [mixed_frames(rng, rng.randrange(100, 300), big=False) for _ in range(40)]
    chunks = split_runs(rng, streams)
    data, runs = b2.make_runs(chunks)
    want = O.process_batch(cfg, data, runs)
    ctx = _ctx(b2, tile_bytes=8192)
    for rep in range(3):
        assert_same(ctx.process_batch(data, runs), want, "default rep %d" % rep)
    monkeypatch.setenv("B2_ONEPASS", "off")
    ctx_off = _ctx(b2, tile_bytes=8192)
    assert_same(ctx_off.process_batch(data, runs), want, "B2_ONEPASS=off")
    assert not ctx_off.batch_info()["onepass"]
