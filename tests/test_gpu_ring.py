"""GPU: the persistent latency kernel (b2_ring_submit / b2_ring_wait) gives bit-exact the results of the oracle — same
small batches as the latency path of b2_process_batch, but with no launch / memcpy / synchronise per batch."""
import random
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import _oracle as O  # noqa: E402
from _compare import assert_same  # noqa: E402
from _traffic import SEED, echo_frame, mixed_frames, rnd62, split_runs  # noqa: E402


def test_ring_small_batches_bit_exact_pipelined_and_relaunched():
    import brpc_b200 as b2
    from brpc_b200.abi import PinnedBuffer
    rng = random.Random(SEED + 91)
    ctx = b2.Context(device=0, max_batch_bytes=8 << 20, max_msgs=1 << 14, max_runs=1024)
    cfg = O.make_config()
    ctx.ring_start()
    batches = []
    for k in range(60):
        streams = [mixed_frames(rng, rng.randrange(1, 6)) for _ in range(rng.randrange(1, 40))]
        chunks = split_runs(rng, streams)
        while sum(len(c) + 16 for c in chunks) > (120 << 10):
            chunks.pop()
        batches.append(b2.make_runs(chunks))
    # one at a time, pageable input (staged through the slot) ...
    for data, runs in batches[:20]:
        t = ctx.ring_submit(data, runs)
        assert_same(ctx.ring_wait(t), O.process_batch(cfg, data, runs), "ring one-at-a-time")
    # ... pinned input read in place, several tickets in flight ...
    pins = []
    for data, runs in batches[20:]:
        p = PinnedBuffer(max(len(data), 16)); p.array[:len(data)] = data; pins.append(p)
    inflight = []
    for k, (data, runs) in enumerate(batches[20:]):
        inflight.append((ctx.ring_submit(None, runs, ptr=pins[k].ptr, nbytes=len(data)), data, runs))
        if len(inflight) == 6:
            t, d, r = inflight.pop(0)
            assert_same(ctx.ring_wait(t), O.process_batch(cfg, d, r), "ring pipelined")
    for t, d, r in inflight:
        assert_same(ctx.ring_wait(t), O.process_batch(cfg, d, r), "ring drain")
    n0 = ctx.ring_launches()
    assert n0 >= 1
    # ... the kernel retires when idle and comes back with the next submission
    time.sleep(0.2)
    data, runs = batches[3]
    t = ctx.ring_submit(data, runs)
    assert_same(ctx.ring_wait(t), O.process_batch(cfg, data, runs), "ring after idle")
    assert ctx.ring_launches() == n0 + 1
    # by-reference replies through the ring
    ctx.set_modes(0, 1)
    frames = [echo_frame(rng, i, rnd62(rng, 1024)) for i in range(64)]
    data, runs = b2.make_runs(frames)
    t = ctx.ring_submit(data, runs)
    rs, msgs, resp, info = ctx.ring_wait(t)
    o_rs, o_msgs, o_resp = O.process_batch(cfg, data, runs)
    assert len(msgs) == 64 and np.all(info["refs"]["src_len"] == 1024)
    for i in range(64):
        p = int(info["refs"]["prefix_len"][i]); so = int(info["refs"]["src_off"][i])
        got = bytes(resp[int(msgs["resp_off"][i]):int(msgs["resp_off"][i]) + p]) + bytes(data[so:so + 1024])
        assert got == bytes(o_resp[int(o_msgs["resp_off"][i]):int(o_msgs["resp_off"][i]) + int(o_msgs["resp_len"][i])])
    ctx.set_modes(0, 0)
    # a batch whose replies overflow the compact block is served by the big pipeline from inside ring_wait
    frames = [echo_frame(rng, i, b"") for i in range(1500)]
    data, runs = b2.make_runs([b"".join(frames)])
    t = ctx.ring_submit(data, runs)
    assert_same(ctx.ring_wait(t), O.process_batch(cfg, data, runs), "ring overflow fallback")
    ctx.ring_stop()
    ctx.close()
