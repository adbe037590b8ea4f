"""Adversarial probe (not product): requests whose payloads are themselves valid-looking baidu_std frame chains, so that the
speculative tile search keeps picking fakes and k_resolve has to re-walk.  Prints stage times next to a benign batch."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import brpc_b200
import _oracle as O
from _traffic import echo_frame, rnd62
rng = random.Random(5)
def batch(adversarial):
    streams = []
    for s in range(64):
        fr = []
        while sum(len(f) for f in fr) < (1 << 20):
            if adversarial:
                inner = b"".join(echo_frame(rng, 100 + k, rnd62(rng, rng.choice([10, 200, 700]))) for k in range(rng.randrange(2, 8)))
                fr.append(echo_frame(rng, len(fr), rnd62(rng, rng.randrange(0, 50)) + inner))
            else:
                fr.append(echo_frame(rng, len(fr), rnd62(rng, 1500)))
        streams.append(b"".join(fr))
    return brpc_b200.make_runs(streams)
ctx = brpc_b200.Context(device=0, max_batch_bytes=128 << 20, max_msgs=1 << 20, max_runs=64)
for name, adv in (("benign", False), ("adversarial", True)):
    data, runs = batch(adv)
    rs, msgs, resp, info = ctx.process_batch(data, runs)
    ors, om, oresp = O.process_batch(O.make_config(), data, runs)
    assert len(msgs) == len(om) and np.array_equal(msgs["frame_off"], om["frame_off"])
    ctx.upload(data, runs)
    acc = {}
    for _ in range(5):
        ctx.execute()
        for k, v in ctx.stage_times(): acc.setdefault(k, []).append(v)
    print(name, "msgs", len(msgs), "MB", round(len(data) / 1e6, 1), {k: round(1000 * sum(v) / len(v)) for k, v in acc.items()})
