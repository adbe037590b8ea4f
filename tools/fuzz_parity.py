"""One-off fuzz run (not part of the test suite): many seeds of mixed / corrupted baidu_std + streaming traffic through the
tile pipeline and the one-launch path, device vs oracle bit for bit.  Usage: python tools/fuzz_parity.py [seconds]"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import brpc_b200
import _oracle as O
from _compare import assert_same
from _traffic import mixed_frames, split_runs, echo_frame, rnd62

def run(budget=60.0, base_seed=900000):
    t0 = time.time(); seed = 0; msgs_total = 0; batches = 0
    ctxs = {}
    while time.time() - t0 < budget:
        seed += 1
        rng = random.Random(base_seed + seed)
        tile = rng.choice([0, 512, 1024, 4096, 8192, 16384])
        os.environ["B2_SMALL"] = rng.choice(["on", "off"])
        key = (tile, os.environ["B2_SMALL"])
        if key not in ctxs:
            ctxs[key] = brpc_b200.Context(device=0, max_batch_bytes=64 << 20, max_msgs=1 << 20, max_runs=8192, tile_bytes=tile, server_identity=b"10.1.2.3:8000")
        ctx = ctxs[key]
        streams = []
        for s in range(rng.randrange(1, 120)):
            kind = rng.random()
            if kind < 0.6: fr = mixed_frames(rng, rng.randrange(1, 60), big=rng.random() < 0.2)
            elif kind < 0.8: fr = [echo_frame(rng, k, b"r" * rng.choice([0, 3, 16, 40])) for k in range(rng.randrange(1, 600))]
            else: fr = [echo_frame(rng, k, rnd62(rng, rng.choice([1000, 1024, 5000]))) for k in range(rng.randrange(1, 40))]
            b = bytearray(b"".join(fr))
            if rng.random() < 0.25 and len(b) > 20:
                for _ in range(rng.randrange(1, 4)):
                    b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
            if rng.random() < 0.1: b = b[rng.randrange(0, 12):]
            streams.append([bytes(b)])
        chunks = split_runs(rng, streams)
        data, runs = brpc_b200.make_runs(chunks)
        runs["preferred_proto"] = np.array([rng.choice([-1, -1, 1, 2]) for _ in range(len(runs))], dtype=np.int32)
        dev = ctx.process_batch(data, runs)
        orc = O.process_batch(O.make_config(server_identity=b"10.1.2.3:8000"), data, runs)
        assert_same(dev, orc, "fuzz seed %d tile %d small %s" % (seed, tile, os.environ["B2_SMALL"]))
        msgs_total += len(dev[1]); batches += 1
    return batches, seed, msgs_total


if __name__ == "__main__":
    b, sd, m = run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0)
    print("fuzz ok: %d batches (%d seeds), %d messages, device == oracle everywhere" % (b, sd, m))
