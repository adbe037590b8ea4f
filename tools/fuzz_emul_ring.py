"""One-off fuzz run on the CPU (not part of the test suite): the submit ring of the EMULATED library (persistent k_ring on its own thread) fed
with random small batches of the fuzz traffic — pageable input staged through the slot or pinned input read in place, up to 6 tickets in
flight, replies copied / by reference / as iovec pairs, batches over the compact block's capacity (served by the big pipeline from inside
b2_ring_wait), idle retirements — against the oracle.  Usage: python tools/fuzz_emul_ring.py [seconds] [base seed]"""
import ctypes, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("B2_RING_IDLE_MS", "5")
import fuzz_emul as E                     # (points ctypes at the emulated library)
import numpy as np
import brpc_b200
from brpc_b200.abi import PinnedBuffer, ECHO_METHOD
from _compare import assert_same, MSG_FIELDS, RUN_FIELDS
from test_gpu_modes import replies
O = E.O; F = E.F


def run(budget, base_seed):
    t0 = time.time(); seed = 0; total = 0; batches = 0
    while time.time() - t0 < budget:
        seed += 1
        rng = random.Random(base_seed + seed)
        mask = rng.choice([F.ALL, (1 << 1) | (1 << 2), (1 << 1) | (1 << 3)])
        m = dict(ECHO_METHOD, response_checksum_type=rng.choice([0, 0, 1]), response_compress_type=rng.choice([0, 0, 1]), echo_attachment=rng.choice([0, 1]))
        sth = rng.choice([0, 1]); rm = rng.choice([0, 0, 1, 2])
        ctx = brpc_b200.Context(device=0, max_batch_bytes=8 << 20, max_msgs=1 << 14, max_runs=1024, methods=[m], stream_handler=sth)
        ctx.set_protocols(mask); ctx.set_modes(0, rm)
        cfg = O.make_config(methods=[m], protocols=mask, stream_handler=sth)
        ctx.ring_start()
        inflight = []; keep = []

        def finish(item):
            t, data, runs, pin = item
            rs, msgs, resp, info = ctx.ring_wait(t)
            o_rs, o_msgs, o_resp = O.process_batch(cfg, data, runs)
            what = "ring seed %d rm %d" % (base_seed + seed, rm)
            if rm == 0:
                assert_same((rs, msgs, resp), (o_rs, o_msgs, o_resp), what); return len(msgs)
            for f in RUN_FIELDS: assert np.array_equal(rs[f], o_rs[f]), what + " run." + f
            assert len(msgs) == len(o_msgs), what
            for f in MSG_FIELDS: assert np.array_equal(msgs[f], o_msgs[f]), what + " msgs." + f
            want = replies(data, o_msgs, o_resp, None)
            if rm == 1:
                got = replies(pin.array if pin else data, msgs, resp, info["refs"])
            else:
                iov = info["iov"]; answered = (msgs["status"] == 0) | (msgs["status"] == 1)
                got = [b"".join(ctypes.string_at(int(iov["base"][j]), int(iov["len"][j])) for j in (2 * k, 2 * k + 1) if iov["len"][j]) for k in range(len(msgs))]
                want = [w if answered[k] else b"" for k, w in enumerate(want)]
            for k, (g, w) in enumerate(zip(got, want)): assert g == w, "%s reply %d differs (status %d)" % (what, k, msgs["status"][k])
            return len(msgs)

        for k in range(rng.randrange(3, 14)):
            client = rng.random() < 0.2
            chunks = []; room = rng.choice([4 << 10, 40 << 10, 125 << 10])
            for s in range(rng.randrange(1, 50)):
                b = bytearray(b"".join(E.frame(rng, j, client, sth) for j in range(rng.randrange(1, 8))))
                if rng.random() < 0.25 and len(b) > 20: b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
                if rng.random() < 0.3: b = b[:rng.randrange(len(b) + 1)]
                if len(b) + 16 > room: continue
                room -= len(b) + 16; chunks.append(bytes(b))
            if rng.random() < 0.07: chunks = [b"".join(E.echo_frame(rng, i, b"") for i in range(1400))]      # over the compact block: big pipeline
            if not chunks: continue
            data, runs = brpc_b200.make_runs(chunks)
            runs["flags"] = 1 if client else 0
            runs["preferred_proto"] = np.array([rng.choice([-1, -1, 1, 2, 3]) for _ in chunks], dtype=np.int32)
            alone = len(data) // 40 > 900               # may not fit the compact block: b2_ring_wait then runs the big pipeline, which wants the other tickets collected
            if alone:
                for it in inflight: total += finish(it); batches += 1
                inflight = []
            pin = None
            if rm == 2 or rng.random() < 0.5:                # (iovec entries point into the caller's bytes: pinned input)
                pin = PinnedBuffer(max(len(data), 16)); pin.array[:len(data)] = data; keep.append(pin)
                t = ctx.ring_submit(None, runs, ptr=pin.ptr, nbytes=len(data))
            else:
                t = ctx.ring_submit(data, runs)
            inflight.append((t, data, runs, pin))
            if alone or len(inflight) >= rng.randrange(1, 7): total += finish(inflight.pop(0)); batches += 1
            if rng.random() < 0.1: time.sleep(0.02)                                  # (long enough for the kernel to retire)
        for it in inflight: total += finish(it); batches += 1
        ctx.ring_stop(); ctx.close()
        for p in keep: p.free()
    return batches, seed, total


if __name__ == "__main__":
    b, sd, m = run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 9700000)
    print("ring fuzz ok: %d batches (%d contexts), %d messages, emulated ring == oracle everywhere" % (b, sd, m))
