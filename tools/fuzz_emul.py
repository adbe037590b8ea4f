"""One-off fuzz run on the CPU (not part of the test suite): mixed / corrupted / truncated traffic of all five framings, codecs, attachments,
limits, both sides — the generator of tools/fuzz_small_host.py — through the C ABI of the EMULATED library (tests/cpp/libb2rpc_emul.so, see
tests/test_emulated_library.py), over random tile sizes, the one-launch path on / off, k_fused on / off, against the oracle bit for bit.
Usage: python tools/fuzz_emul.py [seconds] [base seed]"""
import ctypes, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import emul_runner
ctypes.CDLL = emul_runner.EmulCDLL
import numpy as np
import brpc_b200
import _oracle as O
from _compare import assert_same
from brpc_b200.abi import ECHO_METHOD
import fuzz_small_host as F
from _compare import MSG_FIELDS, RUN_FIELDS
from test_gpu_modes import replies
from brpc_b200.abi import PinnedBuffer

MODES = [(1, 0), (0, 1), (1, 1), (0, 2), (1, 2)]
from _traffic import mixed_frames, echo_frame, rnd62


def frame(rng, j, client, sth):
    """fuzz_small_host's frames plus: every branch of ProcessRpcRequest (mixed_frames), frames larger than k_fused's staging buffer,
    streaming DATA frames whose payload is a (sometimes damaged) snappy stream when the stream handler decompresses"""
    c = rng.random()
    if c < 0.25:
        return b"".join(mixed_frames(rng, 1, big=rng.random() < 0.15))
    if c < 0.32 and sth:
        f = echo_frame(rng, j, rnd62(rng, rng.choice([0, 30, 900, 5000])), compress_type=1)
        z = bytearray(f[12 + int.from_bytes(f[8:12], "big"):])                       # a snappy stream
        if rng.random() < 0.3 and z: z[rng.randrange(len(z))] ^= 1 << rng.randrange(8)
        return O.pack_stream_frame(rng.randrange(1 << 40), rng.choice([-1, 77]), 3, rng.choice([None, True, False]), bytes(z))
    return F.one_frame(rng, j, client)


def check_modes(ctx, data, runs, orc, rng, what):
    """B2_INPUT_PULL / B2_RESP_BY_REF / B2_RESP_IOVEC give what the copy path gives (tests/test_gpu_modes.py's comparison, on this traffic)"""
    o_rs, o_msgs, o_resp = orc
    pin = PinnedBuffer(max(len(data), 16)); pin.array[:len(data)] = data
    want = replies(data, o_msgs, o_resp, None)
    for im, rm in rng.sample(MODES, 2):
        ctx.set_modes(im, rm)
        rs, msgs, resp, info = ctx.process_batch_ptr(pin.ptr, len(data), runs)
        tag = "%s input=%d resp=%d" % (what, im, rm)
        for f in RUN_FIELDS:
            assert np.array_equal(rs[f], o_rs[f]), tag + " run." + f
        assert len(msgs) == len(o_msgs), tag
        for f in MSG_FIELDS:
            assert np.array_equal(msgs[f], o_msgs[f]), tag + " msgs." + f
        if rm == 2 and len(msgs):
            iov = info["iov"]; answered = (msgs["status"] == 0) | (msgs["status"] == 1)
            for k in range(len(msgs)):
                g = b"".join(ctypes.string_at(int(iov["base"][j]), int(iov["len"][j])) for j in (2 * k, 2 * k + 1) if iov["len"][j])
                assert g == (want[k] if answered[k] else b""), "%s iovec reply %d differs (status %d)" % (tag, k, msgs["status"][k])
            continue
        got = replies(pin.array, msgs, resp, info["refs"])
        for k, (g, w) in enumerate(zip(got, want)):
            assert g == w, "%s reply %d differs (status %d)" % (tag, k, msgs["status"][k])
    ctx.set_modes(0, 0); pin.free()


def run(budget, base_seed, first=0):
    t0 = time.time(); seed = first; total = 0; batches = 0
    while time.time() - t0 < budget:
        seed += 1
        rng = random.Random(base_seed + seed)
        mask = rng.choice([F.ALL, (1 << 1) | (1 << 2), (1 << 1) | (1 << 3), (1 << 3) | (1 << 4) | (1 << 12)])
        m = dict(ECHO_METHOD, response_checksum_type=rng.choice([0, 0, 1]), response_compress_type=rng.choice([0, 0, 1]), echo_attachment=rng.choice([0, 1]))
        identity = rng.choice([None, b"10.1.2.3:8000"]); sth = rng.choice([0, 1]); max_body = rng.choice([0, 0, 600, 3000])
        tile = rng.choice([0, 512, 1024, 4096, 16384])
        os.environ["B2_SMALL"] = rng.choice(["on", "off", "off"]); os.environ["B2_FUSED"] = rng.choice(["on", "on", "off"])
        ms = [m] + ([dict(ECHO_METHOD, method_name=b"Echo2", handler=0)] if rng.random() < 0.5 else [])          # (Echo2: a host-handled method)
        ctx = brpc_b200.Context(device=0, max_batch_bytes=8 << 20, max_msgs=1 << 16, max_runs=1024, tile_bytes=tile, server_identity=identity,
                                methods=ms, stream_handler=sth, max_body_size=max_body)
        ctx.set_protocols(mask)
        cfg = O.make_config(methods=ms, server_identity=identity, protocols=mask, max_body_size=max_body, stream_handler=sth)
        client = rng.random() < 0.25
        chunks = []
        for s in range(rng.randrange(1, 90)):
            b = bytearray(b"".join(frame(rng, j, client, sth) for j in range(rng.randrange(1, 25))))
            if rng.random() < 0.3 and len(b) > 20:
                for _ in range(rng.randrange(1, 4)): b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
            if rng.random() < 0.1: b = b[rng.randrange(0, 13):]
            if rng.random() < 0.4: b = b[:rng.randrange(len(b) + 1)]
            chunks.append(bytes(b))
        data, runs = brpc_b200.make_runs(chunks)
        runs["preferred_proto"] = np.array([rng.choice([-1, -1, 1, 2, 3, 4, 12]) for _ in chunks], dtype=np.int32)
        runs["flags"] = 1 if client else 0
        for rep in range(2):                                   # (the second pass runs with the adapted tile size and, maybe, the other pipeline)
            dev = ctx.process_batch(data, runs)
            assert_same(dev, O.process_batch(cfg, data, runs), "fuzz seed %d tile %d small %s fused %s pass %d" % (base_seed + seed, tile, os.environ["B2_SMALL"], os.environ["B2_FUSED"], rep))
        if seed % 3 == 0:
            print("modes at seed", base_seed + seed, flush=True) if os.environ.get("B2_FUZZ_VERBOSE") else None
            check_modes(ctx, data, runs, O.process_batch(cfg, data, runs), rng, "fuzz seed %d modes" % (base_seed + seed))
        total += len(dev[1]); batches += 1
        ctx.close()
    return batches, seed, total


if __name__ == "__main__":
    b, sd, m = run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 9000000, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    print("fuzz ok: %d batches (%d seeds), %d messages, emulated library == oracle everywhere" % (b, sd, m))
