"""One-off fuzz run on the CPU (not part of the test suite): the whole k_small kernel on the emulated 512-thread block
(tests/test_device_small_host.py builds it) against the oracle, over many seeds of mixed / corrupted traffic of all five framings,
codecs, attachments, body-size limits, preferred-index settings, server and client side.
Usage: python tools/fuzz_small_host.py [seconds] [base seed]"""
import os, random, struct, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _oracle as O
from _compare import assert_same
from _traffic import mixed_frames, echo_frame, echo_pb, raw_request_frame, raw_response_frame, rnd62
import test_device_small_host as T
from brpc_b200.abi import ECHO_METHOD

ALL = (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4) | (1 << 12)


def hulu(meta, payload): return b"HULU" + struct.pack("<II", len(meta) + len(payload), len(meta)) + meta + payload
def sofa(meta, payload): return b"SOFA" + struct.pack("<IQQ", len(meta), len(payload), len(meta) + len(payload)) + meta + payload
def nshead(body, log_id=7): return struct.pack("<HHI16sIII", 1, 2, log_id, b"b2-test", 0xfb709394, 0, len(body)) + body


def one_frame(rng, i, client):
    c = rng.random()
    body = rnd62(rng, rng.choice([0, 1, 20, 100, 700, 2500]))
    if c < 0.45:
        kw = {}
        if rng.random() < 0.3: kw["checksum_type"] = 1
        if rng.random() < 0.3: kw["compress_type"] = 1
        if rng.random() < 0.3: kw["attachment"] = rnd62(rng, rng.choice([1, 30, 500]))
        return echo_frame(rng, i, body, **kw)
    if c < 0.60:
        fmt = rng.choice([2, 3])
        co = zlib.compressobj(rng.choice([0, 1, 6, 9]), zlib.DEFLATED, 31 if fmt == 2 else 15)
        z = co.compress(echo_pb(body)) + co.flush()
        if rng.random() < 0.3:
            z = bytearray(z); z[rng.randrange(len(z))] ^= 1 << rng.randrange(8); z = bytes(z)
        if rng.random() < 0.1: z = z[:rng.randrange(len(z) + 1)]
        mk = raw_response_frame if client else raw_request_frame
        return mk(z, 5000 + i, compress_type=fmt, attachment=rng.choice([b"", b"a" * 9]))
    if c < 0.70: return b"".join(mixed_frames(rng, 1))
    meta = rnd62(rng, rng.choice([0, 5, 40]))
    if c < 0.78: return hulu(meta, body)
    if c < 0.86: return sofa(meta, body)
    if c < 0.93: return nshead(body, log_id=i)
    if c < 0.96: return O.pack_stream_frame(rng.randrange(1 << 30), -1, rng.randrange(5), None, body)
    return bytes(rng.randrange(256) for _ in range(rng.randrange(1, 40)))


def run(budget, base_seed):
    sh = T.sh.__wrapped__()
    t0 = time.time(); seed = 0; total = 0; batches = 0
    while time.time() - t0 < budget:
        seed += 1
        rng = random.Random(base_seed + seed)
        mask = rng.choice([ALL, (1 << 1) | (1 << 2), (1 << 1) | (1 << 3), (1 << 3) | (1 << 4) | (1 << 12)])
        ms = [dict(ECHO_METHOD, response_checksum_type=rng.choice([0, 0, 1]), response_compress_type=rng.choice([0, 0, 1]), echo_attachment=rng.choice([0, 1]))]
        identity = rng.choice([None, b"10.1.2.3:8000"]); sth = rng.choice([0, 1])
        max_body = rng.choice([0, 0, 600, 3000])
        k = sh.sh_create(max_body, mask, 0, sth, identity)
        for m in ms:
            sh.sh_add_method(k, m["service_full_name"], m["service_name"], m["method_name"], m["request_type_name"], m["handler"], m["echo_attachment"],
                             m["response_checksum_type"], m["response_compress_type"])
        cfg = O.make_config(methods=ms, server_identity=identity, protocols=mask, max_body_size=max_body, stream_handler=sth)
        client = rng.random() < 0.25
        chunks = []; budget_b = 120 << 10
        for s in range(rng.randrange(1, 60)):
            b = bytearray(b"".join(one_frame(rng, j, client) for j in range(rng.randrange(1, 10))))
            if rng.random() < 0.3 and len(b) > 20:
                for _ in range(rng.randrange(1, 4)): b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
            if rng.random() < 0.1: b = b[rng.randrange(0, 13):]
            if rng.random() < 0.4: b = b[:rng.randrange(len(b) + 1)]
            if len(b) + 16 > budget_b: break
            budget_b -= len(b) + 16
            chunks.append(bytes(b))
        if not chunks: continue
        pref = np.array([rng.choice([-1, -1, 1, 2, 3, 4, 12]) for _ in chunks], dtype=np.int32)
        flg, dev, (data, runs) = T.k_small(sh, k, chunks, preferred=pref, flags=1 if client else 0, max_msgs=1024)
        sh.sh_destroy(k)
        if flg: continue                                              # over the compact block's capacity: the library takes the tile pipeline
        assert_same(dev, O.process_batch(cfg, data, runs), "fuzz seed %d" % (base_seed + seed))
        total += len(dev[1]); batches += 1
    return batches, seed, total


if __name__ == "__main__":
    b, sd, m = run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 7000000)
    print("fuzz ok: %d batches (%d seeds), %d messages, emulated k_small == oracle everywhere" % (b, sd, m))
