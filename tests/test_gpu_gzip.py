"""GPU: gzip / zlib request and response bodies (policy/gzip_compress.cpp:75-89 = GzipInputStream + parse) through the whole path
against the oracle, whose inflate is pinned to the system zlib driven like GzipInputStream (tests/test_oracle_gzip.py).  Streams of
every block type and level, with CRC32C over the compressed body, attachments, concatenated members, corruptions, truncations,
outputs beyond the 64 KiB call buffer, and bodies over the device's size limit (left to the host)."""
import random
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import _oracle as O  # noqa: E402
from _compare import assert_same  # noqa: E402
from _traffic import SEED, echo_pb, raw_request_frame, raw_response_frame, rnd62, echo_frame  # noqa: E402

GZIP, ZLIB = 2, 3


def _deflate(data, fmt, level=6, strategy=0):
    c = zlib.compressobj(level, zlib.DEFLATED, 31 if fmt == GZIP else 15, 8, strategy)
    return c.compress(data) + c.flush()


def _crc_be(body):
    c = O.crc32c(body)
    masked = (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff
    return masked.to_bytes(4, "big")


def _bodies(rng, n_big=True):
    """(compress_type, body bytes) pairs around EchoRequest{message}"""
    out = []
    sizes = [0, 1, 10, 100, 1000, 5000, 40000] + ([70000, 200000] if n_big else [])
    for n in sizes:
        for msg in (rnd62(rng, n), b"r" * n, (rnd62(rng, 61) * (n // 61 + 1))[:n]):
            pb = echo_pb(msg)
            for fmt in (GZIP, ZLIB):
                level, strategy = rng.choice([(0, 0), (1, 0), (6, 0), (9, 0), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)])
                out.append((fmt, _deflate(pb, fmt, level, strategy)))
    return out


def _corrupt(rng, body):
    b = bytearray(body); c = rng.random()
    if not b:
        return bytes(b)
    if c < 0.3:
        del b[rng.randrange(len(b)):]
    elif c < 0.75:
        for _ in range(rng.randrange(1, 3)):
            b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
    elif c < 0.9:
        b += bytes(rng.randrange(256) for _ in range(rng.randrange(1, 30)))
    else:
        b[-1 - rng.randrange(min(8, len(b)))] ^= 0x21                 # the trailer checks
    return bytes(b)


def _streams(rng, frame_of, n_streams=48):
    pool = _bodies(rng)
    streams = []; cid = 1
    for s in range(n_streams):
        fr = []
        for _ in range(rng.randrange(1, 8)):
            fmt, body = rng.choice(pool)
            k = rng.random()
            if k < 0.25:
                body = _corrupt(rng, body)
            elif k < 0.32:                                           # a second member behind the first (multiple fields: the last message wins)
                body = body + _deflate(echo_pb(rnd62(rng, rng.choice([3, 300]))), fmt)
            elif k < 0.36:
                fmt = GZIP if fmt == ZLIB else ZLIB                  # the other wrapper
            att = rng.choice([b"", b"", rnd62(rng, 37)])
            with_crc = rng.random() < 0.3
            cks = _crc_be(body) if with_crc else None
            if with_crc and rng.random() < 0.1:
                cks = bytes([cks[0] ^ 1]) + cks[1:]
            fr.append(frame_of(body, cid, compress_type=fmt, checksum_value=cks, checksum_type=1 if with_crc else 0, attachment=att))
            cid += 1
            if rng.random() < 0.3:
                fr.append(echo_frame(rng, cid, rnd62(rng, rng.choice([5, 500])))); cid += 1       # plain traffic in between
        streams.append(b"".join(fr))
    return streams


@pytest.mark.parametrize("tile", [512, 8192])
def test_gzip_zlib_requests_through_the_path(tile):
    import brpc_b200
    rng = random.Random(SEED + 31)
    ms = [dict(brpc_b200.ECHO_METHOD, response_checksum_type=rng.choice([0, 1]))]
    ctx = brpc_b200.Context(device=0, max_batch_bytes=64 << 20, max_msgs=1 << 16, max_runs=4096, max_resp_bytes=256 << 20, tile_bytes=tile, methods=ms)
    cfg = O.make_config(methods=ms)
    streams = _streams(rng, raw_request_frame)
    data, runs = brpc_b200.make_runs(streams)
    dev = ctx.process_batch(data, runs)
    orc = O.process_batch(cfg, data, runs, resp_cap=256 << 20)
    assert_same(dev, orc, "tile=%d" % tile)
    st = dev[1]["status"]; ct = dev[1]["compress_type"]
    assert ((st == 0) & (ct >= 2)).sum() > 60 and ((st == 1) & (ct >= 2)).sum() > 15


def test_gzip_requests_on_the_latency_path_and_with_snappy_replies():
    import brpc_b200
    rng = random.Random(SEED + 32)
    for resp_compress in (0, 1):
        ms = [dict(brpc_b200.ECHO_METHOD, response_compress_type=resp_compress)]
        ctx = brpc_b200.Context(device=0, max_batch_bytes=8 << 20, max_msgs=1 << 14, max_runs=256, methods=ms)
        cfg = O.make_config(methods=ms)
        for trial in range(6):
            streams = []
            for s in range(rng.randrange(1, 12)):
                msg = rnd62(rng, rng.choice([0, 7, 300, 3000]))
                fmt = rng.choice([GZIP, ZLIB]); body = _deflate(echo_pb(msg), fmt, rng.choice([0, 1, 6]))
                if rng.random() < 0.2:
                    body = _corrupt(rng, body)
                streams.append(raw_request_frame(body, 100 + s, compress_type=fmt, attachment=rng.choice([b"", b"att"])))
            data, runs = brpc_b200.make_runs(streams)
            assert data.nbytes <= 128 << 10                      # the k_small / ring batch size
            assert_same(ctx.process_batch(data, runs), O.process_batch(cfg, data, runs), "small trial %d compress %d" % (trial, resp_compress))


def test_gzip_zlib_responses_on_client_sockets():
    import brpc_b200
    rng = random.Random(SEED + 33)
    ctx = brpc_b200.Context(device=0, max_batch_bytes=64 << 20, max_msgs=1 << 16, max_runs=4096, max_resp_bytes=256 << 20)
    cfg = O.make_config()
    streams = _streams(rng, raw_response_frame, n_streams=32)
    data, runs = brpc_b200.make_runs(streams)
    runs["flags"] = 1                                                # B2_RUN_CLIENT
    dev = ctx.process_batch(data, runs)
    orc = O.process_batch(cfg, data, runs, resp_cap=256 << 20)
    assert_same(dev, orc, "client")
    st = dev[1]["status"]
    assert (st == 8).sum() > 40                                      # B2_MSG_RESPONSE_UNZ


def test_bodies_beyond_the_device_limit_go_to_the_host():
    import brpc_b200
    rng = random.Random(SEED + 34)
    ctx = brpc_b200.Context(device=0, max_batch_bytes=64 << 20, max_msgs=1 << 12, max_runs=64, max_resp_bytes=128 << 20)
    cfg = O.make_config()
    bomb = _deflate(echo_pb(b"\x00" * (8 << 20)), GZIP)                       # 8 MiB out of ~8 KiB
    just_fits = _deflate(echo_pb(b"q" * ((1 << 20) - 8)), ZLIB)
    just_over = _deflate(echo_pb(b"q" * ((1 << 20) - 3)), ZLIB)
    long_in = _deflate(echo_pb(bytes(rng.getrandbits(8) for _ in range((1 << 20) + 100))), GZIP, 0)
    frames = [raw_request_frame(b, i + 1, compress_type=GZIP if b[:1] == b"\x1f" else ZLIB) for i, b in enumerate([bomb, just_fits, just_over, long_in])]
    data, runs = brpc_b200.make_runs([b"".join(frames)])
    dev = ctx.process_batch(data, runs)
    orc = O.process_batch(cfg, data, runs, resp_cap=128 << 20)
    assert_same(dev, orc, "limits")
    assert list(dev[1]["status"]) == [6, 0, 6, 6], list(dev[1]["status"])     # B2_MSG_UNSUPPORTED = 6
