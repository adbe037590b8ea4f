"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle, bit-exact.

Covers the cases the reference's own tests exercise for this path:
frames split at every offset and several frames per read
(test/brpc_input_messenger_unittest.cpp:95-221), the content/compress/checksum
matrix of test/brpc_server_unittest.cpp:1690-1880 (PB x none x {none,crc32c}),
fuzz-style corruptions (test/fuzzing/fuzz_baidu_rpc.cpp), empty and maximum inputs.
"""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import _oracle as O  # noqa: E402
from _compare import assert_same  # noqa: E402
from _traffic import SEED, echo_frame, mixed_frames, patch_meta, rnd62, split_runs  # noqa: E402


@pytest.fixture(scope="module")
def b2():
    import brpc_b200
    return brpc_b200


def make_ctx(b2, tile_bytes=0, max_body_size=0, methods=None, identity=None, max_batch=64 << 20, max_msgs=1 << 20):
    methods = [b2.ECHO_METHOD] if methods is None else methods
    return b2.Context(device=0, max_batch_bytes=max_batch, max_msgs=max_msgs, max_runs=8192, tile_bytes=tile_bytes,
                      max_body_size=max_body_size, methods=methods, server_identity=identity)


def run_both(b2, ctx, chunks, cfg=None, preferred=None, what=""):
    data, runs = b2.make_runs(chunks)
    if preferred is not None:
        runs["preferred_proto"] = preferred
    dev = ctx.process_batch(data, runs)
    cfg = cfg or O.make_config()
    orc = O.process_batch(cfg, data, runs)
    assert_same(dev, orc, what)
    return dev, orc


def test_single_socket_whole_frames(b2):
    ctx = make_ctx(b2)
    rng = random.Random(SEED)
    frames = [echo_frame(rng, i, b"hello world") for i in range(10)]
    dev, _ = run_both(b2, ctx, [b"".join(frames)])
    rs, msgs, resp, info = dev
    assert rs["n_msgs"][0] == 10 and rs["consumed"][0] == sum(map(len, frames)) and rs["parse_error"][0] == 2
    assert np.all(msgs["status"] == 0) and info["n_launches"] >= 1


def test_empty_inputs(b2):
    ctx = make_ctx(b2)
    run_both(b2, ctx, [b""])
    run_both(b2, ctx, [b"", b"", b""])
    rng = random.Random(1)
    run_both(b2, ctx, [b"", echo_frame(rng, 1, b"x"), b""])
    data, runs = b2.make_runs([])
    rs, msgs, resp, _ = ctx.process_batch(data, runs[:0])
    assert len(rs) == 0 and len(msgs) == 0


def test_truncated_header_every_length(b2):
    ctx = make_ctx(b2)
    rng = random.Random(2)
    f = echo_frame(rng, 3, b"abc")
    for pref in (-1, 1, 2):
        chunks = [f[:n] for n in range(0, 14)] + [f + f[:n] for n in range(0, 14)] + [b"P", b"PR", b"S", b"ST", b"STRM", b"X", b"PRPX", b"PXPC"]
        run_both(b2, ctx, chunks, preferred=pref, what="pref=%d" % pref)


def test_split_at_every_offset(b2, monkeypatch):
    monkeypatch.setenv("B2_SMALL", "off")
    """One frame stream cut at every byte offset (brpc_input_messenger_unittest.cpp chunking discipline)."""
    ctx = make_ctx(b2, tile_bytes=512)
    rng = random.Random(3)
    stream = b"".join(echo_frame(rng, i, rnd62(rng, rng.choice([0, 5, 40, 300]))) for i in range(6))
    chunks = [stream[:n] for n in range(0, len(stream) + 1)]
    run_both(b2, ctx, chunks)


def test_mixed_matrix_many_sockets(b2, monkeypatch):
    for tile in (512, 2048, 8192):
        # small batches normally take the one-launch k_small path; B2_SMALL=off sends them through the
        # tile pipeline instead, so both implementations face the same traffic
        monkeypatch.setenv("B2_SMALL", "off" if tile == 2048 else "on")
        ctx = make_ctx(b2, tile_bytes=tile, identity=b"10.1.2.3:8000")
        rng = random.Random(SEED + tile)
        streams = [mixed_frames(rng, rng.randrange(1, 40)) for _ in range(200)]
        chunks = split_runs(rng, streams)
        cfg = O.make_config(server_identity=b"10.1.2.3:8000")
        dev, _ = run_both(b2, ctx, chunks, cfg=cfg, what="tile=%d" % tile)
        assert len(set(dev[1]["status"].tolist())) >= 5


def test_method_options(b2):
    """echo_attachment off, response checksum on (example/echo_c++/server.cpp:31,72-82), host-handled method."""
    rng = random.Random(5)
    ms = [dict(b2.ECHO_METHOD, echo_attachment=0, response_checksum_type=1),
          dict(service_full_name=b"example.Other", service_name=b"Other", method_name=b"Call",
               request_type_name=b"example.OtherRequest", handler=0, echo_attachment=0,
               response_checksum_type=0, response_compress_type=0)]
    ctx = make_ctx(b2, methods=ms)
    cfg = O.make_config(methods=ms)
    streams = []
    for s in range(50):
        fr = mixed_frames(rng, 20)
        fr += [echo_frame(rng, 7, b"zzz", service=b"example.Other", method=b"Call"),
               echo_frame(rng, 8, b"zzz", service=b"Other", method=b"Call")]
        rng.shuffle(fr)
        streams.append(fr)
    run_both(b2, ctx, split_runs(rng, streams), cfg=cfg)


def test_large_frames_span_tiles(b2):
    ctx = make_ctx(b2, tile_bytes=1024)
    rng = random.Random(6)
    streams = [[echo_frame(rng, i, rnd62(rng, rng.choice([3000, 10000, 70000, 5])), attachment=rnd62(rng, rng.choice([0, 0, 5000])))
                for i in range(12)] for _ in range(20)]
    run_both(b2, ctx, split_runs(rng, streams))


def test_payload_full_of_fake_frames(b2, monkeypatch):
    monkeypatch.setenv("B2_SMALL", "off")
    """Payloads that contain valid-looking baidu_std frames: the speculative tile scan must never
    change the result (SURVEY §7 'hard parts')."""
    rng = random.Random(7)
    for tile in (512, 1024, 4096):
        ctx = make_ctx(b2, tile_bytes=tile)
        streams = []
        for s in range(40):
            fr = []
            for i in range(10):
                inner = b"".join(echo_frame(rng, 100 + k, rnd62(rng, rng.choice([10, 200, 700]))) for k in range(rng.randrange(1, 8)))
                junk = rnd62(rng, rng.randrange(0, 50))
                fr.append(echo_frame(rng, i, junk + inner + b"PRPC" + rnd62(rng, 3) + b"STRM\x00\x00\x00\x05\x00\x00\x00\x01" + inner[:rng.randrange(0, 60)]))
            streams.append(fr)
        run_both(b2, ctx, split_runs(rng, streams), what="tile=%d" % tile)


def test_corrupted_streams(b2, monkeypatch):
    monkeypatch.setenv("B2_SMALL", "off")           # the speculative tile pipeline ...
    ctx = make_ctx(b2, tile_bytes=512, max_body_size=1 << 20)
    monkeypatch.setenv("B2_SMALL", "on")            # ... and the one-launch k_small path, same streams
    ctx_small = make_ctx(b2, tile_bytes=512, max_body_size=1 << 20)
    cfg = O.make_config(max_body_size=1 << 20)
    rng = random.Random(8)
    chunks = []
    for s in range(300):
        fr = [echo_frame(rng, i, rnd62(rng, rng.choice([0, 30, 500]))) for i in range(8)]
        k = rng.randrange(len(fr)); f = bytearray(fr[k]); c = rng.random()
        if c < 0.2:   f[rng.randrange(4)] ^= 0x20                                  # bad magic -> TRY_OTHERS, socket closed
        elif c < 0.4: f[4:8] = (2 << 20).to_bytes(4, "big")                        # body_size > max_body_size -> TOO_BIG_DATA
        elif c < 0.6: f[8:12] = (int.from_bytes(f[4:8], "big") + 1 + rng.randrange(50)).to_bytes(4, "big")  # meta > body: pop + TRY_OTHERS
        elif c < 0.7: f[0:4] = b"RDMA"
        elif c < 0.8: f = bytearray(b"STRM") + f[4:]                               # baidu body under a stream magic
        else:         f[12 + rng.randrange(20)] ^= 0xff
        fr[k] = bytes(f)
        if rng.random() < 0.3:
            fr.insert(rng.randrange(len(fr)), O.pack_stream_frame(5, 6, 3, True, b"data"))
        chunks.append(b"".join(fr))
    for pref in (-1, 1, 2):
        run_both(b2, ctx, chunks, cfg=cfg, preferred=pref, what="pref=%d" % pref)
        for k in range(0, len(chunks), 40):         # k_small takes batches of <= 128 KB
            run_both(b2, ctx_small, chunks[k:k + 40], cfg=cfg, preferred=pref, what="small pref=%d k=%d" % (pref, k))


def test_meta_edge_encodings(b2):
    """Unknown fields, groups, repeated occurrences, over-long varints inside RpcMeta."""
    ctx = make_ctx(b2)
    rng = random.Random(9)
    base = echo_frame(rng, 1, b"payload-bytes")
    muts = [lambda m: m + b"\x98\x06\x01",                       # unknown varint field 99
            lambda m: b"\xa3\x06\x08\x01\xa4\x06" + m,           # unknown group 100 {1: 1}
            lambda m: b"\xa3\x06\xa3\x06\xa4\x06\xa4\x06" + m,   # nested groups
            lambda m: b"\xa3\x06" + m,                           # unterminated group -> parse fails
            lambda m: m + b"\x20\x81\x80\x80\x80\x80\x80\x80\x80\x80\x01",  # 10-byte varint correlation_id (last wins)
            lambda m: m + b"\x20" + b"\x80" * 10 + b"\x01",      # 11-byte varint -> fails
            lambda m: m + b"\x50\x07",                           # content_type = 7: closed enum, stays PB
            lambda m: m + b"\x50\x01",                           # content_type = JSON -> unsupported
            lambda m: m + b"\x0a\x00",                           # second (empty) request sub-message merges
            lambda m: m + b"\x32\x02\x08\x01",                   # chunk_info missing chunk_id -> fails
            lambda m: m + b"\x32\x04\x08\x01\x10\x02",           # chunk_info ok
            lambda m: m + b"\x42\x00",                           # stream_settings missing stream_id -> fails
            lambda m: m + b"\x42\x06\x08\x05\x22\x02\x01\x02",   # stream_settings packed extra ids
            lambda m: m + b"\x4a\x06\x0a\x01k\x12\x01v",         # user_fields entry
            lambda m: m + b"\x3a\x03abc",                        # authentication_data
            lambda m: m + b"\x00",                               # tag 0 -> fails
            lambda m: m + b"\x0c",                               # stray END_GROUP -> fails
            lambda m: m + b"\x0e",                               # wire type 6 -> fails
            lambda m: m + b"\x2d\x01\x02\x03\x04",               # attachment_size with wire type 5: unknown field, skipped
            lambda m: m + b"\x28\xff\xff\xff\xff\x0f",           # attachment_size = -1
            lambda m: m + b"\x28\x05",                           # attachment_size = 5
            lambda m: m + b"\x28\x90\x4e",                       # attachment_size > request -> EREQUEST
            lambda m: m[:-2] + b"\x62\x02\xab\xcd",              # 2-byte checksum_value travels back
            lambda m: m + b"\x62\x85\x01" + b"Q" * 133,          # long checksum_value (2-byte length varint)
            ]
    chunks = [patch_meta(base, f) for f in muts]
    chunks += [patch_meta(base, lambda m: b"")]                   # empty meta: no request -> ENOMETHOD ""/""
    run_both(b2, ctx, chunks)
    run_both(b2, ctx, [b"".join(chunks[:3] + chunks[4:5])])


def test_echo_body_edge_encodings(b2):
    ctx = make_ctx(b2)
    def frame(body, att=b""):
        meta = bytes.fromhex("0a1e0a136578616d706c652e4563686f5365727669636512044563686f18b960180020c0c48780705000580062 00".replace(" ", ""))
        if att:
            meta += b"\x28" + bytes([len(att)])
        return b"PRPC" + (len(meta) + len(body) + len(att)).to_bytes(4, "big") + len(meta).to_bytes(4, "big") + meta + body + att
    bodies = [b"\x0a\x03abc", b"", b"\x0a\x00", b"\x0a\x03abc\x0a\x02zz",            # last occurrence wins
              b"\x10\x05\x0a\x03abc", b"\x0a\x03abc\x10\x05",                       # unknown field before / after
              b"\x0a\x05abc", b"\x08\x01", b"\x0a\x03abc\x00", b"\x0b\x0a\x01x\x0c\x0a\x01y"]
    chunks = [frame(b) for b in bodies] + [frame(b, b"ATTACH") for b in bodies]
    run_both(b2, ctx, chunks)


def test_baseline_1k_workload_against_oracle(b2):
    """BASELINE configs[1] shape at a size the oracle handles in seconds: 64 sockets, 1 KB payload,
    runs cut mid-frame, default tile."""
    from brpc_b200 import press
    ctx = make_ctx(b2, max_batch=80 << 20, max_msgs=1 << 17)
    for kind, cks, att in [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 0, 64)]:
        sp = press.spec(payload_bytes=1024, payload_kind=kind, checksum_type=cks, attachment_bytes=att)
        n_sock, run_bytes = 64, (1 << 20) - 123
        data = np.zeros(n_sock * (1 << 20), dtype=np.uint8)
        runs, n_full = press.fill_batch(sp, data, n_sock, run_bytes)
        dev = ctx.process_batch(data, runs)
        orc = O.process_batch(O.make_config(), data, runs)
        assert_same(dev, orc, "kind=%d cks=%d att=%d" % (kind, cks, att))
        assert len(dev[1]) == n_full and np.all(dev[1]["status"] == 0)


def test_full_size_properties(b2):
    """BASELINE full size (256 MiB batch): size-independent properties instead of a full oracle diff:
    message count, per-message length law, echo round trip of a sample, and responses re-parse."""
    from brpc_b200 import press
    sp = press.spec(payload_bytes=1024)
    n_sock, run_bytes = 64, 4 << 20
    ctx = make_ctx(b2, max_batch=n_sock * run_bytes + (1 << 20), max_msgs=1 << 19)
    data = np.zeros(n_sock * run_bytes, dtype=np.uint8)
    runs, n_full = press.fill_batch(sp, data, n_sock, run_bytes)
    rs, msgs, resp, _ = ctx.process_batch(data, runs)
    assert len(msgs) == n_full == int(rs["n_msgs"].sum())
    assert np.all(msgs["status"] == 0) and np.all(rs["parse_error"] == 2)
    # request frame = 12 + meta + 1027; response frame = 12 + meta' + 1027: both laws at once
    assert np.all(msgs["body_size"] - msgs["meta_size"] == 1027)
    cid_len = np.ceil(np.log2(msgs["correlation_id"].astype(np.float64) + 1) / 7).astype(np.int64)
    assert np.all(msgs["resp_len"] == 12 + (4 + 2 + 1 + cid_len + 2 + 2 + 2) + 1027)
    # the oracle on a 1/64 sample (one whole run)
    sub = runs[5:6].copy(); off = int(sub["offset"][0]); sub["offset"] = 0
    o_rs, o_msgs, o_resp = O.process_batch(O.make_config(), data[off:off + run_bytes], sub)
    a, n = int(rs["first_msg"][5]), int(rs["n_msgs"][5])
    assert n == len(o_msgs)
    from _compare import gather
    assert np.array_equal(gather(resp, msgs["resp_off"][a:a + n], msgs["resp_len"][a:a + n]), o_resp)
    # every response payload is 1024 x 'r' at its tail
    tails = gather(resp, msgs["resp_off"] + msgs["resp_len"] - 1024, np.full(len(msgs), 1024, np.uint32))
    assert np.all(tails == ord("r"))


def test_messenger_carries_partial_frames(b2):
    """Host mirror: bytes fed in arbitrary chunks over several polls give the same messages as one shot."""
    ctx = make_ctx(b2, tile_bytes=512)
    rng = random.Random(11)
    streams = {s: b"".join(mixed_frames(rng, 30)) for s in range(8)}
    m = b2.GpuInputMessenger(ctx)
    for s in streams:
        m.add_socket(s)
    pos = {s: 0 for s in streams}
    got = {s: [] for s in streams}
    while any(pos[s] < len(streams[s]) for s in streams):
        for s in streams:
            n = rng.randrange(0, 3000)
            m.feed(s, streams[s][pos[s]:pos[s] + n]); pos[s] += n
        for sid, d, body in m.poll():
            got[sid].append((int(d["status"]), int(d["correlation_id"]), body))
    for s in streams:
        chunks = [streams[s]]
        data, runs = b2.make_runs(chunks)
        o_rs, o_msgs, o_resp = O.process_batch(O.make_config(), data, runs)
        if o_rs["parse_error"][0] != 2:
            continue      # a corrupted frame closes the socket; how much was cut before depends on the read pattern
        exp = [(int(x["status"]), int(x["correlation_id"]), bytes(o_resp[int(x["resp_off"]):int(x["resp_off"]) + int(x["resp_len"])])) for x in o_msgs]
        assert got[s] == exp


def test_small_batch_overflow_redo(b2):
    """> 1024 tiny messages in < 128 KB: the compact latency block overflows and the batch is
    transparently redone on the normal path; results identical to the oracle."""
    ctx = make_ctx(b2)
    rng = random.Random(12)
    frames = [echo_frame(rng, i, b"") for i in range(1500)]
    assert sum(map(len, frames)) < 128 << 10
    dev, _ = run_both(b2, ctx, [b"".join(frames[:700]), b"".join(frames[700:])])
    assert len(dev[1]) == 1500
    # and right at the edge
    run_both(b2, ctx, [b"".join(frames[:1024])])
    run_both(b2, ctx, [b"".join(frames[:1025])])


def test_client_side_responses(b2):
    """Client-side sockets (CreatedByConnect): CutInputMessage's baidu_std<->streaming fallback rules and
    ProcessRpcResponse.  The response streams are what the oracle's server writes for mixed requests
    (OK / error replies, crc32c and snappy responses), re-read as a client would."""
    rng = random.Random(13)
    for opts in (dict(), dict(response_checksum_type=1), dict(response_compress_type=1), dict(response_checksum_type=1, response_compress_type=1)):
        ms = [dict(b2.ECHO_METHOD, **opts)]
        srv = O.make_config(methods=ms, server_identity=b"127.0.0.1:8002")
        streams = [mixed_frames(rng, 25) for _ in range(40)]
        data, runs = b2.make_runs(split_runs(rng, streams, cut_tail=False))
        rs, msgs, resp = O.process_batch(srv, data, runs)
        # per-socket response streams; sprinkle stream frames and a corrupted response
        client_chunks = []
        for r in range(len(runs)):
            m = msgs[int(rs["first_msg"][r]):int(rs["first_msg"][r]) + int(rs["n_msgs"][r])]
            parts = [bytes(resp[int(x["resp_off"]):int(x["resp_off"]) + int(x["resp_len"])]) for x in m if x["resp_len"] > 0 and x["status"] in (0, 1)]
            if r % 5 == 0 and parts:
                parts.insert(len(parts) // 2, O.pack_stream_frame(77, 78, 3, False, b"stream-data"))
            if r % 7 == 0 and parts:
                p = bytearray(parts[-1]); p[-1] ^= 0x33; parts[-1] = bytes(p)
            if r % 11 == 0 and parts:
                p = bytearray(parts[0]); p[0] = ord("X"); parts[0] = bytes(p)      # bad magic on a client socket
            b = b"".join(parts)
            if r % 3 == 0 and len(b) > 20:
                b = b[:len(b) - rng.randrange(1, 20)]
            client_chunks.append(b)
        ctx = make_ctx(b2, tile_bytes=1024)
        cdata, cruns = b2.make_runs(client_chunks)
        cruns["flags"] = 1
        for pref in (1, 2, -1):
            cruns["preferred_proto"] = pref
            dev = ctx.process_batch(cdata, cruns)
            orc = O.process_batch(O.make_config(), cdata, cruns)
            assert_same(dev, orc, "opts=%s pref=%d" % (opts, pref))
            st = set(dev[1]["status"].tolist())
            assert 7 in st or 8 in st
            if opts.get("response_compress_type"):
                assert 8 in st
        # the in-place message of an OK response equals the request's message
        ok = dev[1][(dev[1]["status"] == 7) & (dev[1]["error_code"] == 0) & (dev[1]["resp_len"] > 0)]
        assert len(ok) > 10 or opts.get("response_compress_type")


def test_dense_and_sparse_tiles(b2):
    """Tiles holding far more frames than the speculative walk keeps (tiny requests) next to tiles holding a few
    (1 KB requests) and tiles with none (64 KB requests): k_frame_table's copy path and its re-walk path together."""
    rng = random.Random(SEED + 99)
    for tile in (0, 2048, 8192):
        ctx = make_ctx(b2, tile_bytes=tile)
        streams = []
        for s in range(24):
            fr = []
            while sum(len(f) for f in fr) < 300_000:
                kind = rng.random()
                if kind < 0.5:
                    fr += [echo_frame(rng, len(fr) + k, b"r" * rng.choice([0, 1, 16])) for k in range(rng.randrange(20, 200))]
                elif kind < 0.9:
                    fr += [echo_frame(rng, len(fr) + k, rnd62(rng, 1024)) for k in range(rng.randrange(1, 30))]
                else:
                    fr.append(echo_frame(rng, len(fr), rnd62(rng, 65536)))
            streams.append(b"".join(fr)[:rng.randrange(250_000, 300_000)])
        dev, _ = run_both(b2, ctx, streams, what="dense/sparse tile=%d" % tile)
        assert len(dev[1]) > 20000
        # the context now knows the average request is small: the second batch takes the warp-staged (dense) tile walk
        dev2, _ = run_both(b2, ctx, streams[::-1], what="dense/sparse tile=%d, second batch" % tile)
        assert len(dev2[1]) == len(dev[1])
        tiny = [b"".join(echo_frame(rng, k, b"r" * rng.choice([0, 1, 16, 64])) for k in range(3000))[:rng.randrange(150_000, 200_000)] for _ in range(16)]
        run_both(b2, ctx, tiny, what="tiny requests tile=%d" % tile)
