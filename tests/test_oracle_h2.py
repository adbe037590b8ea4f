"""CPU: the h2 server-side oracle against transcripts worked out by hand from the reference
(src/brpc/policy/http2_rpc_protocol.cpp) — what it must WriteAck and which requests it must hand on."""
import random

import numpy as np

import _oracle as O
import _h2traffic as T

SETTINGS_REPLY = bytes.fromhex("00000c040000000000" "000200000000" "000400040000" "000004080000000000" "000f0001")
ACK = bytes.fromhex("000000040100000000")


def hdrs(blob, m):
    return O.parse_header_records(blob[m["headers_off"]:m["headers_off"] + m["headers_len"]])


def test_preface_settings_ping_and_errors():
    c = O.H2Conn()
    err, cons, msgs, ctrl, blob, mfs, sws = c.consume(T.PREFACE[:10])
    assert (err, cons, ctrl) == (2, 0, b"")                               # NOT_ENOUGH_DATA, nothing popped
    err, cons, msgs, ctrl, blob, mfs, sws = c.consume(T.PREFACE + T.settings([(5, 32768), (4, 70000)]) + T.frame(6, 0, 0, b"12345678"))
    assert err == 2 and cons == 24 + 9 + 12 + 17
    # SETTINGS{ENABLE_PUSH=0, INITIAL_WINDOW_SIZE=256K} + WINDOW_UPDATE(1M - 65535) (:230-259), the ack (:905-911), the pong (:942-949)
    assert ctrl == SETTINGS_REPLY + ACK + bytes.fromhex("000008060100000000") + b"12345678"
    assert (mfs, sws) == (32768, 70000)
    # PRIORITY is not supported: GOAWAY(last_stream=-1, PROTOCOL_ERROR) and its 5 payload bytes are parsed as the next frame head
    err, cons, msgs, ctrl, blob, _, _ = c.consume(T.frame(2, 0, 1, bytes(5)))
    assert ctrl == bytes.fromhex("000008070000000000" "ffffffff" "00000001") and cons == 9 and err == 2
    c2 = O.H2Conn()
    assert c2.consume(b"GET / HTTP/1.1\r\n")[0] == 1                     # TRY_OTHERS
    c3 = O.H2Conn(); c3.consume(T.PREFACE)
    assert c3.consume(T.frame(0, 0, 1, bytes(20000))[:200])[0] == 5     # frame longer than max_frame_size: ABSOLUTELY_WRONG
    assert c3.consume(T.frame(11, 0, 1, b""))[0] == 5                   # unknown frame type
    # DATA on a stream nobody opened: RST_STREAM(STREAM_CLOSED)
    c4 = O.H2Conn(); c4.consume(T.PREFACE)
    err, cons, msgs, ctrl, blob, _, _ = c4.consume(T.frame(0, 1, 7, b"abc"))
    assert ctrl == bytes.fromhex("000004030000000007" "00000005") and cons == 12 and len(msgs) == 0


def test_unary_grpc_request_with_continuation_padding_and_trailers():
    rng = random.Random(5)
    enc = T.HpackEncoder(rng)
    c = O.H2Conn()
    c.consume(T.PREFACE + T.settings())
    fr = T.request_frames(rng, enc, 1, message=b"x" * 3000, split_headers=True, pad=True, priority=True, chunk=700, trailers=True,
                          extra=[(b"grpc-timeout", b"5S")])
    stream = b"".join(fr)
    err, cons, msgs, ctrl, blob, _, _ = c.consume(stream)
    assert err == 2 and cons == len(stream) and len(msgs) == 1 and ctrl == b""
    m = msgs[0]
    h = hdrs(blob, m)
    assert h[0] == (b":method", b"POST") and (b":path", b"/example.EchoService/Echo") in h and h[-1] == (b"x-trailer", b"done") and m["n_headers"] == len(h)
    assert m["stream_id"] == 1 and m["http_method"] == 3 and m["content_type"] == 2 and m["method_idx"] == 0
    assert m["flags"] == 1 | 2 | 8 and m["msg_len"] == 3003 and m["body_len"] == 3008
    assert bytes(blob[m["msg_off"]:m["msg_off"] + m["msg_len"]]) == b"\x0a\xb8\x17" + b"x" * 3000
    assert bytes(blob[m["path_off"]:m["path_off"] + m["path_len"]]) == b"/example.EchoService/Echo"
    # the second call on the connection hits the dynamic table (indexed :path / content-type)
    fr2 = T.request_frames(rng, enc, 3, message=b"again")
    err, cons, msgs, ctrl, blob, _, _ = c.consume(b"".join(fr2))
    assert len(msgs) == 1 and msgs[0]["method_idx"] == 0 and msgs[0]["stream_id"] == 3


def test_every_split_point_gives_the_same_transcript():
    rng = random.Random(11)
    stream = b"".join(T.connection_script(rng, n_calls=6, violations=0.3))
    whole = O.H2Conn()
    err, cons, msgs, ctrl, blob, _, _ = whole.consume(stream)
    ref = (ctrl, [(int(m["stream_id"]), bytes(blob[m["headers_off"]:m["headers_off"] + m["headers_len"]]),
                   bytes(blob[m["body_off"]:m["body_off"] + m["body_len"]]), int(m["flags"]), int(m["method_idx"])) for m in msgs])
    for step in (1, 7, 64, 1000):
        c = O.H2Conn(); buf = b""; ctrl2 = b""; got = []; fed = 0
        while fed < len(stream):
            buf += stream[fed:fed + step]; fed += step
            e, cons, msgs, ct, blob, _, _ = c.consume(buf)
            buf = buf[cons:]; ctrl2 += ct
            got += [(int(m["stream_id"]), bytes(blob[m["headers_off"]:m["headers_off"] + m["headers_len"]]),
                     bytes(blob[m["body_off"]:m["body_off"] + m["body_len"]]), int(m["flags"]), int(m["method_idx"])) for m in msgs]
            if e not in (2,):
                break
        assert (ctrl2, got) == ref, step


def test_grpc_response_bytes_first_and_later_calls():
    """H2UnsentResponse::AppendAndDestroySelf + PackH2Message: the first response on a connection carries literal
    (incrementally indexed) content-type / grpc-status fields, later ones are fully indexed (hpack.cpp:696-726)."""
    c = O.H2Conn()
    c.consume(T.PREFACE + T.settings())
    r1 = c.pack_response(1, b"\x0a\x05hello")
    assert r1 == (bytes.fromhex("000013010400000001") + b"\x88\x5f\x10application/grpc" +
                  bytes.fromhex("00000c000000000001") + b"\x00\x00\x00\x00\x07\x0a\x05hello" +
                  bytes.fromhex("00000f010500000001") + b"\x40\x0bgrpc-status\x010")
    r2 = c.pack_response(3, b"\x0a\x05hello")
    assert r2 == bytes.fromhex("000002010400000003" "88bf" "00000c000000000003") + b"\x00\x00\x00\x00\x07\x0a\x05hello" + bytes.fromhex("000001010500000003" "be")
    r3 = c.pack_response(5, b"", grpc_status=12, grpc_message=b"no%20method")
    assert r3.endswith(b"\x7e\x0212\x40\x0cgrpc-message\x0bno%20method") and bytes.fromhex("000005000000000005" "0000000000") in r3
    # a body longer than the peer's max_frame_size is cut into DATA frames; the last one does not end the stream (trailers do)
    big = c.pack_response(7, b"z" * 40000)
    assert big.count(bytes.fromhex("004000000000000007")) == 2 and bytes.fromhex("001c45000000000007") in big
    # peer announced header_table_size 0: never-indexed literals (index_policy, :1718-1720)
    c2 = O.H2Conn(); c2.consume(T.PREFACE + T.settings([(1, 0)]))
    assert c2.pack_response(1, b"", grpc=False, content_type=b"text/plain") == bytes.fromhex("000012010500000001") + b"\x18\x03200\x1f\x10\x0atext/plain"
    # connection window exhausted: RST_STREAM(FLOW_CONTROL_ERROR) instead of the response (:1706-1712)
    c3 = O.H2Conn(); c3.consume(T.PREFACE + T.settings())
    assert c3.pack_response(1, b"y" * 70000) == bytes.fromhex("000004030000000001" "00000003")


def test_settings_case_of_the_reference_unit_test():
    """test/brpc_http_rpc_protocol_unittest.cpp:1382-1412 (HttpTest.http2_settings): a SETTINGS frame serialized from
    H2Settings{header_table_size 8192, max_concurrent_streams 1024, stream_window_size 2^29-1} is answered with exactly one
    9-byte SETTINGS frame, flags ACK, stream 0, and the remote settings take the values.  SerializeH2Settings (:213-250)
    writes the non-default fields in id order; enable_push=false differs from DEFAULT_ENABLE_PUSH, so id 2 is present."""
    frame = T.settings([(1, 8192), (2, 0), (3, 1024), (4, (1 << 29) - 1)])
    assert len(frame) == 9 + 24
    c = O.H2Conn()
    c.consume(T.PREFACE)                                   # the unit test forces H2_CONNECTION_READY; here the preface does
    err, cons, msgs, ctrl, blob, mfs, sws = c.consume(frame)
    assert (err, cons, len(msgs)) == (2, len(frame), 0)
    assert ctrl == ACK and ctrl[3] == 4 and ctrl[4] == 1 and ctrl[5:9] == b"\x00\x00\x00\x00"
    assert sws == (1 << 29) - 1 and mfs == 16384
    # max_frame_size outside [DEFAULT_MAX_FRAME_SIZE, MAX_OF_MAX_FRAME_SIZE] is invalid (ParseH2Settings :193-199, the bounds
    # HttpTest.http2_invalid_settings checks on the server options): connection error, GOAWAY(PROTOCOL_ERROR)
    for bad in (16383, 16777216):
        c2 = O.H2Conn(); c2.consume(T.PREFACE)
        err, cons, msgs, ctrl, *_ = c2.consume(T.settings([(5, bad)]))
        assert ctrl == bytes.fromhex("000008070000000000" "ffffffff" "00000001")
