"""oracle/b2_oracle_gzip.c (restated inflate + GzipInputStream) against the system zlib driven like GzipInputStream
(tests/_gzipstream.py): valid streams of every block type, and corrupted / truncated / concatenated ones — what counts is the
exact byte string the protobuf parser would get to see (gzip_compress.cpp:75-89)."""
import random
import struct
import zlib

import pytest

import _gzipstream as G
import _oracle as O


def _payloads(rng):
    text = b" ".join(rng.choice([b"echo", b"brpc", b"socket", b"message", b"iobuf", b"attachment", b"x" * 40]) for _ in range(3000))
    return [b"", b"a", b"hello world", bytes(rng.getrandbits(8) for _ in range(5000)), text, b"\x00" * 70000,
            bytes(rng.getrandbits(8) for _ in range(70000)), text * 12, (b"ab" * 40000) + bytes(rng.getrandbits(8) for _ in range(3000))]


def _deflate(data, fmt, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8):
    c = zlib.compressobj(level, zlib.DEFLATED, 31 if fmt == G.GZIP else 15, mem, strategy)
    return c.compress(data) + c.flush()


def _check(stream, fmt):
    want = G.gzip_input_stream(stream, fmt)
    got = O.gzip_input_stream(stream, fmt)
    assert got == want, (len(stream), len(got), len(want))
    return want


@pytest.mark.parametrize("fmt", [G.GZIP, G.ZLIB])
def test_valid_streams_all_block_types(fmt):
    rng = random.Random(11)
    for data in _payloads(rng):
        for level, strategy in [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE), (6, zlib.Z_FILTERED)]:
            s = _deflate(data, fmt, level, strategy)
            assert _check(s, fmt) == data
    # sync-flushed pieces (empty stored blocks between the blocks) and small windows
    c = zlib.compressobj(6, zlib.DEFLATED, 31 if fmt == G.GZIP else 9)
    parts = [c.compress(b"part one " * 50) + c.flush(zlib.Z_SYNC_FLUSH), c.compress(b"part two " * 70) + c.flush(zlib.Z_FULL_FLUSH), c.compress(b"tail") + c.flush()]
    assert _check(b"".join(parts), fmt) == b"part one " * 50 + b"part two " * 70 + b"tail"


def _gzip_with_header(data, extra=None, name=None, comment=None, hcrc=False, flags_extra=0, method=8):
    flg = (4 if extra is not None else 0) | (8 if name is not None else 0) | (16 if comment is not None else 0) | (2 if hcrc else 0) | flags_extra
    h = bytes([0x1f, 0x8b, method, flg]) + struct.pack("<IBB", 12345, 0, 3)
    if extra is not None:
        h += struct.pack("<H", len(extra)) + extra
    if name is not None:
        h += name + b"\0"
    if comment is not None:
        h += comment + b"\0"
    if hcrc:
        h += struct.pack("<H", zlib.crc32(h) & 0xffff)
    raw = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = raw.compress(data) + raw.flush()
    return h + body + struct.pack("<II", zlib.crc32(data), len(data) & 0xffffffff)


def test_gzip_header_fields_and_trailer_checks():
    data = b"header field test " * 100
    for kw in [dict(), dict(extra=b"\x01\x02abcd"), dict(name=b"file.bin"), dict(comment=b"a comment"), dict(hcrc=True),
               dict(extra=b"", name=b"", comment=b"", hcrc=True), dict(extra=b"x" * 300, name=b"n" * 100, comment=b"c" * 50, hcrc=True)]:
        s = _gzip_with_header(data, **kw)
        assert _check(s, G.GZIP) == data
        for cut in range(0, min(len(s), 480)):                       # truncated inside every header field
            _check(s[:cut], G.GZIP)
        for cut in range(len(s) - 12, len(s)):                       # ... and inside the trailer
            _check(s[:cut], G.GZIP)
    bad = bytearray(_gzip_with_header(data, hcrc=True)); bad[10] ^= 1                  # header crc mismatch
    assert _check(bytes(bad), G.GZIP) == b""
    assert _check(_gzip_with_header(data, flags_extra=0x20), G.GZIP) == b""              # reserved flag
    assert _check(_gzip_with_header(data, method=7), G.GZIP) == b""
    s = bytearray(_gzip_with_header(data)); s[-5] ^= 0x10                               # data check
    assert _check(bytes(s), G.GZIP) == b""
    s = bytearray(_gzip_with_header(data)); s[-1] ^= 0x10                               # length check
    assert _check(bytes(s), G.GZIP) == b""
    # a zlib stream offered as gzip and the other way round; FDICT; window size 8 + 8 > 15
    assert _check(_deflate(data, G.ZLIB), G.GZIP) == b""
    assert _check(_deflate(data, G.GZIP), G.ZLIB) == b""
    z = bytearray(_deflate(data, G.ZLIB)); z[1] |= 0x20; z[1] = (z[1] & 0xe0) | ((31 - ((z[0] << 8 | (z[1] & 0xe0)) % 31)) % 31)
    assert _check(bytes(z) , G.ZLIB) == b""
    assert _check(bytes([0x88, 0x1c]) + _deflate(data, G.ZLIB)[2:], G.ZLIB) == b""


@pytest.mark.parametrize("fmt", [G.GZIP, G.ZLIB])
def test_concatenated_members_and_trailing_bytes(fmt):
    rng = random.Random(5)
    a, b, c = b"first member " * 30, bytes(rng.getrandbits(8) for _ in range(70000)), b""
    sa, sb, sc = _deflate(a, fmt), _deflate(b, fmt, 1), _deflate(c, fmt)
    assert _check(sa + sb, fmt) == a + b
    assert _check(sa + sc + sb + sa, fmt) == a + b + a
    assert _check(sa + b"garbage after the member", fmt) == a            # silently ignored by the Message-parsing variant
    assert _check(sa + sb[:len(sb) // 2], fmt) == a + G.gzip_input_stream(sb[:len(sb) // 2], fmt)
    assert _check(sa + b"\x1f", fmt) == a
    assert _check(sb + b"\x00" * 7, fmt) == b


@pytest.mark.parametrize("fmt", [G.GZIP, G.ZLIB])
def test_truncated_and_corrupted_streams(fmt):
    rng = random.Random(77)
    small = [b"tiny", b"abcabcabcabcabcabc" * 9, bytes(rng.getrandbits(8) for _ in range(300))]
    for data in small:
        for level, strategy in [(0, 0), (6, 0), (6, zlib.Z_FIXED)]:
            s = _deflate(data, fmt, level, strategy)
            for cut in range(len(s) + 1):
                _check(s[:cut], fmt)
            for pos in range(len(s)):
                for bit in (0, 3, 7):
                    t = bytearray(s); t[pos] ^= 1 << bit
                    _check(bytes(t), fmt)
    # bigger streams: what survives an error depends on the 64 KiB call boundaries
    big = [(b"0123456789abcdef" * 5000) + bytes(rng.getrandbits(8) for _ in range(90000)) + b"z" * 140000,
           bytes(rng.getrandbits(8) for _ in range(200000)), b"\x00" * 300000]
    delivered_partial = 0
    for data in big:
        for level, strategy in [(0, 0), (1, 0), (6, 0), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY)]:
            s = _deflate(data, fmt, level, strategy)
            for _ in range(40):
                t = bytearray(s)
                pos = rng.randrange(len(t)); t[pos] ^= 1 << rng.randrange(8)
                w = _check(bytes(t), fmt)
                delivered_partial += 0 < len(w) < len(data)
            for _ in range(10):
                _check(s[:rng.randrange(len(s))], fmt)
    assert delivered_partial > 20


def test_handcrafted_code_length_sets():
    """Dynamic-block headers zlib rejects or accepts in its own way: incomplete / over-subscribed sets, a single 1-bit code,
    no distance codes, repeats running over the end, missing end-of-block."""
    rng = random.Random(3)

    class Bits:
        def __init__(self): self.v = 0; self.n = 0
        def put(self, val, k):
            self.v |= (val & ((1 << k) - 1)) << self.n; self.n += k
        def bytes(self): return self.v.to_bytes((self.n + 7) // 8, "little")

    def wrap(raw):
        return bytes([0x78, 0x9c]) + raw
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    outcomes = {}
    for trial in range(3000):
        b = Bits()
        # a stored block first: its bytes survive a truncated header behind it but not an erroring one (same inflate() call)
        b.put(0, 1); b.put(0, 2); b.put(0, 5); b.put(5, 16); b.put(5 ^ 0xffff, 16)
        for ch in b"keep!": b.put(ch, 8)
        b.put(1, 1); b.put(2, 2)
        nlen = rng.choice([257, 258, 260, 286, 287, 288]); ndist = rng.choice([1, 2, 5, 30, 31, 32]); ncode = rng.choice([4, 8, 19])
        b.put(nlen - 257, 5); b.put(ndist - 1, 5); b.put(ncode - 4, 4)
        mode = rng.randrange(5)
        cll = [0] * 19
        if mode == 0:
            pass                                            # all zero: a table of invalid codes
        elif mode == 1:
            cll[rng.choice([0, 1, 2, 18])] = 1               # one code only: incomplete, rejected for this alphabet
        elif mode == 2:
            for s in rng.sample(range(19), 2): cll[s] = 1    # complete, two symbols
        else:
            for s in rng.sample(range(19), rng.choice([3, 4, 6, 8])): cll[s] = rng.choice([1, 2, 3, 3, 4])
        for i in range(ncode):
            b.put(cll[order[i]], 3)
        for _ in range(rng.randrange(1, 60)):
            b.put(rng.getrandbits(8), 8)
        raw = b.bytes()
        outcomes[_check(wrap(raw), G.ZLIB)] = outcomes.get(_check(wrap(raw), G.ZLIB), 0) + 1
        for cut in range(10, len(raw), 3):
            w = _check(wrap(raw[:cut]), G.ZLIB)
            outcomes[w] = outcomes.get(w, 0) + 1
    assert outcomes.get(b"", 0) > 1000 and outcomes.get(b"keep!", 0) > 1000 and len(outcomes) >= 2
