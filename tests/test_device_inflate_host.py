"""CPU: the DEVICE inflate source (brpc_b200/csrc/b2_inflate.cuh) compiled for the host (tests/cpp/inflate_host.cc, a test harness) against the
system zlib driven like GzipInputStream (tests/_gzipstream.py) and against the oracle's sizing bound — over the same stream families as
tests/test_oracle_gzip.py: every block type and level, header fields, concatenated members, truncation at every byte, bit flips, hand-built
code-length sets.  What the GPU tests check on whole messages is checked here on the decoder itself, stream by stream."""
import ctypes as C
import os
import random
import subprocess
import zlib

import pytest

import _gzipstream as G
import _oracle as O
import test_oracle_gzip as T

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def dev():
    so = os.path.join(HERE, "cpp", "libinflate_host.so")
    src = os.path.join(HERE, "cpp", "inflate_host.cc")
    hdr = os.path.join(ROOT, "brpc_b200", "csrc", "b2_inflate.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    lib.dev_gz_input_stream.restype = C.c_uint32
    lib.dev_gz_input_stream.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_char_p, C.c_uint32, C.POINTER(C.c_int)]
    lib.dev_gz_sizing_bound.restype = C.c_uint32
    lib.dev_gz_sizing_bound.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.POINTER(C.c_int)]
    lib.dev_gz_max_out.restype = C.c_uint32
    return lib


def check(lib, stream, fmt):
    """sizing pass then the real pass with exactly the reserved room — as decode_one / pack_one do"""
    big = C.c_int(0)
    bound = lib.dev_gz_sizing_bound(stream, len(stream), fmt, C.byref(big))
    want_bound = O.lib.orc_gzip_sizing_bound(stream, len(stream), fmt, lib.dev_gz_max_out())
    if big.value:
        assert want_bound > lib.dev_gz_max_out()
        return None
    assert bound == want_bound, (len(stream), bound, want_bound)
    out = C.create_string_buffer(max(1, bound))
    n = lib.dev_gz_input_stream(stream, len(stream), fmt, out, bound, C.byref(big))
    want = G.gzip_input_stream(stream, fmt)
    assert not big.value and n <= bound and out.raw[:n] == want, (len(stream), n, len(want))
    return want


O.lib.orc_gzip_sizing_bound.restype = C.c_size_t
O.lib.orc_gzip_sizing_bound.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_size_t]


@pytest.mark.parametrize("fmt", [G.GZIP, G.ZLIB])
def test_device_source_on_valid_streams(dev, fmt):
    rng = random.Random(11)
    for data in T._payloads(rng):
        for level, strategy in [(0, 0), (1, 0), (6, 0), (9, 0), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)]:
            assert check(dev, T._deflate(data, fmt, level, strategy), fmt) == data
    a, b = b"first member " * 30, bytes(rng.getrandbits(8) for _ in range(70000))
    sa, sb = T._deflate(a, fmt), T._deflate(b, fmt, 1)
    assert check(dev, sa + sb + sa, fmt) == a + b + a
    assert check(dev, sa + b"garbage after the member", fmt) == a
    big = T._deflate(b"\x00" * (3 << 20), fmt)                      # beyond the device limit: left to the host
    assert check(dev, big, fmt) is None


def test_device_source_on_gzip_headers(dev):
    data = b"header field test " * 100
    for kw in [dict(), dict(extra=b"\x01\x02abcd"), dict(name=b"file.bin"), dict(comment=b"a comment"), dict(hcrc=True),
               dict(extra=b"x" * 300, name=b"n" * 100, comment=b"c" * 50, hcrc=True)]:
        s = T._gzip_with_header(data, **kw)
        assert check(dev, s, G.GZIP) == data
        for cut in list(range(0, min(len(s), 480))) + list(range(len(s) - 12, len(s))):
            check(dev, s[:cut], G.GZIP)
    bad = bytearray(T._gzip_with_header(data, hcrc=True)); bad[10] ^= 1
    assert check(dev, bytes(bad), G.GZIP) == b""
    for tweak in (-5, -1):
        s = bytearray(T._gzip_with_header(data)); s[tweak] ^= 0x10
        assert check(dev, bytes(s), G.GZIP) == b""
    assert check(dev, T._deflate(data, G.ZLIB), G.GZIP) == b"" and check(dev, T._deflate(data, G.GZIP), G.ZLIB) == b""


@pytest.mark.parametrize("fmt", [G.GZIP, G.ZLIB])
def test_device_source_on_truncated_and_corrupted_streams(dev, fmt):
    rng = random.Random(77)
    for data in [b"tiny", b"abcabcabcabcabcabc" * 9, bytes(rng.getrandbits(8) for _ in range(300))]:
        for level, strategy in [(0, 0), (6, 0), (6, zlib.Z_FIXED)]:
            s = T._deflate(data, fmt, level, strategy)
            for cut in range(len(s) + 1):
                check(dev, s[:cut], fmt)
            for pos in range(len(s)):
                for bit in (0, 3, 7):
                    t = bytearray(s); t[pos] ^= 1 << bit
                    check(dev, bytes(t), fmt)
    big = [(b"0123456789abcdef" * 5000) + bytes(rng.getrandbits(8) for _ in range(90000)) + b"z" * 140000, b"\x00" * 300000]
    partial = 0
    for data in big:
        for level, strategy in [(0, 0), (1, 0), (6, 0), (6, zlib.Z_FIXED)]:
            s = T._deflate(data, fmt, level, strategy)
            for _ in range(25):
                t = bytearray(s); t[rng.randrange(len(t))] ^= 1 << rng.randrange(8)
                w = check(dev, bytes(t), fmt)
                partial += w is not None and 0 < len(w) < len(data)
            for _ in range(6):
                check(dev, s[:rng.randrange(len(s))], fmt)
    assert partial > 10


def test_device_source_on_handcrafted_code_length_sets(dev):
    rng = random.Random(3)
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    seen = set()
    for trial in range(1500):
        v = 0; nbits = 0

        def put(val, k):
            nonlocal v, nbits
            v |= (val & ((1 << k) - 1)) << nbits; nbits += k
        put(0, 1); put(0, 2); put(0, 5); put(5, 16); put(5 ^ 0xffff, 16)
        for ch in b"keep!":
            put(ch, 8)
        put(1, 1); put(2, 2)
        put(rng.choice([0, 1, 3, 29, 30, 31]), 5); put(rng.choice([0, 1, 4, 29, 30, 31]), 5); ncode = rng.choice([4, 8, 19]); put(ncode - 4, 4)
        cll = [0] * 19
        mode = rng.randrange(5)
        if mode == 1:
            cll[rng.choice([0, 1, 2, 18])] = 1
        elif mode == 2:
            for s_ in rng.sample(range(19), 2): cll[s_] = 1
        elif mode >= 3:
            for s_ in rng.sample(range(19), rng.choice([3, 4, 6, 8])): cll[s_] = rng.choice([1, 2, 3, 3, 4])
        for i in range(ncode):
            put(cll[order[i]], 3)
        for _ in range(rng.randrange(1, 60)):
            put(rng.getrandbits(8), 8)
        raw = bytes([0x78, 0x9c]) + v.to_bytes((nbits + 7) // 8, "little")
        seen.add(check(dev, raw, G.ZLIB))
        for cut in range(12, len(raw), 3):
            seen.add(check(dev, raw[:cut], G.ZLIB))
    assert b"" in seen and b"keep!" in seen
