"""CPU: the oracle's HPACK decoder against the RFC 7541 Appendix C vectors the reference asserts
(test/brpc_hpack_unittest.cpp:30-550 -> tests/golden/hpack_vectors.json) and basic h2 frame scanning."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def vectors():
    return json.load(open(os.path.join(HERE, "golden", "hpack_vectors.json")))


def test_rfc7541_appendix_c(oracle):
    n = 0
    for t in vectors()["unittest"]:
        hp = oracle.HPack(t["max_table_size"])           # one decoder across the steps: the dynamic table persists
        for s in t["steps"]:
            st, hdrs = hp.decode_block(bytes.fromhex(s["bytes_hex"]))
            assert st == 0, (t["test"], st)
            want = [(a.lower().encode(), b.encode()) for a, b in s["headers"]]
            assert hdrs == want, (t["test"], hdrs, want)
            n += len(hdrs)
    assert n >= 50


def test_seed_corpus_does_not_crash_and_is_deterministic(oracle):
    for s in vectors()["seed_corpus"]:
        b = bytes.fromhex(s["hex"])
        r1 = oracle.HPack().decode_block(b); r2 = oracle.HPack().decode_block(b)
        assert r1 == r2 and r1[0] in (0, 1, -1)


def test_hpack_errors_and_table_semantics(oracle):
    hp = oracle.HPack(4096)
    assert hp.decode_block(b"\x80")[0] == -1                       # index 0
    assert hp.decode_block(b"\xbe")[0] == -1                       # index 62 with an empty dynamic table
    assert hp.decode_block(b"\x3f\xe2\x1f")[0] == -1               # table size update 4097 > 4096
    assert hp.decode_block(b"\x3f\xe1\x1f\x82") == (0, [(b":method", b"GET")])     # update to 4096, then a field
    assert hp.decode_block(b"\x40\x01A\x01b\xbe") == (0, [(b"a", b"b"), (b"a", b"b")])   # lower-cased name, then indexed 62
    assert hp.decode_block(b"\x00\x03abc")[0] == -1                # value missing
    assert hp.decode_block(b"\x00\x03ab")[0] == -1                 # name truncated: DecodeWithKnownPrefix maps 0 to -1
    assert hp.decode_block(b"\x82\xff")[0] == 1                    # second field's integer runs out of bytes
    assert hp.decode_block(b"\x00\x81\xff\x01a")[0] == -1          # huffman: 8 bits of padding
    st, h = oracle.HPack(64).decode_block(b"\x40\x0acustom-key\x0dcustom-header" * 2 + b"\xbe")
    assert st == 0 and len(h) == 3 and h[2] == (b"custom-key", b"custom-header")   # 55-byte entry: the second add evicts the first
    st, h = oracle.HPack(40).decode_block(b"\x40\x0acustom-key\x0dcustom-header\xbe")
    assert st == -1 and len(h) == 1                                # larger than the table: not stored (hpack.cpp:160-163), 62 is empty
