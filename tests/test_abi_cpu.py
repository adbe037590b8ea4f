"""CPU tests of the boundary: the C-ABI library loads, exports every symbol that
include/b2rpc.h declares, struct sizes match, and it fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported():
    import brpc_b200
    hdr = open(os.path.join(ROOT, "include", "b2rpc.h")).read()
    declared = sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(brpc_b200.lib, name), "libb2rpc.so does not export %s" % name
    from brpc_b200.abi import ABI_SYMBOLS
    assert sorted(ABI_SYMBOLS) == declared


def test_struct_sizes_match_header():
    from brpc_b200 import MSG_DT, RUN_DT, RUN_STATUS_DT
    assert (RUN_DT.itemsize, RUN_STATUS_DT.itemsize, MSG_DT.itemsize) == (24, 32, 64)
    import _oracle
    assert _oracle.MSG_DT == MSG_DT and _oracle.RUN_DT == RUN_DT and _oracle.RUN_STATUS_DT == RUN_STATUS_DT


def test_fails_loudly_without_gpu():
    import brpc_b200
    if _have_gpu():
        pytest.skip("GPU present")
    with pytest.raises(brpc_b200.B2Error) as e:
        brpc_b200.Context(device=0, max_batch_bytes=1 << 20, max_msgs=1024, max_runs=16)
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_version_string():
    import brpc_b200
    assert b"sm_100a" in brpc_b200.lib.b2_version()


def test_press_frames_match_oracle_packer(oracle):
    """The load generator (tools/rpc_press.cpp) and the oracle's PackRpcRequest restatement agree byte for byte."""
    from brpc_b200 import press
    for n, att, cks in [(0, 0, 0), (16, 0, 0), (1024, 0, 0), (1024, 5, 0), (100, 0, 1), (4096, 33, 1)]:
        sp = press.spec(payload_bytes=n, attachment_bytes=att, checksum_type=cks)
        for i in [0, 1, 127, 128, 5000, (3 << 32) + 77]:
            f = press.frame(sp, i)
            attachment = bytes((ord("A") + (i + k) % 26) for k in range(att))
            g = oracle.pack_echo_request(log_id=i & 0x3fff, correlation_id=((i & 0xfffff) << 32) | (i % 7 + 1),
                                         message=b"r" * n, attachment=attachment, checksum_type=cks)
            assert f == g, (n, att, cks, i)
    sp = press.spec(payload_bytes=1024)
    # SURVEY §8 / BASELINE.md §3: 1 KB request frame = 1085 B with a 5-byte correlation_id varint;
    # the generator's ids give 5..8-byte varints
    assert 1085 <= len(press.frame(sp, 12345)) <= 1088
