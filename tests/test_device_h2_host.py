"""CPU: the DEVICE h2 / HPACK source (brpc_b200/csrc/b2_h2.cuh: k_h2_consume, the server side of ParseH2Message) built for the host
(tests/cpp/gen_h2_host.py writes the harness, a "warp" of one thread) and driven through the same multi-batch scenario as
tests/test_gpu_h2_server.py, against the oracle — over more seeds than a GPU run affords, under AddressSanitizer-tight buffers, and with the
device state memory pre-filled with two different patterns: a result that depends on memory the kernel never wrote shows as a difference."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from brpc_b200.abi import H2_MSG_DT, H2_RUN_STATUS_DT, RUN_DT  # noqa: E402
from brpc_b200.messenger import make_runs  # noqa: E402
from _h2server_scenario import run_scenario  # noqa: E402


@pytest.fixture(scope="module")
def h2lib():
    cpp = os.path.join(HERE, "cpp")
    so = os.path.join(cpp, "libh2_host.so")
    deps = [os.path.join(cpp, "gen_h2_host.py"), os.path.join(cpp, "h2_host_prelude.h"), os.path.join(ROOT, "brpc_b200", "csrc", "b2_h2.cuh"),
            os.path.join(ROOT, "brpc_b200", "csrc", "b2_kernels.cuh"), os.path.join(ROOT, "brpc_b200", "csrc", "b2_core.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call([sys.executable, os.path.join(cpp, "gen_h2_host.py")])
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-w", "-I", os.path.join(cpp, "stub"), "-I", os.path.join(ROOT, "include"),
                               "-o", so, os.path.join(cpp, "h2_host.cc")])
    lib = C.CDLL(so)
    lib.h2h_create.restype = C.c_void_p
    lib.h2h_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint]
    lib.h2h_destroy.argtypes = [C.c_void_p]
    lib.h2h_add_method.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    lib.h2h_conn_reset.argtypes = [C.c_void_p, C.c_uint32]
    lib.h2h_consume.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    return lib


class HostH2Context:
    """b2_h2_conn_reset / b2_h2_process_batch around the host-built kernel, laid out like brpc_b200/csrc/b2_api.cu lays the device side out"""
    def __init__(self, lib, fill, max_conns=128, pending=8, stream_bytes=69632):      # the library's defaults (b2_h2_configure)
        self.lib = lib
        self.h = lib.h2h_create(max_conns, pending, stream_bytes, fill)
        lib.h2h_add_method(self.h, b"example.EchoService.Echo", b"EchoService", b"example.EchoService", 1)
        self.fill = fill

    def h2_conn_reset(self, i):
        self.lib.h2h_conn_reset(self.h, i)

    def h2_process_batch(self, data, runs, msg_cap=None, out_cap=None, out=None):
        data = np.ascontiguousarray(data, dtype=np.uint8); runs = np.ascontiguousarray(runs, dtype=RUN_DT)
        n = len(runs)
        msg_cap = msg_cap or max(64, 64 * n)
        out_cap = out_cap or max(1 << 16, n * (1 << 17))
        region = (out_cap // n) & ~63; per_run = msg_cap // n
        # (exact-size input with a recognisable tail: the kernel may look at most a few bytes past a run, never use them)
        inp = np.concatenate([data, np.full(64, self.fill, np.uint8)])
        dev_out = np.full(region * n, self.fill, np.uint8)
        rs = np.zeros(n, H2_RUN_STATUS_DT); dmsgs = np.zeros(per_run * n, H2_MSG_DT)
        self.lib.h2h_consume(self.h, inp.ctypes.data, runs.ctypes.data, n, rs.ctypes.data, dmsgs.ctypes.data, per_run, dev_out.ctypes.data, region)
        msgs = []
        total = 0
        for r in range(n):
            k = int(rs["n_msgs"][r])
            msgs.append(dmsgs[r * per_run:r * per_run + k])
            rs["first_msg"][r] = total; total += k
        return rs, (np.concatenate(msgs) if msgs else np.zeros(0, H2_MSG_DT)), dev_out

    def close(self):
        self.lib.h2h_destroy(self.h)


@pytest.mark.parametrize("fill", [0x00, 0xa5, 0xff])
def test_device_h2_source_clean_and_violating_traffic(h2lib, fill):
    mk = lambda: HostH2Context(h2lib, fill)
    msgs, ctrl, errors = run_scenario(mk, make_runs, n_conns=48, n_calls=20, violations=0.0, seed=20260921, step_choices=[1, 9, 100, 1500, 5000, 20000])
    assert msgs > 48 * 15 and ctrl > 48 * 43 and not errors
    msgs, ctrl, errors = run_scenario(mk, make_runs, n_conns=64, n_calls=24, violations=0.25, seed=20260922, step_choices=[3, 50, 700, 4000, 30000])
    assert msgs > 300 and ctrl > 64 * 43


def test_device_h2_source_many_seeds(h2lib):
    total = 0
    for seed in range(100, 130):
        mk = lambda: HostH2Context(h2lib, seed & 0xff)
        m, c, e = run_scenario(mk, make_runs, n_conns=24, n_calls=10, violations=0.3, seed=seed, step_choices=[2, 17, 300, 2500, 12000])
        total += m
    assert total > 1500


def _records(buf):
    out = []; p = 0
    while p < len(buf):
        nl = buf[p] | (buf[p + 1] << 8); vl = buf[p + 2] | (buf[p + 3] << 8)
        out.append((bytes(buf[p + 4:p + 4 + nl]), bytes(buf[p + 4 + nl:p + 4 + nl + vl]))); p += 4 + nl + vl
    return out


def test_device_hpack_source_on_rfc_vectors_and_against_the_oracle(h2lib):
    """hpack_decode_block of b2_h2.cuh (host build): the RFC 7541 Appendix C vectors the reference asserts (test/brpc_hpack_unittest.cpp), then
    blocks made of valid pieces, mutations and truncations on a long-lived table — status and headers equal the oracle's, block after block."""
    import json
    import random
    import _oracle as O
    h2lib.h2h_hpack_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    h2lib.h2h_hpack_reset.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    ctx = HostH2Context(h2lib, 0x5a, max_conns=4, pending=1, stream_bytes=256)
    out = C.create_string_buffer(1 << 16); ol = C.c_uint32(); nh = C.c_uint32()

    def dev(b, cap=1 << 16):
        st = h2lib.h2h_hpack_decode(ctx.h, 0, b, len(b), out, cap, C.byref(ol), C.byref(nh))
        return st, _records(out.raw[:ol.value])
    vec = json.load(open(os.path.join(HERE, "golden", "hpack_vectors.json")))
    n = 0
    for t in vec["unittest"]:
        h2lib.h2h_hpack_reset(ctx.h, 0, t["max_table_size"])
        for s in t["steps"]:
            st, hdrs = dev(bytes.fromhex(s["bytes_hex"]))
            assert st == 0 and hdrs == [(a.lower().encode(), b.encode()) for a, b in s["headers"]], t["test"]
            n += len(hdrs)
    assert n >= 50
    rng = random.Random(4242)
    pieces = [bytes.fromhex(s["bytes_hex"]) for t in vec["unittest"] for s in t["steps"]] + [bytes.fromhex(s["hex"]) for s in vec["seed_corpus"]]
    h2lib.h2h_hpack_reset(ctx.h, 0, 4096); orc = O.HPack(4096)
    agree = 0
    for k in range(6000):
        b = bytearray(rng.choice(pieces))
        c = rng.random()
        if c < 0.25 and b: del b[rng.randrange(len(b)):]
        elif c < 0.5 and b: b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        elif c < 0.6: b += rng.choice(pieces)
        elif c < 0.65: b = bytearray(bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 20))))
        want = orc.decode_block(bytes(b)); got = dev(bytes(b))
        assert got[0] == want[0], (k, bytes(b).hex(), got[0], want[0])
        if want[0] == 0:
            assert got[1] == want[1], (k, bytes(b).hex())
            agree += 1
        else:                                   # a failed block leaves the connection dead in both: start both tables afresh
            h2lib.h2h_hpack_reset(ctx.h, 0, 4096); orc = O.HPack(4096)
    assert agree > 1500
    ctx.close()
