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
