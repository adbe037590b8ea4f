"""CPU: the product library ITSELF — brpc_b200/csrc/b2_api.cu with every kernel of b2_kernels.cuh / b2_h2.cuh — built for the host on an emulated
CUDA execution model (tests/cpp/cuda_emul.h: blocks of host threads, warp collectives and __syncthreads() as barriers, TMA / cp.async copies
done at issue time; tests/cpp/gen_emul_lib.py rewrites the launches and the inline PTX, nothing else), then GPU test files run against it
UNCHANGED through the C ABI (tests/emul_runner.py points this one process's ctypes loader at the emulated build; the package itself cannot
load it and has no CPU fallback).  What this checks: the data flow of the kernels and the host-side orchestration of every ABI call, bit for
bit against the oracle, on a machine without a GPU — with B2_EMUL_ASYNC=late also that every staging buffer is waited for before it is read
or refilled (the asynchronous copies then land at the latest moment the kernel's own waits allow).  What it cannot check: speed.
Default: a subset sized for the CPU suite; B2_LONG_TESTS=1 runs every GPU test file (about 25 minutes on 8 cores; results of the full run,
also under AddressSanitizer, are in profiles/r2_emulator.md)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CPP = os.path.join(HERE, "cpp")
LONG = os.environ.get("B2_LONG_TESTS") == "1"

QUICK = ["test_gpu_zz_late_fixes.py", "test_gpu_h2.py", "test_gpu_crc32c.py", "test_gpu_dump.py"]


def build(out=None, extra=()):
    so = out or os.path.join(CPP, "libb2rpc_emul.so")
    deps = [os.path.join(CPP, f) for f in ("gen_emul_lib.py", "cuda_emul.h")] + \
           [os.path.join(ROOT, "brpc_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "brpc_b200", "csrc"))] + [os.path.join(ROOT, "include", "b2rpc.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call([sys.executable, os.path.join(CPP, "gen_emul_lib.py")])
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-w", "-pthread", *extra, "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "brpc_b200", "csrc"), "-o", so, os.path.join(CPP, "emul_lib.cc"), "-ldl"])
    return so


def run_files(files, timeout, select=None, **more):
    env = dict(os.environ, B2_EMUL_LIB=build(), B2_FUZZ_SECONDS="150", **more)
    p = subprocess.run([sys.executable, os.path.join(HERE, "emul_runner.py"), *[os.path.join(HERE, f) for f in files], "-m", "gpu", "-q", "-x",
                        "-p", "no:cacheprovider", *(["-k", select] if select else [])], cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, tail
    return tail


def test_gpu_test_files_pass_on_the_emulated_library():
    tail = run_files(QUICK, 900)
    assert " passed" in tail and "failed" not in tail and "skipped" not in tail


def test_asynchronous_copies_landing_as_late_as_the_waits_allow():
    tail = run_files(["test_gpu_zz_late_fixes.py", "test_gpu_dump.py"], 900, select="off-on or dump", B2_EMUL_ASYNC="late")      # (k_fused, k_pack_tma, k_pack_slow)
    assert " passed" in tail and "failed" not in tail


def test_smoke_on_the_emulated_library():
    code = ("import ctypes, sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import emul_runner; ctypes.CDLL = emul_runner.EmulCDLL; "
            "import __graft_entry__ as g; g.smoke()" % (ROOT, HERE))
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, B2_EMUL_LIB=build()), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "bit-exact vs oracle" in p.stdout, (p.stdout + p.stderr)[-2000:]


@pytest.mark.skipif(not LONG, reason="B2_LONG_TESTS=1: every GPU test file on the emulator (~25 min)")
def test_every_gpu_test_file_on_the_emulated_library():
    files = sorted(f for f in os.listdir(HERE) if f.startswith("test_gpu_") and f.endswith(".py"))
    for f in files:
        run_files([f], 3000)


FAKE_NCCL = r"""
#include <stddef.h>
#include <stdint.h>
// a stand-in for ncclAllReduce as two ranks holding the same counters would see it: checks what b2_counters_allreduce passes, doubles in place
extern "C" int b2_fake_nccl_calls = 0;
extern "C" int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void* stream) {
    if (send != recv || dtype != 4 /*ncclInt64*/ || op != 0 /*ncclSum*/ || comm != (void*)0x1234 || !stream) return 5;   /* ncclInvalidArgument */
    int64_t* p = (int64_t*)recv;
    for (size_t i = 0; i < count; i++) p[i] *= 2;
    b2_fake_nccl_calls++;
    return 0;
}
"""


def test_counters_allreduce_through_a_stand_in_nccl(tmp_path):
    """b2_counters_allreduce (bvar Adder semantics across GPUs: one in-place ncclAllReduce of the device-resident counters) had no GPU run this
    round: here the emulated library calls a stand-in ncclAllReduce that checks the arguments and plays a second rank with equal counters."""
    src = tmp_path / "fake_nccl.cc"; so = tmp_path / "libfake_nccl.so"
    src.write_text(FAKE_NCCL)
    subprocess.check_call(["g++", "-O1", "-fPIC", "-shared", "-o", str(so), str(src)])
    code = r'''
import ctypes, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
fake = ctypes.CDLL(%r, mode=ctypes.RTLD_GLOBAL)
import emul_runner; ctypes.CDLL = emul_runner.EmulCDLL
import numpy as np, brpc_b200
from brpc_b200 import abi, press
ctx = brpc_b200.Context(device=0, max_batch_bytes=8 << 20, max_msgs=1 << 14, max_runs=64)
data = np.zeros(8 * (64 << 10), dtype=np.uint8)
runs, n_full = press.fill_batch(press.spec(payload_bytes=1024, payload_kind=1), data, 8, (64 << 10) - 77)
ctx.process_batch(data, runs)
before = list(ctx.counters())
abi.lib.b2_counters_allreduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert abi.lib.b2_counters_allreduce(ctx._h, ctypes.c_void_p(0x1234)) == 0, abi.lib.b2_last_error()
after = list(ctx.counters())
assert before[1] == n_full and after == [2 * v for v in before], (before, after)
assert ctypes.c_int.in_dll(fake, "b2_fake_nccl_calls").value == 1
assert abi.lib.b2_counters_allreduce(ctx._h, None) != 0
print("allreduce ok", before[:3], after[:3])
''' % (ROOT, HERE, str(so))
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, B2_EMUL_LIB=build()), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "allreduce ok" in p.stdout, (p.stdout + p.stderr)[-2000:]
