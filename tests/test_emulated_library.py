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

QUICK = ["test_gpu_zz_empty_reply_checksum.py", "test_gpu_h2.py", "test_gpu_crc32c.py", "test_gpu_dump.py"]


def build(out=None, extra=()):
    so = out or os.path.join(CPP, "libb2rpc_emul.so")
    deps = [os.path.join(CPP, f) for f in ("gen_emul_lib.py", "cuda_emul.h")] + \
           [os.path.join(ROOT, "brpc_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "brpc_b200", "csrc"))] + [os.path.join(ROOT, "include", "b2rpc.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call([sys.executable, os.path.join(CPP, "gen_emul_lib.py")])
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-w", "-pthread", *extra, "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "brpc_b200", "csrc"), "-o", so, os.path.join(CPP, "emul_lib.cc"), "-ldl"])
    return so


def run_files(files, timeout, **more):
    env = dict(os.environ, B2_EMUL_LIB=build(), B2_FUZZ_SECONDS="150", **more)
    p = subprocess.run([sys.executable, os.path.join(HERE, "emul_runner.py"), *[os.path.join(HERE, f) for f in files], "-m", "gpu", "-q", "-x",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, tail
    return tail


def test_gpu_test_files_pass_on_the_emulated_library():
    tail = run_files(QUICK, 900)
    assert " passed" in tail and "failed" not in tail and "skipped" not in tail


def test_asynchronous_copies_landing_as_late_as_the_waits_allow():
    tail = run_files(["test_gpu_zz_empty_reply_checksum.py", "test_gpu_dump.py"], 900, B2_EMUL_ASYNC="late")
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
