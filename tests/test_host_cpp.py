"""The C++ host side (brpc_b200/host: b2::IOBuf, b2::GpuInputMessenger) — builds tests/cpp/host_test
with g++ and runs it: `cpu` (IOBuf contract, allocator hook; no GPU) and `gpu` (messenger end to end,
responses byte-identical to the oracle)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "host_test")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "host_test.cc")
    deps = [src, os.path.join(ROOT, "brpc_b200", "host", "iobuf.h"), os.path.join(ROOT, "brpc_b200", "host", "input_messenger.h"), os.path.join(ROOT, "brpc_b200", "host", "h2_messenger.h"), os.path.join(ROOT, "brpc_b200", "host", "protocol.h")]
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-Wall", "-o", BIN, src,
                           "-L" + os.path.join(ROOT, "brpc_b200"), "-lb2rpc",
                           "-L" + os.path.join(ROOT, "brpc_b200", "tools"), "-lb2press",
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "brpc_b200"),
                           "-Wl,-rpath," + os.path.join(ROOT, "brpc_b200", "tools"),
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])


TBIN = os.path.join(ROOT, "tests", "cpp", "transport_test")


def _build_transport():
    src = os.path.join(ROOT, "tests", "cpp", "transport_test.cc")
    deps = [src] + [os.path.join(ROOT, "brpc_b200", "host", h) for h in ("iobuf.h", "input_messenger.h", "gpu_transport.h")]
    if os.path.exists(TBIN) and all(os.path.getmtime(TBIN) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-Wall", "-pthread", "-o", TBIN, src, "-L/usr/local/cuda/lib64", "-lcudart",
                           "-L" + os.path.join(ROOT, "brpc_b200"), "-lb2rpc",
                           "-L" + os.path.join(ROOT, "brpc_b200", "tools"), "-lb2press",
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "brpc_b200"),
                           "-Wl,-rpath," + os.path.join(ROOT, "brpc_b200", "tools"),
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])


def test_socket_write_queue_cpp():
    """Socket::Write / StartWrite / KeepWrite (socket.cpp:1604-1889): 8 producer threads, one socket, tiny send buffer."""
    import brpc_b200.press  # noqa: F401
    _build_transport()
    out = subprocess.run([TBIN, "queue"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "write queue ok" in out.stdout


@pytest.mark.gpu
def test_gpu_transport_cpp():
    """b2::GpuTransport over socketpairs: pipelined rounds, pull + by-reference replies (and the copy modes): every byte the clients
    read back equals the oracle's response stream."""
    import brpc_b200.press  # noqa: F401
    _build_transport()
    out = subprocess.run([TBIN, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("transport ok") == 7      # 3 Socket::Write variants, 4 with a reply sink (3 of them on the device-written iovec list)


def test_iobuf_contract_cpp():
    import brpc_b200.press  # noqa: F401  (builds libb2press.so if missing)
    _build()
    out = subprocess.run([BIN, "cpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "iobuf ok" in out.stdout and "adapters ok" in out.stdout


@pytest.mark.gpu
def test_gpu_input_messenger_cpp():
    import brpc_b200.press  # noqa: F401
    _build()
    out = subprocess.run([BIN, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "messenger ok" in out.stdout and "protocol shim ok" in out.stdout
