"""Test infrastructure: runs pytest with brpc_b200's ctypes loader pointed at tests/cpp/libb2rpc_emul.so — the product library's own sources built
for the host on an emulated CUDA execution model (tests/cpp/cuda_emul.h) — so that the GPU test files can run on a machine without a GPU.
Only this process is affected; the product package knows nothing about it.  Usage: python tests/emul_runner.py <pytest arguments>"""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.environ.get("B2_EMUL_LIB") or os.path.join(HERE, "cpp", "libb2rpc_emul.so")       # (B2_EMUL_LIB: a sanitizer build of the same file)
_CDLL = ctypes.CDLL


class EmulCDLL(_CDLL):
    def __init__(self, name, *a, **kw):
        if name and os.path.basename(str(name)) == "libb2rpc.so":
            name = EMUL
        super().__init__(name, *a, **kw)


if __name__ == "__main__":
    ctypes.CDLL = EmulCDLL
    import pytest
    sys.exit(pytest.main(sys.argv[1:]))
