"""brpc_b200 — B200-native implementation of apache/brpc's message-processing hot path.

The product is the C-ABI library `libb2rpc.so` (include/b2rpc.h) built from
`csrc/` for sm_100a; this package is the thin host-side mirror used by the
tests and the benchmark.  There is no CPU implementation in here: importing
works without a GPU, every compute call raises `B2Error(B2_E_NO_DEVICE)`.
"""
from .abi import (B2Error, Context, Method, ECHO_METHOD, lib, lib_path,  # noqa: F401
                  RUN_DT, RUN_STATUS_DT, MSG_DT)
from .messenger import GpuInputMessenger, make_runs  # noqa: F401

__all__ = ["B2Error", "Context", "Method", "ECHO_METHOD", "GpuInputMessenger", "make_runs",
           "RUN_DT", "RUN_STATUS_DT", "MSG_DT", "lib", "lib_path"]
