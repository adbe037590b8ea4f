"""Latency probe (not product): where the ~70 us of a 64-request b2_process_batch go."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import brpc_b200
from brpc_b200 import press
from brpc_b200.abi import PinnedBuffer
torch.cuda.set_device(0)
N = 64
ctx = brpc_b200.Context(device=0, max_batch_bytes=1 << 20, max_msgs=4096, max_runs=N)
sp = press.spec(payload_bytes=1024)
f = press.frame(sp, 1); stride = (len(f) + 15) // 16 * 16
buf = PinnedBuffer(N * stride); runs = np.zeros(N, dtype=brpc_b200.RUN_DT)
for s in range(N):
    fr = press.frame(sp, (s << 32) + 7); buf.array[s * stride:s * stride + len(fr)] = np.frombuffer(fr, np.uint8); runs[s] = (s, s * stride, len(fr), 1, 0)
def med(fn, n=2000):
    for _ in range(100): fn()
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); t.append((time.perf_counter() - t0) * 1e6)
    t.sort(); return t[len(t) // 2]
whole = med(lambda: ctx.process_batch_ptr(buf.ptr, N * stride, runs))
def up_only(): ctx.upload_ptr(buf.ptr, N * stride, runs); torch.cuda.synchronize()
def up_exec(): ctx.upload_ptr(buf.ptr, N * stride, runs); ctx.launch(); ctx.wait()
def up_exec_dl(): ctx.upload_ptr(buf.ptr, N * stride, runs); ctx.launch(); ctx.wait(); ctx.download()
a, b, c = med(up_only), med(up_exec), med(up_exec_dl)
ctx.upload_ptr(buf.ptr, N * stride, runs)
kern = med(lambda: (ctx.launch(), ctx.wait()))
noop = med(lambda: torch.cuda.synchronize())
print("process_batch %.1f us | upload+sync %.1f | upload+kernel+wait %.1f | +download %.1f | launch+wait alone %.1f | bare sync %.1f" % (whole, a, b, c, kern, noop))
