"""latency probe (not product): 64 connections x one 1 KB request per batch — blocking b2_process_batch vs the persistent
ring (b2_ring_submit + b2_ring_wait), pinned host buffers."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import brpc_b200
from brpc_b200 import press
from brpc_b200.abi import PinnedBuffer
N = 64
ctx = brpc_b200.Context(device=0, max_batch_bytes=1 << 20, max_msgs=4096, max_runs=N)
sp = press.spec(payload_bytes=1024); f = press.frame(sp, 1); stride = (len(f) + 15) // 16 * 16
buf = PinnedBuffer(N * stride); runs = np.zeros(N, dtype=brpc_b200.RUN_DT)
for s in range(N):
    fr = press.frame(sp, (s << 32) + 7); buf.array[s * stride:s * stride + len(fr)] = np.frombuffer(fr, np.uint8); runs[s] = (s, s * stride, len(fr), 1, 0)
def pct(v, p): v = sorted(v); return v[int(len(v) * p)]
for by_ref in (0, 1):
    ctx.set_modes(0, by_ref)
    for _ in range(50): ctx.process_batch_ptr(buf.ptr, N * stride, runs)
    lat = []
    for _ in range(3000):
        t0 = time.perf_counter(); r = ctx.process_batch_ptr(buf.ptr, N * stride, runs); lat.append((time.perf_counter() - t0) * 1e6)
    print("by_ref=%d blocking call: p50 %.1f us  p99 %.1f us" % (by_ref, pct(lat, .5), pct(lat, .99)))
    ctx.ring_start()
    for _ in range(50): ctx.ring_wait(ctx.ring_submit(None, runs, ptr=buf.ptr, nbytes=N * stride))
    lat = []
    for _ in range(3000):
        t0 = time.perf_counter(); r = ctx.ring_wait(ctx.ring_submit(None, runs, ptr=buf.ptr, nbytes=N * stride)); lat.append((time.perf_counter() - t0) * 1e6)
    assert len(r[1]) == N and np.all(r[1]["status"] == 0)
    print("by_ref=%d ring         : p50 %.1f us  p99 %.1f us  (kernel (re)starts: %d)" % (by_ref, pct(lat, .5), pct(lat, .99), ctx.ring_launches()))
    for ring in (0, 1):
        us = np.sort(ctx.latency_probe(buf.ptr, N * stride, runs, 3000, ring))
        print("by_ref=%d C-timed %s: p50 %.1f us  p99 %.1f us  min %.1f" % (by_ref, "ring    " if ring else "blocking", us[1500], us[2970], us[0]))
    t = ctx.ring_submit(None, runs, ptr=buf.ptr, nbytes=N * stride); ctx.ring_wait(t)
    print("   device phases (ns since doorbell seen): header %d, pulled %d, body done %d, pushed %d" % tuple(ctx.ring_phase_ns(t)))
    ctx.ring_stop()
