"""e2e probe (not product): msgs/s through b2_batch_submit/collect with pinned host buffers in the four
(input, resp) mode combinations, `depth` batches in flight.  python tools/e2e_modes_probe.py [run_mib] [depth]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, brpc_b200
run_mib = int(sys.argv[1]) if len(sys.argv) > 1 else 4
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bench.pin_to_gpu_numa(0)
buf, data, runs, n_full, nbytes = bench.build_batch(run_mib, 0)
mk = lambda: brpc_b200.Context(device=0, max_batch_bytes=nbytes + (1 << 20), max_msgs=n_full + 4096, max_runs=64, max_resp_bytes=nbytes + 96 * n_full + (4 << 20))
ctxs = [mk() for _ in range(depth)]
for im, rm in [(0, 0), (0, 1), (1, 0), (1, 1)]:
    for cx in ctxs:
        cx.set_modes(im, rm)
        for _ in range(2):
            cx.process_batch_ptr(buf.ptr, nbytes, runs)
    steps = 30
    t0 = time.perf_counter()
    for s in range(depth): ctxs[s].submit_ptr(buf.ptr, nbytes, runs)
    for s in range(steps):
        cx = ctxs[s % depth]
        rs, msgs, resp, info = cx.collect()
        assert len(msgs) == n_full
        if s + depth < steps: cx.submit_ptr(buf.ptr, nbytes, runs)
    dt = (time.perf_counter() - t0) / steps
    d2h = len(msgs) * 64 + len(rs) * 32 + int(rs["resp_bytes"].sum()) + (16 * len(msgs) if rm else 0)
    print("input=%s resp=%s: %.3f ms/step  %.1f M msgs/s  (d2h %.1f MB/step, kernel %.3f ms)" % (["copy", "pull"][im], ["copy", "by_ref"][rm], dt * 1e3, n_full / dt / 1e6, d2h / 1e6, info["kernel_ms"]))
    ctxs[0].set_modes(im, rm); ctxs[0].upload_ptr(buf.ptr, nbytes, runs)
    for _ in range(3):
        ctxs[0].execute()
    print("   stages:", ", ".join("%s %.1f" % (n, ms * 1e3) for n, ms in ctxs[0].stage_times()))
