"""resident-path probe (not product): time per pass with `depth` batches in flight, and the stage times of one pass."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, brpc_b200
payload = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
buf, data, runs, n_full, nbytes = bench.build_batch(4, 0, payload=payload)
mk = lambda: brpc_b200.Context(device=0, max_batch_bytes=nbytes + (1 << 20), max_msgs=n_full + 4096, max_runs=64, max_resp_bytes=2 * nbytes + (8 << 20))
ctxs = [mk() for _ in range(3)]
for cx in ctxs:
    cx.process_batch_ptr(buf.ptr, nbytes, runs); cx.upload_ptr(buf.ptr, nbytes, runs)
for depth in (1, 2, 3):
    for s in range(20): ctxs[s % depth].launch()
    for cx in ctxs: cx.wait()
    n = 600
    for s in range(n): ctxs[s % depth].launch()
    for cx in ctxs[:depth]: cx.wait()
    ms = max(ctxs[0].elapsed_ms_to(cx) for cx in ctxs[:depth]) / n
    print("depth %d: %.1f us/pass  %.2f G msgs/s  whole-path frac %.3f" % (depth, ms * 1e3, n_full / ms / 1e6, (2209.43 * n_full / (ms * 1e-3) / 1e9) / 6578.7 if payload == 1024 else 0))
for _ in range(3): ctxs[0].execute()
print("stages:", ", ".join("%s %.1f" % (n, ms * 1e3) for n, ms in ctxs[0].stage_times()), ctxs[0].batch_info())
