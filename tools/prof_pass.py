"""profiling helper (not product): upload the bench batch once and run a few resident passes — the short command ncu wraps.
python tools/prof_pass.py [passes] [payload] [checksum] [pull|iovec]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, brpc_b200
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
payload = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cks = int(sys.argv[3]) if len(sys.argv) > 3 else 0
pull = len(sys.argv) > 4 and sys.argv[4] in ("pull", "iovec")
iovec = len(sys.argv) > 4 and sys.argv[4] == "iovec"
buf, data, runs, n_full, nbytes = bench.build_batch(4, 0, payload=payload, checksum=cks)
ctx = brpc_b200.Context(device=0, max_batch_bytes=nbytes + (1 << 20), max_msgs=n_full + 4096, max_runs=64, max_resp_bytes=2 * nbytes + (8 << 20))
if pull:
    ctx.set_modes(1, 2 if iovec else 1)
rs, msgs, resp, info = ctx.process_batch_ptr(buf.ptr, nbytes, runs)      # (sets the adaptive tile size)
rs, msgs, resp, info = ctx.process_batch_ptr(buf.ptr, nbytes, runs)
ctx.upload_ptr(buf.ptr, nbytes, runs)
for _ in range(passes):
    ms, n = ctx.execute()
print("msgs", len(msgs), "pass ms", ms, "launches", n, ctx.stage_times())
