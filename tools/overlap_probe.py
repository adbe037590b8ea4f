"""Timing experiment (not product): how well do the front stages of one batch overlap the pack of another?
Uses the debug B2_STAGE_MASK (1 front stages, 2 k_pack_tma, 4 k_pack_slow), re-read at every upload."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brpc_b200
import bench

def main():
    torch.cuda.set_device(0)
    buf, data, runs, n_full, nbytes = bench.build_batch(4, 0, payload=1024, checksum=0, kind=0)
    def mk(mask):
        c = brpc_b200.Context(device=0, max_batch_bytes=nbytes + (1 << 20), max_msgs=n_full + 4096, max_runs=bench.N_SOCKETS,
                              tile_bytes=0, max_resp_bytes=nbytes + 96 * n_full + (4 << 20))
        os.environ["B2_STAGE_MASK"] = "7"
        c.upload_ptr(buf.ptr, nbytes, runs); c.launch(); c.wait()      # prime every intermediate buffer
        os.environ["B2_STAGE_MASK"] = str(mask)
        c.upload_ptr(buf.ptr, nbytes, runs)
        os.environ["B2_STAGE_MASK"] = "7"
        return c
    def timeit(ctxs, steps=200):
        for i in range(10 * len(ctxs)): ctxs[i % len(ctxs)].launch()
        for c in ctxs: c.wait()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(steps * len(ctxs)): ctxs[i % len(ctxs)].launch()
        for c in ctxs: c.wait()
        torch.cuda.synchronize()
        return round((time.perf_counter() - t0) * 1e3 / steps, 4)
    out = {}
    full = [mk(7) for _ in range(3)]
    out["full_x1"] = timeit(full[:1]); out["full_x2_per_round"] = timeit(full[:2]); out["full_x3_per_round"] = timeit(full[:3])
    front = [mk(1) for _ in range(3)]; pack = [mk(2) for _ in range(2)]; slow = mk(4)
    out["front_x1"] = timeit(front[:1]); out["front_x2_per_round"] = timeit(front[:2]); out["front_x3_per_round"] = timeit(front[:3])
    out["pack_x1"] = timeit(pack[:1]); out["pack_x2_per_round"] = timeit(pack[:2]); out["slow_x1"] = timeit([slow])
    out["front+pack_per_round"] = timeit([front[0], pack[0]])
    out["front+front+pack_per_round"] = timeit([front[0], front[1], pack[0]])
    out["2front+2pack_per_round"] = timeit([front[0], pack[0], front[1], pack[1]])
    print(json.dumps(out, indent=1), flush=True)

if __name__ == "__main__":
    main()
