"""Builds profiles/r2_summary.md from the JSON lines / ncu summaries brought back in gpurun_out/ (copies them to profiles/ first)."""
import json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def load(name):
    for d in (P, G):
        p = os.path.join(d, name)
        if os.path.exists(p) and os.path.getsize(p):
            try:
                return json.loads(open(p).read().strip().splitlines()[-1])
            except Exception:
                pass
    return None


def keep(name):
    src = os.path.join(G, name)
    if os.path.exists(src) and os.path.getsize(src):
        shutil.copy(src, os.path.join(P, name))


for n in ["r2_bench_n1.json", "r2_bench_reference.json", "r2_bench_n2.json", "r2_bench_n4.json", "r2_bench_n8.json", "r2_launches.csv", "r2_pytest_gpu.txt", "r2_sweep.jsonl",
          "r2_grpc_n1.json", "r2_grpc_n4.json", "r2_stream_n1.json", "r2_stream_n8.json", "r2_sweep_table.md", "r2_smoke.txt", "r2_transport.jsonl", "r2_sanitizer.txt", "r2_scale_n2.json", "r2_scale_n4.json", "r2_scale_n8.json"]:
    keep(n)
out = ["# Round 2 — measured on B200 (driver-independent runs of this round; commands in `profiles/r2_capture.sh`, `r2_sweep.sh`)", ""]
b = load("r2_bench_n1.json"); r = load("r2_bench_reference.json")
if b:
    rp = b["roofline_pipeline"]; e = b["e2e_modes"]
    out += ["## N = 1, 64 connections x 4 MiB, 1 KB payloads (`python bench.py`)", "",
            "| quantity | value |", "|---|---|",
            "| `value` (resident, %s, %d passes per step) | %.3f G msgs/s, %.1f us per pass |" % (b["config"]["value_path"], b["config"]["passes_per_step"], b["value"] / 1e9, 1e3 * b["ms_per_step"] / b["config"]["passes_per_step"]),
            "| whole-path roofline | %.3f of %.1f GB/s (%.0f B/msg algorithmic) |" % (rp["frac"], b["roofline"]["peak"], rp["algorithmic_bytes_per_msg"]),
            "| dominant kernel `%s` | %.1f us, %.0f GB/s = **%.3f** of the measured HBM peak |" % (b["roofline"]["kernel"], 1e3 * b["roofline"]["kernel_ms"], b["roofline"]["achieved"], b["roofline"]["frac"]),
            "| stages of one pass (us) | %s |" % ", ".join("%s %.1f" % (k, 1e3 * v) for k, v in rp["stage_ms"].items()),
            "| `e2e` pull + by-ref | **%.1f M msgs/s** (%.2f ms/step, H2D %.1f MB + D2H %.1f MB per step) |" % (e["pull_by_ref"]["value"] / 1e6, e["pull_by_ref"]["ms_per_step"], e["pull_by_ref"]["h2d_bytes_per_step"] / 1e6, e["pull_by_ref"]["d2h_bytes_per_step"] / 1e6),
            "| e2e pull + iovec list (what `b2::GpuTransport` runs) | %.1f M msgs/s |" % (e["pull_iovec"]["value"] / 1e6) if "pull_iovec" in e else "| | |",
            "| e2e copy + by-ref | %.1f M msgs/s (H2D %.1f MB, D2H %.1f MB) |" % (e["copy_by_ref"]["value"] / 1e6, e["copy_by_ref"]["h2d_bytes_per_step"] / 1e6, e["copy_by_ref"]["d2h_bytes_per_step"] / 1e6),
            "| e2e copy both ways (round 1's mode) | %.1f M msgs/s (H2D %.1f MB, D2H %.1f MB) |" % (e["copy"]["value"] / 1e6, e["copy"]["h2d_bytes_per_step"] / 1e6, e["copy"]["d2h_bytes_per_step"] / 1e6)]
    if b.get("e2e_messenger") and "msgs_per_s" in b["e2e_messenger"]:
        m = b["e2e_messenger"]; out += ["| e2e through `b2::GpuTransport` (C++, %d groups / %d host threads, writev to /dev/null) | %.1f M msgs/s |" % (m.get("groups", 0), m.get("host_threads", 0), m["msgs_per_s"] / 1e6)]
    if b.get("latency"):
        l = b["latency"]; out += ["| latency, 64 x 1 request, ring (%s) | p50 %.1f us, p99 **%.1f us**, %.3f launches per batch |" % (l.get("mode", ""), l["p50_us"], l["p99_us"], l["kernel_launches_per_batch"]),
                                   "| latency, blocking `b2_process_batch` | p50 %.1f us, p99 %.1f us |" % (l["blocking_call"]["p50_us"], l["blocking_call"]["p99_us"])]
    out += ["| clocks during the timed regions | %s |" % json.dumps(b["clocks"])]
    if b.get("cpu_baseline"): out += ["| oracle port, 1 host core | %.2f M msgs/s |" % (b["cpu_baseline"]["value"] / 1e6)]
if r:
    out += ["| reference arm (oracle port, %d host threads) | %.1f M msgs/s |" % (r["cpu_baseline"]["cores"], r["value"] / 1e6)]
    if b: out += ["| e2e / reference arm | **%.2fx** (copy both ways: %.2fx); resident / reference arm %.1fx |" % (b["e2e"]["value"] / r["value"], b["e2e_modes"]["copy"]["value"] / r["value"], b["value"] / r["value"])]
out += [""]
f = load("r2_bench_n1_final_build.json")
if f:
    rp = f["roofline_pipeline"]
    out += ["## Same run on the final build (`k_fused` prefetches the next tile into L2; `python bench.py --no-cpu-baseline --no-latency`)", "",
            "| quantity | value |", "|---|---|",
            "| `value` | **%.3f G msgs/s**, %.1f us per pass with %d batches in flight |" % (f["value"] / 1e9, 1e3 * f["ms_per_step"] / f["config"]["passes_per_step"], f["config"].get("pipeline_depth", 2)),
            "| dominant kernel `%s` | %.1f us, %.0f GB/s = **%.3f** of the measured HBM peak (ncu DRAM traffic %.0f MB vs %.0f MB algorithmic) |" % (f["roofline"]["kernel"], 1e3 * f["roofline"]["kernel_ms"], f["roofline"]["achieved"], f["roofline"]["frac"], (f["roofline"]["traffic"] or 0) / 1e6, f["roofline"]["algorithmic_bytes_per_launch"] / 1e6),
            "| whole-path roofline | %.3f; stages of one pass (us): %s |" % (rp["frac"], ", ".join("%s %.1f" % (k, 1e3 * v) for k, v in rp["stage_ms"].items())),
            "| e2e | pull + by-ref %.1f, pull + iovec %.1f, copy + by-ref %.1f, copy %.1f M msgs/s; through `b2::GpuTransport` %.1f M msgs/s |" % tuple([f["e2e_modes"][k]["value"] / 1e6 for k in ("pull_by_ref", "pull_iovec", "copy_by_ref", "copy")] + [f.get("e2e_messenger", {}).get("msgs_per_s", 0) / 1e6]),
            "| clocks | %s |" % json.dumps(f["clocks"]), ""]
sc = [(n, load("r2_bench_n%d.json" % n)) for n in (1, 2, 4, 8)]
if any(x for _, x in sc[1:]):
    out += ["## Scaling (torchrun, one process per GPU, weak: 64 connections per GPU)", "", "| N | value G msgs/s | e2e pull_by_ref M msgs/s | e2e copy M msgs/s | by wall clock G msgs/s |", "|---|---|---|---|---|"]
    for n, x in sc:
        if x: out += ["| %d | %.2f | %.1f | %.1f | %.2f |" % (n, x["value"] / 1e9, x["e2e"]["value"] / 1e6, x["e2e_modes"]["copy"]["value"] / 1e6, x.get("value_by_wall_clock", 0) / 1e9)]
    out += [""]
for title, names in (("h2 / gRPC 4 KB unary echo (BASELINE configs[3])", ["r2_grpc_n1.json", "r2_grpc_n4.json"]), ("256 KiB snappy streaming frames (BASELINE configs[4])", ["r2_stream_n1.json", "r2_stream_n8.json"])):
    rows = [(n, load(n)) for n in names]
    if any(x for _, x in rows):
        out += ["## " + title, ""]
        for n, x in rows:
            if x: out += ["* `%s`: N=%d, %.3g %s, %.3f ms/step%s" % (n, x["n_gpus"], x["value"], x["unit"], x["ms_per_step"], (", counters all-reduced over NCCL: %s" % x["counters_allreduced"]) if "counters_allreduced" in x else "")]
        out += [""]
p = os.path.join(G, "r2_sweep_table.md")
if os.path.exists(p):
    out += ["## rpc_press payload sweep (BASELINE configs[2], one GPU)", "", open(p).read(), ""]
p = os.path.join(P, "r2_ncu_summary.json")
if os.path.exists(p):
    k = json.load(open(p))["kernels"]
    out += ["## ncu `--set full`, one launch of every kernel of a pass (`profiles/r2_ncu_summary.json`)", "", "| kernel | us | DRAM read MB | DRAM written MB | issue active % | warps active % | regs |", "|---|---|---|---|---|---|---|"]
    for name, v in k.items():
        out += ["| `%s` | %.1f | %.1f | %.1f | %.1f | %.1f | %d |" % (name, v.get("duration_us", 0), v.get("dram_read_MB", 0), v.get("dram_write_MB", 0), v.get("issue_active_pct", 0), v.get("warps_active_pct", 0), int(v.get("registers", 0)))]
    out += [""]
notes = os.path.join(P, "r2_notes.md")
if os.path.exists(notes):
    out += [open(notes).read()]
open(os.path.join(P, "r2_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))

