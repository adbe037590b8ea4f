"""Summarise .ncu-rep captures (read here with `ncu -i ... --page raw --csv`) into profiles/r2_ncu_summary.json:
per kernel launch — duration, DRAM bytes read / written, PCIe / sysmem reads when present, issue-active and warps-active
percentages, registers, executed warp instructions.  python tools/ncu_summary.py out.json rep1 [rep2 ...]"""
import csv, io, json, subprocess, sys

WANT = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read_MB", "dram__bytes_write.sum": "dram_write_MB",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "launch__registers_per_thread": "registers", "smsp__inst_executed.sum": "warp_instructions", "launch__grid_size": "grid", "launch__block_size": "block",
        "lts__t_sectors_srcunit_tex_aperture_sysmem_op_read.sum": "sysmem_read_sectors", "pcie__read_bytes.sum": "pcie_read_bytes", "pcie__write_bytes.sum": "pcie_write_bytes",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct"}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("b2::", "")}
        for i, h in enumerate(hdr):
            if h in WANT and r[i] != "":
                v = float(r[i].replace(",", ""))
                u = units[i]
                if WANT[h].endswith("_MB") and u == "Gbyte": v *= 1000
                if WANT[h].endswith("_MB") and u == "Kbyte": v /= 1000
                if WANT[h].endswith("_MB") and u == "byte": v /= 1e6
                if WANT[h] == "duration_us" and u == "ms": v *= 1000
                if WANT[h] == "duration_us" and u in ("ns", "nsecond"): v /= 1000
                d[WANT[h]] = v
        yield d


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    kernels = {}
    for rep in reps:
        for d in rows_of(rep):
            k = d.pop("kernel")
            kernels.setdefault(k, []).append(d)
    summary = {"workload": {"payload_bytes": 1024, "run_mib": 4, "connections": 64, "note": "bench.py batch: 64 connections x 4 MiB, 246 976 messages, tools/prof_pass.py"},
               "captures": reps, "kernels": {k: v[-1] for k, v in kernels.items()}, "all_launches": kernels}
    json.dump(summary, open(out, "w"), indent=1)
    for k, v in summary["kernels"].items():
        print("%-28s %8.1f us  dram %7.1f + %7.1f MB  issue %5.1f%%  warps %5.1f%%  regs %3d" % (k, v.get("duration_us", 0), v.get("dram_read_MB", 0), v.get("dram_write_MB", 0),
              v.get("issue_active_pct", 0), v.get("warps_active_pct", 0), int(v.get("registers", 0))))


if __name__ == "__main__":
    main()
