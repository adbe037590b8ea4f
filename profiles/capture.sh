#!/bin/bash
# Round-end evidence capture (run under gpurun from the repo root): tests, the bench line, the reference arm,
# the ncu launch list of one bench pass and full captures of the two heaviest kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r1_pytest_gpu.txt
python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
python bench.py --impl reference --steps 20 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv \
    python bench.py --steps 4 --warmup 3 --pipeline 1 --no-cpu-baseline --no-latency > gpurun_out/r1_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_pack_tma|k_decode" -s 16 -c 2 -o gpurun_out/r1_full \
    python bench.py --steps 2 --warmup 3 --pipeline 1 --no-cpu-baseline --no-latency > gpurun_out/r1_full.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r1_smoke.txt 2>&1
