#!/bin/bash
# Round-2 evidence capture (run under gpurun from the repo root, one GPU): tests, the bench line, the reference arm, the ncu launch list of
# a few resident passes and full captures of every kernel of the final pipeline (copy path and pull path).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv python tools/prof_pass.py 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_tile_search|k_tile_walk|k_resolve|k_fused|k_pack_slow" -s 5 -c 5 -o gpurun_out/r2_prof_pipeline \
    python tools/prof_pass.py 1 > gpurun_out/r2_prof_pipeline.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"k_tile_walk_pull|k_decode|k_pack_tma|k_scan_blocks|k_frame_table|k_emit_iov" -s 12 -c 6 -o gpurun_out/r2_prof_pull \
    python tools/prof_pass.py 1 1024 0 iovec > gpurun_out/r2_prof_pull.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_crc_verify" -s 1 -c 1 -o gpurun_out/r2_prof_crc \
    python tools/prof_pass.py 1 1024 1 > gpurun_out/r2_prof_crc.log 2>&1
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2_smoke.txt 2>&1
# memcheck over the paths added this round (gzip / zlib inflate, host-produced replies).  NOT the persistent-kernel tests: a k_ring that
# polls host memory under the sanitizer does not come back in any reasonable time (that is how this round's last capture lost its box).
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -m gpu -q -x \
    tests/test_gpu_replies.py "tests/test_gpu_gzip.py::test_bodies_beyond_the_device_limit_go_to_the_host" > gpurun_out/r2_sanitizer.txt 2>&1
tail -5 gpurun_out/r2_sanitizer.txt
