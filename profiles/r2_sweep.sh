#!/bin/sh
# rpc_press-style payload sweep (BASELINE configs[2]: 1024 sockets per GPU x 256 KiB pending, 11 sizes x {plain, crc32c}) on one GPU; writes gpurun_out/r2_sweep.jsonl + a table
: > gpurun_out/r2_sweep.jsonl
for p in 64 128 256 512 1024 2048 4096 8192 16384 32768 65536; do
  python bench.py --sockets 1024 --run-mib 0.25 --payload $p --steps 50 --min-seconds 0.3 --no-cpu-baseline --no-latency >> gpurun_out/r2_sweep.jsonl 2>> gpurun_out/r2_sweep.err
done
for p in 64 1024 16384; do
  python bench.py --sockets 1024 --run-mib 0.25 --payload $p --checksum 1 --steps 50 --min-seconds 0.3 --no-cpu-baseline --no-latency >> gpurun_out/r2_sweep.jsonl 2>> gpurun_out/r2_sweep.err
done
python bench.py --sockets 1024 --run-mib 0.25 --payload 1024 --payload-kind 1 --steps 50 --min-seconds 0.3 --no-cpu-baseline --no-latency >> gpurun_out/r2_sweep.jsonl 2>> gpurun_out/r2_sweep.err
python - <<'PY'
import json
print("| payload | crc | value M msgs/s | whole-path frac | dominant kernel frac | e2e pull_by_ref M/s | e2e copy M/s | stages us |")
print("|---|---|---|---|---|---|---|---|")
for l in open('gpurun_out/r2_sweep.jsonl'):
    d=json.loads(l); c=d['config']
    print('| %d | %d | %.1f | %.3f | %s %.3f | %.1f | %.1f | %s |' % (c['payload_bytes'], c['request_checksum'], d['value']/1e6, d['roofline_pipeline']['frac'],
        d['roofline']['kernel'], d['roofline']['frac'], d['e2e']['value']/1e6, d['e2e_modes']['copy']['value']/1e6,
        ' '.join('%s %d' % (k, round(v*1000)) for k,v in d['roofline_pipeline']['stage_ms'].items())))
PY
