#!/bin/sh
# rpc_press-style payload sweep (BASELINE configs[2]) on one GPU; writes gpurun_out/sweep_r1.jsonl
: > gpurun_out/sweep_r1.jsonl
for p in 64 256 1024 4096 16384 65536; do
  python bench.py --payload $p --steps 100 --no-cpu-baseline >> gpurun_out/sweep_r1.jsonl 2>> gpurun_out/sweep_r1.err
done
python bench.py --payload 1024 --checksum 1 --steps 100 --no-cpu-baseline >> gpurun_out/sweep_r1.jsonl 2>> gpurun_out/sweep_r1.err
python bench.py --payload 1024 --payload-kind 1 --steps 100 --no-cpu-baseline >> gpurun_out/sweep_r1.jsonl 2>> gpurun_out/sweep_r1.err
python - <<'PY'
import json
for l in open('gpurun_out/sweep_r1.jsonl'):
    d=json.loads(l); c=d['config']
    print('payload %6d cks %d | %8.1f M msgs/s | %6.3f ms/step | pipe frac %.3f | dom %s %.3f | e2e %6.1f M/s | p99 %.0f us' % (
        c['payload_bytes'], c['request_checksum'], d['value']/1e6, d['ms_per_step'], d['roofline_pipeline']['frac'],
        d['roofline']['kernel'], d['roofline']['frac'], d['e2e']['value']/1e6, d['latency']['p99_us']))
    print('   ', {k: round(v*1000) for k,v in d['roofline_pipeline']['stage_ms'].items()})
PY
