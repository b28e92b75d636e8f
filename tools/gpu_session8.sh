#!/bin/bash
# Full parity suite + bench + warm-cache launch list of one step.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf --timeout 600 2>&1 | tail -30 > gpurun_out/s8_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err
B="python bench.py --steps 2 --warmup 3 --no-graph --no-extra-configs --no-cpu-baseline --sustain-seconds 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/s8_launches.csv $B > gpurun_out/s8_launch_bench.log 2>&1
tail -8 gpurun_out/s8_pytest.log; tail -c 600 gpurun_out/s8_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s8_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['extra']['stage_ms_per_step'], 'conv', d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], 'launches', d['gpu_launches'], 'e2e', d['e2e']['value'])
PY
