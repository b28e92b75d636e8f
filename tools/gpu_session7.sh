#!/bin/bash
# SR iteration: parity tests of the SR paths, then the bench with the up-conv phases interleaved (default) and sequential (R3DP_TC_MIX=0).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 600 -k "sr or tc or engine or torso or large" 2>&1 | tail -30 > gpurun_out/s7_pytest.log
for m in 1 0; do R3DP_TC_MIX=$m timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-configs --sustain-seconds 0 > gpurun_out/s7_bench_mix$m.json 2> gpurun_out/s7_bench_mix$m.err; done
B="python bench.py --steps 2 --warmup 3 --no-graph --no-extra-configs --no-cpu-baseline --sustain-seconds 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/s7_launches.csv $B > gpurun_out/s7_launch_bench.log 2>&1
tail -8 gpurun_out/s7_pytest.log; python - <<'PY'
import json
for m in (1, 0):
    try:
        d=json.loads([l for l in open(f'gpurun_out/s7_bench_mix{m}.json') if l.startswith('{')][-1])
        print('mix', m, d['value'], d['ms_per_step'], d['roofline']['extra']['stage_ms_per_step'], 'conv', d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])
    except Exception as e:
        print('mix', m, 'failed', e); print(open(f'gpurun_out/s7_bench_mix{m}.err').read()[-1500:])
PY
