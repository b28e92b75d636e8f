#!/bin/bash
# Render diet check: parity suite, render A/B, racecheck of the streaming kernel (with the explicit depth barrier), bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf --timeout 600 2>&1 | tail -30 > gpurun_out/s9_pytest.log
timeout 300 python tools/bench_render.py --only tile,stream_d8,stream_d4,stream_d16,stream_d8_two_sets > gpurun_out/s9_render.log 2>&1
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_case.py render > gpurun_out/s9_racecheck_render.log 2>&1; echo "racecheck exit $?" >> gpurun_out/s9_racecheck_render.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
tail -8 gpurun_out/s9_pytest.log; cat gpurun_out/s9_render.log; grep "Error: Race" gpurun_out/s9_racecheck_render.log | sed 's/+0x[0-9a-f]*//g' | sort | uniq -c | cut -c1-200; tail -3 gpurun_out/s9_racecheck_render.log; tail -c 500 gpurun_out/s9_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s9_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['extra']['stage_ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
for k,v in d['roofline']['extra']['configs'].items(): print(k, v['value'], v.get('ms_per_step'))
PY
