#!/bin/bash
# Render-kernel iteration: render parity tests, A/B of the render variants, ncu --set full of the streaming kernel, sample-op roofline via the short bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 600 -k "render or stream or engine or sample or degenerate or trigrid" 2>&1 | tail -30 > gpurun_out/s4_pytest.log
timeout 300 python tools/bench_render.py --only tile,stream_d8,stream_d8_nopf,stream_d4,stream_d16,stream_d8_two_sets > gpurun_out/s4_render.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -f -o gpurun_out/s4_render_stream python tools/bench_render.py --iters 2 --only stream_d8 > gpurun_out/s4_ncu_rs.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err
tail -12 gpurun_out/s4_pytest.log; cat gpurun_out/s4_render.log; tail -c 600 gpurun_out/s4_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s4_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['extra']['stage_ms_per_step'], d['roofline']['extra']['hbm_sample_op']['frac'], d['roofline']['frac'])
PY
