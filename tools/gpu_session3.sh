#!/bin/bash
# Render-kernel iteration: parity suite, A/B of the render variants (single pass and 48+48), ncu --set full of the streaming kernel, short bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf --timeout 600 2>&1 | tail -30 > gpurun_out/s3_pytest.log
timeout 300 python tools/bench_render.py > gpurun_out/s3_render.log 2>&1
timeout 300 python tools/bench_render.py --fine 48 > gpurun_out/s3_render_fine.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -f -o gpurun_out/s3_render_stream python tools/bench_render.py --iters 2 --only stream_d8 > gpurun_out/s3_ncu_rs.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
tail -12 gpurun_out/s3_pytest.log; cat gpurun_out/s3_render.log gpurun_out/s3_render_fine.log; tail -c 600 gpurun_out/s3_bench.err; head -c 400 gpurun_out/s3_bench.json
