#!/bin/bash
# First GPU session of a round: parity suite (all failures, not just the first), render-kernel A/B, the staging upper-bound experiment, bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf --timeout 600 2>&1 | tail -60 > gpurun_out/s1_pytest.log
timeout 300 python tools/bench_render.py > gpurun_out/s1_render.log 2>&1
R3DP_LIB=real3dportrait_b200/lib/libr3dp_b200_exp.so R3DP_RS_FAKE=1 timeout 200 python tools/bench_render.py --iters 20 > gpurun_out/s1_render_fake.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
tail -25 gpurun_out/s1_pytest.log; tail -8 gpurun_out/s1_render.log; tail -3 gpurun_out/s1_render_fake.log; tail -c 1500 gpurun_out/s1_bench.err; head -c 600 gpurun_out/s1_bench.json
