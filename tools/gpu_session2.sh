#!/bin/bash
# Profiling session: re-run the one failed parity test, warm-cache launch list of one bench step, ncu --set full of the render kernels and the SR kernels.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -rf --timeout 300 -k "plane_layouts" 2>&1 | tail -5 > gpurun_out/s2_pytest.log
B="python bench.py --steps 2 --warmup 3 --no-graph --no-extra-configs --no-cpu-baseline --sustain-seconds 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/s2_launches.csv $B > gpurun_out/s2_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -f -o gpurun_out/s2_render_stream python tools/bench_render.py --iters 2 > gpurun_out/s2_ncu_rs.log 2>&1
R3DP_RENDER=tile timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -f -o gpurun_out/s2_render_tile python tools/bench_render.py --iters 2 > gpurun_out/s2_ncu_rt.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc3|fir_tma|upconv_edge|triplane_sample" -s 18 -c 7 -f -o gpurun_out/s2_sr $B > gpurun_out/s2_ncu_sr.log 2>&1
tail -3 gpurun_out/s2_pytest.log; tail -3 gpurun_out/s2_ncu_rs.log; tail -3 gpurun_out/s2_ncu_sr.log; ls -la gpurun_out
