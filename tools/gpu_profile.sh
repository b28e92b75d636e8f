#!/bin/bash
# One GPU: the evidence captures under profiles/ - warm-cache launch list of a step, ncu --set full of the SR launches, the streaming render
# kernel and the stand-alone sampler.   gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh'   then read with tools/ncu_lines.py / ncu -i ... --page raw
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-graph --no-extra-configs --no-cpu-baseline --sustain-seconds 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/prof_launches.csv $B > gpurun_out/prof_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc3|fir_tma|upconv_edge" -s 18 -c 6 -f -o gpurun_out/prof_sr $B > gpurun_out/prof_ncu_sr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_stream -s 3 -c 1 -f -o gpurun_out/prof_render_stream python tools/bench_render.py --iters 2 --only stream_d8 > gpurun_out/prof_ncu_rs.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:triplane_sample -s 3 -c 1 -f -o gpurun_out/prof_sample python tools/bench_ops.py --op sample > gpurun_out/prof_ncu_sample.log 2>&1
ls -la gpurun_out | grep " prof_"
