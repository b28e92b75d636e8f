#!/bin/bash
# Stand-alone sampler A/B (MINB = 2 | 3 | 4), its parity tests and one ncu --set full capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rf --timeout 300 -k "sample" 2>&1 | tail -5 > gpurun_out/s5_pytest.log
for m in 2 3 4; do R3DP_SAMPLE_MINB=$m timeout 200 python tools/bench_ops.py --op sample > gpurun_out/s5_sample_minb$m.log 2>&1; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:triplane_sample -s 3 -c 1 -f -o gpurun_out/s5_sample python tools/bench_ops.py --op sample > gpurun_out/s5_ncu.log 2>&1
tail -3 gpurun_out/s5_pytest.log; cat gpurun_out/s5_sample_minb*.log
