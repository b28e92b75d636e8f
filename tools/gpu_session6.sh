#!/bin/bash
# tc_exact bring-up: SR parity tests (fp16-operand and split-operand paths), then its speed through the bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 600 -k "sr or tc or engine or torso or large" 2>&1 | tail -40 > gpurun_out/s6_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-configs --sr-mode tc_exact > gpurun_out/s6_bench_exact.json 2> gpurun_out/s6_bench_exact.err
tail -30 gpurun_out/s6_pytest.log; tail -c 1000 gpurun_out/s6_bench_exact.err; head -c 500 gpurun_out/s6_bench_exact.json
