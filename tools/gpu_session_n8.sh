#!/bin/bash
# 8-GPU session: exchange paths bit for bit, then the bench at N = 8 (p2p default) and with the NCCL all-gather.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/check_exchange.py > gpurun_out/n8_exchange.log 2>&1
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --steps 32 --warmup 3 --no-cpu-baseline --no-extra-configs --exchange allgather > gpurun_out/n8_bench_allgather.json 2> gpurun_out/n8_bench_allgather.err
tail -6 gpurun_out/n8_exchange.log; tail -c 800 gpurun_out/n8_bench.err; grep "^{" gpurun_out/n8_bench.json | head -c 300; echo; grep "^{" gpurun_out/n8_bench_allgather.json | head -c 300
