#!/bin/bash
# 2-GPU rehearsal of the sweep and the bench under torchrun
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR tools/sweep.py --ds 5,9 --seconds 0.15 --resume gpurun_out/r2g_sweep_resume.json --out gpurun_out/r2g_sweep > gpurun_out/r2g_sweep.log 2>&1; echo "sweep rc=$?"; tail -14 gpurun_out/r2g_sweep.log
timeout 200 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2g_bench_2gpu.json
