#!/bin/bash
# 8-GPU call: BASELINE configs[4] sweep (reduced shots) and the bench lines of both arms
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 360 $TR tools/sweep.py --seconds 0.4 --resume gpurun_out/r2h_sweep_resume.json --out gpurun_out/r2h_sweep > gpurun_out/r2h_sweep.log 2>&1; echo "sweep rc=$?"; tail -16 gpurun_out/r2h_sweep.log
timeout 120 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2h_bench_8gpu.json 2> gpurun_out/r2h_bench_8gpu.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2h_bench_8gpu.json
