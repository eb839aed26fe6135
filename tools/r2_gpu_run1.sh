#!/bin/bash
# Round-2 GPU call 1: the new per-shot parity tests, baseline numbers of the round-1 kernels near threshold,
# ncu captures of cfg3 at p = 0.026 and of the block tile at d = 17.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_logical.py -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" 
tail -5 gpurun_out/r2a_tests.log
python tools/sweep_threshold.py --ds 11 --ps 0.005,0.01,0.02,0.026 --seconds 0.5 > gpurun_out/r2a_sweep_cfg3.txt 2>&1
python tools/sweep_threshold.py --ds 13,17,21,25 --ps 0.015,0.026 --seconds 0.5 > gpurun_out/r2a_sweep_big.txt 2>&1
cat gpurun_out/r2a_sweep_cfg3.txt gpurun_out/r2a_sweep_big.txt
python bench.py --p 0.026 --steps 3 --no-cpu --shots-per-step 1048576 > gpurun_out/r2a_bench_p026.json 2> gpurun_out/r2a_bench_p026.err; echo "bench rc=$?"
cat gpurun_out/r2a_bench_p026.json | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mc_uf -c 1 -o gpurun_out/r2a_p026 -f python bench.py --p 0.026 --steps 1 --no-cpu --shots-per-step 262144 > gpurun_out/r2a_ncu_p026.log 2>&1; echo "ncu rc=$?"
python bench.py --d 17 --p 0.02 --steps 3 --no-cpu --shots-per-step 131072 > gpurun_out/r2a_bench_d17.json 2> gpurun_out/r2a_bench_d17.err; echo "bench d17 rc=$?"
cat gpurun_out/r2a_bench_d17.json | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mc_uf -c 1 -o gpurun_out/r2a_d17 -f python bench.py --d 17 --p 0.02 --steps 1 --no-cpu --shots-per-step 32768 > gpurun_out/r2a_ncu_d17.log 2>&1; echo "ncu d17 rc=$?"
ls -la gpurun_out/*.ncu-rep
