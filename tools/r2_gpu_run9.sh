#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r2i_smoke.log 2>&1
rc=$?; tail -2 gpurun_out/r2i_smoke.log; echo "smoke rc=$rc"
if [ $rc -ne 0 ]; then exit 1; fi
T="timeout 300 python tools/fast_tune.py"
$T --d 11 --p 0.01 --tw '' > gpurun_out/r2i_tune_d11.txt 2>&1; cat gpurun_out/r2i_tune_d11.txt
$T --d 11 --p 0.026 --shots 1048576 --tw '' > gpurun_out/r2i_tune_d11_p026.txt 2>&1; cat gpurun_out/r2i_tune_d11_p026.txt
$T --d 11 --p 0.026 --flags 2 --shots 1048576 --tw '' > gpurun_out/r2i_tune_d11_p026w.txt 2>&1; cat gpurun_out/r2i_tune_d11_p026w.txt
$T --d 17 --p 0.015 --shots 262144 --tw '' > gpurun_out/r2i_tune_d17.txt 2>&1; cat gpurun_out/r2i_tune_d17.txt
timeout 600 python -m pytest tests/test_gpu_logical.py tests/test_gpu_large.py -m gpu -q -x > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2i_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mc_uf_fast -c 1 -o gpurun_out/r2i_fast -f python bench.py --steps 1 --no-cpu --no-workloads --no-python-ref --shots-per-step 1048576 > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"
