#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r2j_smoke.log 2>&1
rc=$?; tail -2 gpurun_out/r2j_smoke.log; echo "smoke rc=$rc"
if [ $rc -ne 0 ]; then exit 1; fi
T="timeout 300 python tools/fast_tune.py"
$T --d 11 --p 0.01 --tw 32 --adj 0,1 --bps 1,2 > gpurun_out/r2j_tune_d11.txt 2>&1; cat gpurun_out/r2j_tune_d11.txt
LUF_FAST_CAP_SCALE=0.8 $T --d 11 --p 0.01 --tw '' > gpurun_out/r2j_tune_d11_c08.txt 2>&1; cat gpurun_out/r2j_tune_d11_c08.txt
LUF_FAST_CAP_SCALE=0.45 $T --d 11 --p 0.01 --tw '' > gpurun_out/r2j_tune_d11_c045.txt 2>&1; cat gpurun_out/r2j_tune_d11_c045.txt
$T --d 11 --p 0.026 --shots 1048576 --tw '' > gpurun_out/r2j_tune_d11_p026.txt 2>&1; cat gpurun_out/r2j_tune_d11_p026.txt
$T --d 11 --p 0.02 --shots 1048576 --tw '' > gpurun_out/r2j_tune_d11_p02.txt 2>&1; cat gpurun_out/r2j_tune_d11_p02.txt
$T --d 13 --p 0.015 --shots 1048576 --tw 32,64 --adj 0 --bps 1,2 > gpurun_out/r2j_tune_d13.txt 2>&1; cat gpurun_out/r2j_tune_d13.txt
$T --d 17 --p 0.015 --shots 262144 --tw 64,128 --adj 0 --bps 1,2 > gpurun_out/r2j_tune_d17.txt 2>&1; cat gpurun_out/r2j_tune_d17.txt
$T --d 17 --p 0.026 --shots 65536 --tw '' > gpurun_out/r2j_tune_d17_p026.txt 2>&1; cat gpurun_out/r2j_tune_d17_p026.txt
$T --d 25 --p 0.015 --shots 65536 --tw 128,256 --adj 0 --bps 1 > gpurun_out/r2j_tune_d25.txt 2>&1; cat gpurun_out/r2j_tune_d25.txt
$T --d 25 --p 0.026 --shots 16384 --tw '' > gpurun_out/r2j_tune_d25_p026.txt 2>&1; cat gpurun_out/r2j_tune_d25_p026.txt
timeout 600 python -m pytest tests/test_gpu_logical.py tests/test_gpu_large.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2j_tests.log
