#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_logical.py tests/test_gpu_uf.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2b_tests.log
timeout 600 python tools/fast_tune.py --d 11 --p 0.01 --baseline > gpurun_out/r2b_tune_d11_p01.txt 2>&1; cat gpurun_out/r2b_tune_d11_p01.txt
timeout 400 python tools/fast_tune.py --d 11 --p 0.026 --baseline --shots 1048576 --tw 16,32 > gpurun_out/r2b_tune_d11_p026.txt 2>&1; cat gpurun_out/r2b_tune_d11_p026.txt
timeout 300 python tools/fast_tune.py --d 13 --p 0.015 --baseline --shots 1048576 --tw 16,32 --bps 1 > gpurun_out/r2b_tune_d13.txt 2>&1; cat gpurun_out/r2b_tune_d13.txt
timeout 300 python tools/fast_tune.py --d 17 --p 0.015 --baseline --shots 262144 --tw 32,64,128 --adj 0 --bps 1 > gpurun_out/r2b_tune_d17.txt 2>&1; cat gpurun_out/r2b_tune_d17.txt
timeout 300 python tools/fast_tune.py --d 17 --p 0.026 --baseline --shots 65536 --tw 32,64,128 --adj 0 --bps 1 > gpurun_out/r2b_tune_d17_p026.txt 2>&1; cat gpurun_out/r2b_tune_d17_p026.txt
timeout 300 python tools/fast_tune.py --d 25 --p 0.015 --baseline --shots 65536 --tw 64,128,256 --adj 0 --bps 1,2 > gpurun_out/r2b_tune_d25.txt 2>&1; cat gpurun_out/r2b_tune_d25.txt
