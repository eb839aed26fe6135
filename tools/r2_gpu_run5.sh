#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -c "
import __graft_entry__ as g
g.smoke()
import localuf_b200 as L
from localuf_b200 import _engine
eng = _engine.Engine(L.Surface(11, 'phenomenological'), device=0)
print('cfg3 fast', eng.mc_run_host(0, 0, 0.01, 1, 0, 1 << 20))
print('cfg3 p026', eng.mc_run_host(0, 0, 0.026, 1, 0, 1 << 18))
" > gpurun_out/r2e_smoke.log 2>&1
rc=$?; tail -4 gpurun_out/r2e_smoke.log; echo "smoke rc=$rc"
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests/test_gpu_logical.py tests/test_gpu_large.py tests/test_gpu_uf.py tests/test_gpu_full_size.py -m gpu -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r2e_tests.log
T="timeout 300 python tools/fast_tune.py"
$T --d 11 --p 0.01 --tw 64 --adj 1 --bps 1,2 > gpurun_out/r2e_tune_d11.txt 2>&1; cat gpurun_out/r2e_tune_d11.txt
$T --d 11 --p 0.005 --baseline --tw '' > gpurun_out/r2e_tune_d11_p005.txt 2>&1; cat gpurun_out/r2e_tune_d11_p005.txt
$T --d 11 --p 0.02 --baseline --shots 1048576 --tw '' > gpurun_out/r2e_tune_d11_p02.txt 2>&1; cat gpurun_out/r2e_tune_d11_p02.txt
$T --d 11 --p 0.026 --baseline --shots 1048576 --tw 32,64 --adj 1 --bps 1 > gpurun_out/r2e_tune_d11_p026.txt 2>&1; cat gpurun_out/r2e_tune_d11_p026.txt
$T --d 11 --p 0.026 --flags 2 --shots 1048576 --tw '' > gpurun_out/r2e_tune_d11_p026_west.txt 2>&1; cat gpurun_out/r2e_tune_d11_p026_west.txt
$T --d 13 --p 0.026 --baseline --shots 262144 --tw '' > gpurun_out/r2e_tune_d13_p026.txt 2>&1; cat gpurun_out/r2e_tune_d13_p026.txt
$T --d 17 --p 0.026 --baseline --shots 65536 --tw '' > gpurun_out/r2e_tune_d17_p026.txt 2>&1; cat gpurun_out/r2e_tune_d17_p026.txt
$T --d 21 --p 0.026 --baseline --shots 32768 --tw '' > gpurun_out/r2e_tune_d21_p026.txt 2>&1; cat gpurun_out/r2e_tune_d21_p026.txt
$T --d 25 --p 0.026 --baseline --shots 16384 --tw '' > gpurun_out/r2e_tune_d25_p026.txt 2>&1; cat gpurun_out/r2e_tune_d25_p026.txt
$T --d 9 --noise 'code capacity' --p 0.05 --baseline --shots 16777216 --tw '' > gpurun_out/r2e_tune_cfg2.txt 2>&1; cat gpurun_out/r2e_tune_cfg2.txt
python bench.py --steps 3 --no-cpu --no-workloads --no-python-ref > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2e_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mc_uf_fast -c 1 -o gpurun_out/r2e_fast -f python bench.py --steps 1 --no-cpu --no-workloads --no-python-ref --shots-per-step 1048576 > gpurun_out/r2e_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mc_uf -c 2 -o gpurun_out/r2e_p026 -f python bench.py --p 0.026 --steps 1 --no-cpu --no-workloads --no-python-ref --shots-per-step 262144 > gpurun_out/r2e_ncu_p026.log 2>&1; echo "ncu p026 rc=$?"
