#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -c "
import __graft_entry__ as g
g.smoke()
import localuf_b200 as L
from localuf_b200 import _engine
eng = _engine.Engine(L.Surface(11, 'phenomenological'), device=0)
print('cfg3 fast', eng.mc_run_host(0, 0, 0.01, 1, 0, 1 << 20))
print('cfg3 west', eng.mc_run_host(0, 2, 0.026, 1, 0, 1 << 18))
eng = _engine.Engine(L.Surface(17, 'phenomenological'), device=0)
print('d17', eng.mc_run_host(0, 0, 0.02, 1, 0, 1 << 14))
eng = _engine.Engine(L.Surface(23, 'circuit-level'), device=0)
print('d23 cl west', eng.mc_run_host(0, 2, 0.004, 1, 0, 1 << 12))
" > gpurun_out/r2d_smoke.log 2>&1
rc=$?; tail -6 gpurun_out/r2d_smoke.log; echo "smoke rc=$rc"
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests/test_gpu_logical.py tests/test_gpu_large.py tests/test_gpu_uf.py tests/test_gpu_full_size.py tests/test_global_scheme.py tests/test_special.py -m gpu -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2d_tests.log
T="timeout 300 python tools/fast_tune.py"
$T --d 11 --p 0.01 --tw '' > gpurun_out/r2d_tune_d11_default.txt 2>&1; cat gpurun_out/r2d_tune_d11_default.txt
LUF_FAST_CAP_SCALE=0.6 $T --d 11 --p 0.01 --tw 32 --adj 1 --bps 1,2 > gpurun_out/r2d_tune_d11_caps06.txt 2>&1; cat gpurun_out/r2d_tune_d11_caps06.txt
LUF_FAST_CAP_SCALE=0.4 $T --d 11 --p 0.01 --tw 32 --adj 1 --bps 1,2 > gpurun_out/r2d_tune_d11_caps04.txt 2>&1; cat gpurun_out/r2d_tune_d11_caps04.txt
$T --d 11 --p 0.01 --tw 32 --adj 0 --bps 2 --spb 16,18 > gpurun_out/r2d_tune_d11_spb.txt 2>&1; cat gpurun_out/r2d_tune_d11_spb.txt
$T --d 13 --p 0.015 --shots 1048576 --tw 32 --adj 0,1 --bps 1 --spb 0,16,20 > gpurun_out/r2d_tune_d13.txt 2>&1; cat gpurun_out/r2d_tune_d13.txt
$T --d 17 --p 0.015 --baseline --shots 262144 --tw 32,64,128 --adj 0 --bps 1 > gpurun_out/r2d_tune_d17.txt 2>&1; cat gpurun_out/r2d_tune_d17.txt
$T --d 17 --p 0.026 --baseline --shots 65536 --tw 32,64,128 --adj 0 --bps 1 > gpurun_out/r2d_tune_d17_p026.txt 2>&1; cat gpurun_out/r2d_tune_d17_p026.txt
$T --d 25 --p 0.015 --baseline --shots 65536 --tw 64,128,256 --adj 0 --bps 1,2 > gpurun_out/r2d_tune_d25.txt 2>&1; cat gpurun_out/r2d_tune_d25.txt
$T --d 25 --p 0.026 --baseline --shots 16384 --tw 128,256 --adj 0 --bps 1 > gpurun_out/r2d_tune_d25_p026.txt 2>&1; cat gpurun_out/r2d_tune_d25_p026.txt
$T --d 25 --noise circuit-level --flags 2 --p 0.004 --shots 16384 --tw 128,256 --adj 0 --bps 1 > gpurun_out/r2d_tune_d25cl.txt 2>&1; cat gpurun_out/r2d_tune_d25cl.txt
python bench.py --steps 3 --no-cpu --no-workloads --no-python-ref > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2d_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mc_uf_fast -c 1 -o gpurun_out/r2d_fast -f python bench.py --steps 1 --no-cpu --no-workloads --no-python-ref --shots-per-step 1048576 > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
