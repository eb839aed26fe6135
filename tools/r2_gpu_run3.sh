#!/bin/bash
# Round-2 GPU call 3: smoke first (abort on a hang), then parity tests and the launch-shape sweeps.
mkdir -p gpurun_out
timeout 180 python -c "
import __graft_entry__ as g
g.smoke()
import localuf_b200 as L
from localuf_b200 import _engine
eng = _engine.Engine(L.Surface(11, 'phenomenological'), device=0)
print('cfg3 fast', eng.mc_run_host(0, 0, 0.01, 1, 0, 1 << 20))
print('cfg3 west', eng.mc_run_host(0, 2, 0.026, 1, 0, 1 << 18))
" > gpurun_out/r2c_smoke.log 2>&1
rc=$?; tail -5 gpurun_out/r2c_smoke.log; echo "smoke rc=$rc"
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests/test_gpu_logical.py tests/test_gpu_uf.py tests/test_gpu_full_size.py tests/test_global_scheme.py tests/test_special.py -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2c_tests.log
timeout 500 python tools/fast_tune.py --d 11 --p 0.01 --baseline > gpurun_out/r2c_tune_d11_p01.txt 2>&1; cat gpurun_out/r2c_tune_d11_p01.txt
timeout 300 python tools/fast_tune.py --d 11 --p 0.026 --baseline --shots 1048576 --tw 16,32 > gpurun_out/r2c_tune_d11_p026.txt 2>&1; cat gpurun_out/r2c_tune_d11_p026.txt
timeout 200 python tools/fast_tune.py --d 13 --p 0.015 --baseline --shots 1048576 --tw 16,32 --bps 1 > gpurun_out/r2c_tune_d13.txt 2>&1; cat gpurun_out/r2c_tune_d13.txt
timeout 200 python tools/fast_tune.py --d 17 --p 0.015 --baseline --shots 262144 --tw 32,64,128 --adj 0 --bps 1 > gpurun_out/r2c_tune_d17.txt 2>&1; cat gpurun_out/r2c_tune_d17.txt
timeout 200 python tools/fast_tune.py --d 17 --p 0.026 --baseline --shots 65536 --tw 32,64,128 --adj 0 --bps 1 > gpurun_out/r2c_tune_d17_p026.txt 2>&1; cat gpurun_out/r2c_tune_d17_p026.txt
timeout 200 python tools/fast_tune.py --d 25 --p 0.015 --baseline --shots 65536 --tw 64,128,256 --adj 0 --bps 1,2 > gpurun_out/r2c_tune_d25.txt 2>&1; cat gpurun_out/r2c_tune_d25.txt
