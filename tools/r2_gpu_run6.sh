#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r2f_smoke.log 2>&1
rc=$?; tail -2 gpurun_out/r2f_smoke.log; echo "smoke rc=$rc"
if [ $rc -ne 0 ]; then exit 1; fi
( time python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err ) 2>&1 | tail -3; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
( time python bench.py --impl reference --steps 2 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err ) 2>&1 | tail -3; cut -c1-600 gpurun_out/r2f_bench_ref.json
timeout 300 python tools/sweep.py --ds 5,9 --seconds 0.2 --resume gpurun_out/r2f_sweep_resume.json --out gpurun_out/r2f_sweep > gpurun_out/r2f_sweep.log 2>&1; echo "sweep rc=$?"; tail -12 gpurun_out/r2f_sweep.log
timeout 120 python tools/sweep.py --ds 5,9 --seconds 0.2 --resume gpurun_out/r2f_sweep_resume.json --out gpurun_out/r2f_sweep > gpurun_out/r2f_sweep2.log 2>&1; echo "sweep resume rc=$?"; tail -4 gpurun_out/r2f_sweep2.log
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2f_tests.log
