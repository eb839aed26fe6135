timeout 800 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for wl in cfg3 cfg3-staged cfg2; do timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu 2>/dev/null > gpurun_out/bench_$wl.json; cut -c1-120 gpurun_out/bench_$wl.json; done
timeout 300 python tools/sweep_threshold.py --ds 9,13,17,25 --ps 0.015,0.026 --seconds 0.5 | tail -9
