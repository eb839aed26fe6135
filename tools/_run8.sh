timeout 600 python -m pytest tests/test_gpu_uf.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | cut -c1-120
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | cut -c1-120
timeout 300 python tools/sweep_threshold.py --ds 9,13,17 --ps 0.015,0.026 --seconds 0.5 | tail -7
