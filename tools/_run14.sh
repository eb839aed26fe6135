timeout 600 python -m pytest tests/test_gpu_uf.py tests/test_gpu_full_size.py tests/test_node_uf.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | cut -c1-120
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | cut -c1-120
