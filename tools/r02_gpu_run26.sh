timeout 300 python tools/kernel_bench.py 2>&1 | grep -E "film_bwd"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "small_geometry or full_geometry" 2>&1 | tail -2
