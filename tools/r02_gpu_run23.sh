for e in fc1 fc1grad dqgelu mulaux plain; do python tools/gemm_shape_bench.py 3072 768 $e 2>&1 | tail -1; done
timeout 300 python -m pytest tests/test_gpu_bench_config.py -x -q -k "tower_gemm" 2>&1 | tail -3
for rep in 1 2; do for env in "TVS_PRE_DGELU=0" "TVS_PRE_DGELU=1"; do
  echo -n "$env  "; env $env python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done; done
