for env in "X=1" "TVS_TEXT_STREAM=0" "TVS_PDL=1" "TVS_MAIN_PRIORITY=1"; do
  echo "== $env"; env $env python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
