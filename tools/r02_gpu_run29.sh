for rep in 1 2 3; do for env in "TVS_LN_FWD=1" "TVS_LN_FWD=0"; do
  echo -n "$env  "; env $env python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done; done
