find oracle/_ref -name "*.pyc" | wc -l; ls -la oracle/_ref
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/z_ref.json 2> gpurun_out/z_ref.err; echo ref rc=$?; python -c "import json; d=json.loads(open('gpurun_out/z_ref.json').read().strip().splitlines()[-1]); print(d['cpu_baseline']['kind'], d['value'], d['cpu_baseline']['cores'])"; grep "bench\]" gpurun_out/z_ref.err
