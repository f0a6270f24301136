set -x
python -m pytest tests -m gpu -x -q > gpurun_out/g1_tests.log 2>&1; tail -3 gpurun_out/g1_tests.log
python bench.py --profile-kernels --timeline gpurun_out/g1_timeline.csv > gpurun_out/g1_bench.json 2> gpurun_out/g1_bench.err; echo rc=$?; cut -c1-300 gpurun_out/g1_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1500 --csv --log-file gpurun_out/g1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/g1_ncu_bench.log 2>&1; echo ncu rc=$?
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
