python -m pytest tests -m gpu -x -q > gpurun_out/g13_tests.log 2>&1; tail -3 gpurun_out/g13_tests.log
python bench.py --profile-kernels --timeline gpurun_out/g13_timeline.csv > gpurun_out/g13_bench.json 2> gpurun_out/g13_bench.err; echo rc=$?; cut -c1-250 gpurun_out/g13_bench.json
