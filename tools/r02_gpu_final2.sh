python -m pytest tests -m gpu -x -q > gpurun_out/y_tests.log 2>&1; tail -3 gpurun_out/y_tests.log
python bench.py --profile-kernels --timeline gpurun_out/y_timeline.csv > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err; echo bench rc=$?; cut -c1-200 gpurun_out/y_bench.json
