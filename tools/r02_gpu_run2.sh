set -x
tools/microbench/tmem_bw > gpurun_out/g2_tmem_bw.txt 2>&1; tail -40 gpurun_out/g2_tmem_bw.txt
python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/g2_tests.log 2>&1; tail -25 gpurun_out/g2_tests.log
