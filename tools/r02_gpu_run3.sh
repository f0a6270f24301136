set -x
python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/g3_tests.log 2>&1; tail -22 gpurun_out/g3_tests.log
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 20 -c 3 -o gpurun_out/g3_attn tests/native/selftest attn 90 > gpurun_out/g3_ncu_attn.log 2>&1; echo rc=$?
tail -3 gpurun_out/g3_ncu_attn.log
ls -la gpurun_out/
