timeout 150 tests/native/selftest all 140 > gpurun_out/z_selftest.txt 2>&1; echo selftest rc=$?; grep -E "FAIL|PASSED|FAILED|WATCHDOG" gpurun_out/z_selftest.txt | head
for g in loss ln attn gemm ffn; do timeout 120 tests/native/selftest $g 110 2>&1 | grep -E "FAIL|PASSED|FAILED|WATCHDOG" | head -3; done
python -m pytest tests -m gpu -x -q > gpurun_out/z_tests.log 2>&1; tail -3 gpurun_out/z_tests.log
