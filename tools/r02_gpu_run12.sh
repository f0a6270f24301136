timeout 120 tests/native/selftest gemm > gpurun_out/g12_gemm.txt 2>&1; echo rc=$?; grep -E "FAIL|PASSED|FAILED" gpurun_out/g12_gemm.txt | head
timeout 300 python -m pytest tests/test_gpu_bench_config.py -x -q -k "tower_gemm" > gpurun_out/g12_pytest.txt 2>&1; tail -5 gpurun_out/g12_pytest.txt
echo "== stream-K on" > gpurun_out/g12_prof.txt; timeout 120 tests/native/selftest gemmprof 90 256 >> gpurun_out/g12_prof.txt 2>&1
echo "== stream-K off" >> gpurun_out/g12_prof.txt; TVS_GEMM_STREAMK=0 timeout 120 tests/native/selftest gemmprof 90 256 >> gpurun_out/g12_prof.txt 2>&1
grep -E "^==|^--|us|FAIL" gpurun_out/g12_prof.txt | head -90
