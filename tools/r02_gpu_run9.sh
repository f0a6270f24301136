timeout 90 tests/native/selftest attn > gpurun_out/g9_attn_new.txt 2>&1; echo rc=$?
grep -E "FAIL|timing|PASSED|FAILED" gpurun_out/g9_attn_new.txt | head -20
LD_LIBRARY_PATH=tools/ab/e9 timeout 90 tests/native/selftest attn 2>&1 | grep -E "timing B=32 S=489 H=12"
timeout 90 tests/native/selftest attn 2>&1 | grep -E "timing B=32 S=489 H=12"
