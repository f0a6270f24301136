timeout 90 tests/native/selftest attn > gpurun_out/g11_attn.txt 2>&1; echo rc=$?
grep -E "FAIL|timing|PASSED|FAILED" gpurun_out/g11_attn.txt | head -20
TVS_ATTN_FWD=3 timeout 90 tests/native/selftest attn 2>&1 | grep -E "timing B=32 S=489 H=12"
timeout 90 tests/native/selftest attn 2>&1 | grep -E "timing B=32 S=489 H=12"
