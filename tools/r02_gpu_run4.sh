set -x
for v in s2 s3; do LD_LIBRARY_PATH=tools/ab/$v timeout 120 tests/native/selftest attn 2>&1 | grep -E "attn timing|FAIL|PASS" ; done
timeout 120 tests/native/selftest attn 2>&1 | grep -E "attn timing|FAIL|PASS"
for v in s2 s3; do LD_LIBRARY_PATH=tools/ab/$v timeout 120 tests/native/selftest attn 2>&1 | grep -E "attn timing|FAIL|PASS" ; done
timeout 120 tests/native/selftest attn 2>&1 | grep -E "attn timing|FAIL|PASS"
