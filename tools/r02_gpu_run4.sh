set -x
for rep in 1 2; do
for v in s2 s3; do echo "== $v" >> gpurun_out/g4_ab.txt; LD_LIBRARY_PATH=tools/ab/$v timeout 120 tests/native/selftest attn 2>&1 | grep -E "attn timing|FAIL" >> gpurun_out/g4_ab.txt; done
echo "== s4 (default)" >> gpurun_out/g4_ab.txt; timeout 120 tests/native/selftest attn 2>&1 | grep -E "attn timing|FAIL" >> gpurun_out/g4_ab.txt
done
cat gpurun_out/g4_ab.txt
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 6 -c 2 -o gpurun_out/g4_attn_bwd tests/native/selftest attn 90 > gpurun_out/g4_ncu_attn.log 2>&1; echo rc=$?
