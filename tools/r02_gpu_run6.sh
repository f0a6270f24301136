rm -f gpurun_out/g6_ab.txt
for rep in 1 2; do
echo "== default" >> gpurun_out/g6_ab.txt; timeout 120 tests/native/selftest attn 2>&1 | grep -E "attn timing B=32 S=489 H=12" >> gpurun_out/g6_ab.txt
for v in e3 e5 e6 e7; do echo "== $v" >> gpurun_out/g6_ab.txt; LD_LIBRARY_PATH=tools/ab/$v timeout 120 tests/native/selftest attn 2>&1 | grep -E "attn timing B=32 S=489 H=12" >> gpurun_out/g6_ab.txt; done
done
cat gpurun_out/g6_ab.txt
