# compute-sanitizer over the native self-test (race / memory checking of the hand-rolled mbarrier / TMEM protocols)
CS=/usr/local/cuda/bin/compute-sanitizer
rm -f gpurun_out/sanitizer.txt
for tool in memcheck synccheck; do
  for what in attncheck ln loss ffn; do
    echo "==== compute-sanitizer --tool $tool selftest $what" >> gpurun_out/sanitizer.txt
    timeout 600 $CS --tool $tool --print-limit 20 tests/native/selftest $what 900 > gpurun_out/san_raw.txt 2>&1
    echo "exit $?" >> gpurun_out/sanitizer.txt
    grep -E "ERROR SUMMARY|Error|error|FAIL|PASSED|FAILED|=========     at|Invalid|hazard|COMPUTE-SANITIZER" gpurun_out/san_raw.txt | head -30 >> gpurun_out/sanitizer.txt
    [ -s gpurun_out/san_raw.txt ] || echo "(no output)" >> gpurun_out/sanitizer.txt
  done
done
echo "==== compute-sanitizer --tool racecheck selftest attncheck (mbarrier / async-proxy traffic is outside racecheck's model: hazards listed here are reviewed by hand)" >> gpurun_out/sanitizer.txt
timeout 600 $CS --tool racecheck --print-limit 20 tests/native/selftest attncheck 900 > gpurun_out/san_raw.txt 2>&1; echo "exit $?" >> gpurun_out/sanitizer.txt
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard|FAIL|PASSED|FAILED" gpurun_out/san_raw.txt | head -40 >> gpurun_out/sanitizer.txt
tail -5 gpurun_out/san_raw.txt >> gpurun_out/sanitizer.txt
cat gpurun_out/sanitizer.txt
# the persistent forward with eight softmax warps: correctness + timing
timeout 120 tests/native/selftest attn > gpurun_out/g14_attn.txt 2>&1; grep -E "FAIL|timing|PASSED|FAILED" gpurun_out/g14_attn.txt | head
