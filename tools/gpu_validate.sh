set -x
python -m pytest tests -m gpu -x -q > gpurun_out/f1_tests.log 2>&1; tail -2 gpurun_out/f1_tests.log
timeout 300 tests/native/selftest all 250 > gpurun_out/f1_selftest.log 2>&1; tail -2 gpurun_out/f1_selftest.log
python bench.py --profile-kernels > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err; echo rc=$?; cut -c1-200 gpurun_out/f1_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f1_ref.json 2> gpurun_out/f1_ref.err; echo rc=$?; cut -c1-300 gpurun_out/f1_ref.json
python bench.py --workload cris_cocoop --no-cpu-baseline --profile-kernels > gpurun_out/f1_cris.json 2> gpurun_out/f1_cris.err; echo rc=$?; cut -c1-200 gpurun_out/f1_cris.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 2000 --csv --log-file gpurun_out/f1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f1_ncu_bench.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o gpurun_out/f1_gemm tests/native/selftest gemmprof 90 256 > gpurun_out/f1_ncu_gemm.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:ffn_tc -s 8 -c 2 -o gpurun_out/f1_ffn tests/native/selftest ffn 90 > gpurun_out/f1_ncu_ffn.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 20 -c 3 -o gpurun_out/f1_attn tests/native/selftest attn 90 > gpurun_out/f1_ncu_attn.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
