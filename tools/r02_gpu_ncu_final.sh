set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1500 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/f_ncu_bench.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwdp -s 40 -c 2 -o gpurun_out/f_attn_bwdp tests/native/selftest attn 90 > gpurun_out/f_ncu_attn1.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd1p -s 20 -c 1 -o gpurun_out/f_attn_fwd1p tests/native/selftest attn 90 > gpurun_out/f_ncu_attn2.log 2>&1; echo rc=$?
ncu --set full --clock-control none -k regex:gemm_bf16 -s 100 -c 1 -o gpurun_out/f_gemm_fc1 tests/native/selftest gemmprof 90 256 > gpurun_out/f_ncu_g1.log 2>&1; echo rc=$?
ncu --set full --clock-control none -k regex:gemm_bf16 -s 124 -c 1 -o gpurun_out/f_gemm_dgelu tests/native/selftest gemmprof 90 256 > gpurun_out/f_ncu_g2.log 2>&1; echo rc=$?
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_bf16 --csv --log-file gpurun_out/f_gemm_traffic.csv tests/native/selftest gemmprof 300 256 > gpurun_out/f_ncu_g3.log 2>&1; echo rc=$?
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:layernorm --csv --log-file gpurun_out/f_ln_traffic.csv tests/native/selftest lnprof 300 > gpurun_out/f_ncu_l1.log 2>&1; echo rc=$?
ncu --set full --clock-control none -k regex:layernorm_bwd_wide -s 5 -c 1 -o gpurun_out/f_ln_bwd tests/native/selftest lnprof 90 > gpurun_out/f_ncu_l2.log 2>&1; echo rc=$?
for v in 1 4; do for c in 2 3 4 6; do echo "== TVS_LN_FWD=$v CTAS=$c" >> gpurun_out/f_ln_ab.txt; TVS_LN_FWD=$v TVS_LN_PIPE_CTAS=$c timeout 100 tests/native/selftest lnprof 2>&1 | grep -E "FAIL|layernorm fwd|PASSED|FAILED" >> gpurun_out/f_ln_ab.txt; [ $v = 1 ] && break; done; done
cat gpurun_out/f_ln_ab.txt
du -sh gpurun_out; ls -la gpurun_out
