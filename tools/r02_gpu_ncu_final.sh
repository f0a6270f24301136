set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1500 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/f_ncu_bench.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwdp -s 40 -c 2 -o gpurun_out/f_attn_bwdp tests/native/selftest attn 90 > gpurun_out/f_ncu_attn1.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd1p -s 20 -c 1 -o gpurun_out/f_attn_fwd1p tests/native/selftest attn 90 > gpurun_out/f_ncu_attn2.log 2>&1; echo rc=$?
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:gemm_bf16_tcgen05_kernel<\(int\)256, \(int\)6, \(bool\)0, \(int\)2' -s 300 -c 59 -o gpurun_out/f_gemm_pair python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/f_ncu_gemm.log 2>&1; echo rc=$?
tail -3 gpurun_out/f_ncu_gemm.log
ncu --set full --clock-control none -k regex:layernorm_bwd_wide -s 30 -c 1 -o gpurun_out/f_ln_bwd python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/f_ncu_ln.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
