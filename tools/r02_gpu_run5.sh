set -x
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 40 -c 2 -o gpurun_out/g5_attn_bwd tests/native/selftest attn 90 > gpurun_out/g5_ncu_attn.log 2>&1; echo rc=$?
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:gemm_bf16_tcgen05_kernel<256, 6, 0, 2' -s 300 -c 59 -o gpurun_out/g5_gemm_pair python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/g5_ncu_gemm.log 2>&1; echo rc=$?
tail -5 gpurun_out/g5_ncu_gemm.log
ls -la gpurun_out
