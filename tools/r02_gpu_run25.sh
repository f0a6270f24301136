timeout 300 python tools/kernel_bench.py > gpurun_out/v_hbm_kernels.txt 2>&1; tail -30 gpurun_out/v_hbm_kernels.txt
