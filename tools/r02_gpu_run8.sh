timeout 60 tools/microbench/umma_rate > gpurun_out/g8_umma_rate.txt 2>&1; echo rc=$?; cat gpurun_out/g8_umma_rate.txt
