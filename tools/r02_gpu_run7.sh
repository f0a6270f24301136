LD_LIBRARY_PATH=tools/ab/e8 timeout 60 tools/microbench/attn_timeline > gpurun_out/g7_timeline.txt 2>&1; echo rc=$?; head -5 gpurun_out/g7_timeline.txt
