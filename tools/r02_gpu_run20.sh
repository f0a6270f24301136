for env in "X=1" "TVS_GEMM_EPILOGUE=s" "TVS_GEMM_EPI=generic"; do echo "== $env"; env $env python tools/gemm_shape_bench.py 768 768 res 2>&1 | tail -1; env $env python tools/gemm_shape_bench.py 768 768 res 256 2>&1 | tail -1; done
python tools/gemm_shape_bench.py 768 768 res 64 2>&1 | tail -1
