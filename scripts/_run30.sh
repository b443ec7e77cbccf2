cd /root/repo
timeout 300 python -m pytest tests/test_gpu_kernels.py -q --timeout 120 -x -k "gemm_layernorm" 2>&1 | tail -8
timeout 120 python scripts/gemm_ln_bench.py 2>&1 | tail -6
