cd /root/repo
SMK_GEMM_CTA_PAIR=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 -x -k "tcgen05_matches" 2>&1 | tail -3
SMK_GEMM_CTA_PAIR=1 python scripts/kernel_bench.py --only fc1,kv,qkv 2>&1 | tail -3
SMK_GEMM_CTA_PAIR=1 SMK_GEMM_BN=256 python scripts/kernel_bench.py --only fc1,kv 2>&1 | tail -2
