cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 300 -k "tcgen05 or bf16x3" 2>&1 | tail -3
for a in 0 1; do echo "TRACE ARES=$a"; SMK_GEMM_ARES=$a timeout 200 python scripts/gemm_trace.py; done
for cfg in "0 0" "1 0" "0 1" "1 1"; do set -- $cfg
  echo "ARES=$1 PAIR=$2"; SMK_GEMM_ARES=$1 SMK_GEMM_CTA_PAIR=$2 timeout 200 python scripts/kernel_bench.py --only qkv,proj,fc1,fc2,kv | cut -c1-100; done
