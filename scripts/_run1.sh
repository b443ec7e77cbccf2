cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 300 -k "tcgen05 or bf16x3" 2>&1 | tail -5
for cfg in "0 0" "1 0" "1 1"; do set -- $cfg; echo "ARES=$1 PAIR=$2"; SMK_GEMM_ARES=$1 SMK_GEMM_CTA_PAIR=$2 timeout 200 python scripts/kernel_bench.py --only qkv,proj,fc1,fc2,kv; done
echo "ARES=1 BN=128"; SMK_GEMM_BN=128 timeout 200 python scripts/kernel_bench.py --only qkv,proj,fc1,kv
echo "ARES=1 BN=192"; SMK_GEMM_BN=192 timeout 200 python scripts/kernel_bench.py --only qkv,proj,fc1,kv
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ares.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_ares.log
