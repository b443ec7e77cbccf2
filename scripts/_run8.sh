cd /root/repo
for d in 0 8 16 1; do echo "DBG=$d (8=nostore 16=stage-but-no-TMA 1=noepi)"; SMK_GEMM_DEBUG=$d SMK_GEMM_ARES=0 timeout 100 python scripts/kernel_bench.py --only qkv,fc1,kv | cut -c1-100; done
