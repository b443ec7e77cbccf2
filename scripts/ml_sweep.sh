S="ml:37888:512:4096:0:0,mlk1:37888:1024:1024:0:0"
for pair in 0 1; do for bn in 128 256; do echo "PAIR=$pair BN=$bn"; SMK_GEMM_CTA_PAIR=$pair SMK_GEMM_BN=$bn timeout 100 python scripts/kernel_bench.py --shapes $S 2>&1 | tail -2; done; done
