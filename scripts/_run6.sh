cd /root/repo
for cfg in "0 0 3 0" "0 0 7 0" "0 0 7 128" "0 0 7 256" "0 1 7 256"; do set -- $cfg
  if [ $4 = 0 ]; then unset SMK_GEMM_BN; else export SMK_GEMM_BN=$4; fi
  echo "ARES=$1 PAIR=$2 DBG=$3 BN=$4"; SMK_GEMM_DEBUG=$3 SMK_GEMM_ARES=$1 SMK_GEMM_CTA_PAIR=$2 timeout 100 python scripts/kernel_bench.py --only qkv,proj,fc1,fc2,kv --shapes big:37888:2048:4096:0:0| cut -c1-100; done
