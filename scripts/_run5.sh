cd /root/repo
for cfg in "0 0 3" "0 0 1" "0 0 2" "1 0 3" "1 0 1" "0 1 3" "0 1 1" "1 1 3" "1 1 1"; do set -- $cfg
  echo "ARES=$1 PAIR=$2 DBG=$3 (1=noepi 2=noload)"; SMK_GEMM_DEBUG=$3 SMK_GEMM_ARES=$1 SMK_GEMM_CTA_PAIR=$2 timeout 100 python scripts/kernel_bench.py --only qkv,proj,fc1,fc2,kv | cut -c1-100; done
