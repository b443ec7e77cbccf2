cd /root/repo
for cfg in "0 1 0" "1 1 0" "1 1 256"; do set -- $cfg
  echo "TRACE ARES=$1 PAIR=$2 BN=$3"; if [ $3 = 0 ]; then unset SMK_GEMM_BN; else export SMK_GEMM_BN=$3; fi
  SMK_GEMM_ARES=$1 SMK_GEMM_CTA_PAIR=$2 timeout 200 python scripts/gemm_trace.py; done
