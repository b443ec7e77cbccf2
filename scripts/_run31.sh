cd /root/repo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ln2 -s 5 -c 1 -f -o gpurun_out/gl2 python scripts/gemm_ln_bench.py > gpurun_out/ncu_gl2.log 2>&1; tail -2 gpurun_out/ncu_gl2.log
