cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 -x -k "tcgen05 or bf16x3" 2>&1 | tail -3
SMK_GEMM_EPI16=0 python scripts/kernel_bench.py --only fc1,kv,qkv 2>&1 | tail -3
python scripts/kernel_bench.py --only fc1,kv,qkv 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_e.log 2>/dev/null | head -8
