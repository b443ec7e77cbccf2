cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 -x -k "layernorm" 2>&1 | tail -2
python scripts/kernel_bench.py 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_g.log 2>/dev/null | head -3
