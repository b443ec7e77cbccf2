cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 -x -k "finalisation or metrics or eval" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_c.log | head -1
