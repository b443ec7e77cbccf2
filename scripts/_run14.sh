cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 300 -x -k "attention_fa" 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x 2>&1 | tail -5
timeout 300 python bench.py --mode bf16x3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x3.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_x3.log 2>/dev/null | head -9
timeout 300 python bench.py --size 384 --batch 128 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_384.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_384.log 2>/dev/null | head -9; tail -2 gpurun_out/bench_384.log | cut -c1-400
