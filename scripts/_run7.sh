cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_u8.log 2>&1; tail -3 gpurun_out/bench_u8.log | cut -c1-600; python scripts/summarize_bench.py gpurun_out/bench_u8.log
