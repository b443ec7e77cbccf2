cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x -k "bf16x3" 2>&1 | tail -12
timeout 300 python bench.py --mode bf16x3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x3.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_x3.log 2>/dev/null | head -9; tail -3 gpurun_out/bench_x3.log | cut -c1-300
grep -A12 bf16x3 gpurun_out/parity_report.json | head -60
