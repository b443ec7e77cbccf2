cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 -x -k "eval or query_iou or metrics" 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cells.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_cells.log 2>/dev/null | head -8
python scripts/step_timeline.py 2>&1 | tail -6
