cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_f.log 2>/dev/null | head -8
python scripts/step_timeline.py > gpurun_out/timeline_f.log 2>&1; tail -10 gpurun_out/timeline_f.log
grep -A6 "vits8" gpurun_out/parity_report.json | head -20
