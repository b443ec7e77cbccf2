cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 -x 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_d.log 2>/dev/null | head -8
python scripts/step_timeline.py > gpurun_out/timeline_d.log 2>&1; tail -1 gpurun_out/timeline_d.log; sed -n 1,12p gpurun_out/timeline_d.log; sed -n 86,100p gpurun_out/timeline_d.log
