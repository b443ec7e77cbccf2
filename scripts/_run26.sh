cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x 2>&1 | tail -3
python scripts/kernel_bench.py --only none --shapes "x:1:384:384:0:0" > /dev/null 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_h.log 2>/dev/null | head -8
SMK_FUSE_LN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h0.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_h0.log 2>/dev/null | head -1
python scripts/step_timeline.py > gpurun_out/timeline_h.log 2>&1; tail -1 gpurun_out/timeline_h.log; sed -n 1,16p gpurun_out/timeline_h.log
grep -A8 '"bf16_nq20_224x224_B8"' gpurun_out/parity_report.json
