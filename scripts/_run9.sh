cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x 2>&1 | tail -3
timeout 100 python scripts/kernel_bench.py --only qkv,proj,fc1,fc2,kv | cut -c1-100
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_a.log
