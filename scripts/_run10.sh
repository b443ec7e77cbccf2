cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q --timeout 600 -x 2>&1 | tail -3
for pdl in 0 1; do SMK_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_pdl$pdl.log | head -1; done
