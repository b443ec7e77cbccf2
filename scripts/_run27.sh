cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 -x -k "layernorm" 2>&1 | tail -2
timeout 300 python bench.py --mode bf16x3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x3.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_x3.log 2>/dev/null | head -9
grep -A10 '"bf16x3_nq20_224x224_B8"' gpurun_out/parity_report.json; grep "evaluator_bf16x3" gpurun_out/parity_report.json
