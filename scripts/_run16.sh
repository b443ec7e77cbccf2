cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x 2>&1 | tail -3
for n in 1 2 4 8; do SMK_DEC_STREAMS=$n timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ds$n.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_ds$n.log | head -1; done
SMK_GEMM_CTA_PAIR=1 python scripts/kernel_bench.py --only proj,fc2 2>&1 | tail -2
SMK_GEMM_CTA_PAIR=1 SMK_GEMM_BN=192 python scripts/kernel_bench.py --only proj,fc2,qkv,fc1 2>&1 | tail -4
SMK_GEMM_CTA_PAIR=1 SMK_GEMM_BN=256 python scripts/kernel_bench.py --only qkv,fc1 2>&1 | tail -2
python scripts/step_timeline.py > gpurun_out/timeline_ds4.log 2>&1; tail -1 gpurun_out/timeline_ds4.log
