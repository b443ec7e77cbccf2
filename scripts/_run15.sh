cd /root/repo
timeout 600 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x -k "384" 2>&1 | tail -5
python scripts/step_timeline.py > gpurun_out/timeline.log 2>&1; tail -2 gpurun_out/timeline.log
python scripts/kernel_bench.py --only proj,fc2,qkv,fc1 2>&1 | tail -5
SMK_GEMM_BN=192 python scripts/kernel_bench.py --only proj,fc2,qkv,fc1 2>&1 | tail -5
python scripts/kernel_bench.py --only none --shapes "saqk:5120:768:1152:0:0,sav:5120:384:1152:0:0,out:5120:384:1152:0:1,l1:5120:1536:1152:2:0,l2:5120:384:4608:0:1,pe:50176:384:768:0:1" 2>&1 | tail -7
timeout 300 python bench.py --size 384 --batch 128 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_384.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_384.log 2>/dev/null | head -9; tail -2 gpurun_out/bench_384.log | cut -c1-300
