cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -q --timeout 600 -x -k "evaluator or pipeline" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample 16 > gpurun_out/bench_i.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_i.log 2>/dev/null | head -1; tail -1 gpurun_out/bench_i.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','cpu_baseline','parity_mode') if k in d})"
