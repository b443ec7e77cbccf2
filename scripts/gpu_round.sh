#!/bin/bash
# One GPU trip: unit tests (tcgen05 ones in their own process so a trapped launch cannot poison the rest),
# model parity, bench.  Everything is logged under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -${TAILN:-6} gpurun_out/$name.log; }
run t_kernels python -m pytest tests/test_gpu_kernels.py -q --timeout 300 -k "not tcgen05 and not bf16x3"
run t_tcgemm  python -m pytest tests/test_gpu_kernels.py -q --timeout 300 -k "tcgen05 or bf16x3"
run t_model32 python -m pytest tests/test_gpu_model.py -q --timeout 600 -k "not bf16"
run t_modelbf python -m pytest tests/test_gpu_model.py -q --timeout 600 -k "bf16"
run smoke python __graft_entry__.py smoke
TAILN=3 run bench_fp32 python bench.py --mode fp32 --batch 64 --steps 3 --warmup 3 --no-cpu-baseline
TAILN=3 run bench_bf16 python bench.py --steps 5 --warmup 3 --cpu-sample 16
cat gpurun_out/parity_report.json 2>/dev/null | head -80
