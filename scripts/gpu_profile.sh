#!/bin/bash
# ncu evidence for profiles/: launch list of one bench run + full-set captures of the top kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -c 600 gpurun_out/plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1100 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 30 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:attn_tc|query_iou|mask_metrics|mask_head|layernorm" -s 40 -c 8 -o gpurun_out/prof_misc $CMD > gpurun_out/ncu_misc.log 2>&1
echo "misc rc=$?"
ls -la gpurun_out/*.ncu-rep
