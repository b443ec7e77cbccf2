#!/bin/bash
# ncu evidence for profiles/: launch list of exactly one bench step + full-set captures of one launch of every
# distinct kernel / shape of the step.  `bench.py --ncu-step` brackets ONE step with cudaProfilerStart/Stop, so with
# --profile-from-start off the launch indices are the step's own (bf16 mode, 224x224, nq 20, 166 launches):
#   0 im2col, 1 patch-embed GEMM, 2 cls/pos rows, 3+7i.. encoder layer i = LN qkv attn proj LN fc1 fc2,
#   87 final LN, 88 memory K/V GEMM, 89 layer-0 state tile, 90-96 decoder layer 0 (cross-attention block + FFN),
#   97+12(l-1).. decoder layer l >= 1, 157 mask logits, 158 mask upsample, 159-163 objectness / features, 164-165 eval
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${TAG:-prof}
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -c 300 gpurun_out/plain.log
CMD="python bench.py --ncu-step --warmup 3"
$CMD > gpurun_out/plain_step.log 2>&1 || { echo "plain step run failed"; tail -5 gpurun_out/plain_step.log; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
cap() { name=$1; skip=$2; count=$3
  ncu --profile-from-start off --set full --clock-control none --import-source on -s $skip -c $count -f -o gpurun_out/${TAG}_$name $CMD > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"; }
cap enc 10 7
cap kv 87 2
cap dec ${DEC_SKIP:-96} 12
cap tail ${TAIL_SKIP:-156} 8
ls -la gpurun_out/*.ncu-rep
# summaries are produced ON the box (gpurun pulls at most 64 MiB back): tables + traffic JSON, then only the encoder-layer
# report (source-level view of the GEMM / attention / LayerNorm kernels) travels home
python scripts/ncu_summary.py launches gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_table.md
python scripts/ncu_summary.py full gpurun_out/${TAG}_enc.ncu-rep gpurun_out/${TAG}_kv.ncu-rep gpurun_out/${TAG}_dec.ncu-rep gpurun_out/${TAG}_tail.ncu-rep > gpurun_out/${TAG}_full_tables.md
python scripts/ncu_summary.py traffic gpurun_out/${TAG}_enc.ncu-rep gpurun_out/${TAG}_kv.ncu-rep gpurun_out/${TAG}_dec.ncu-rep gpurun_out/${TAG}_tail.ncu-rep > gpurun_out/${TAG}_traffic.json
rm -f gpurun_out/${TAG}_kv.ncu-rep gpurun_out/${TAG}_dec.ncu-rep gpurun_out/${TAG}_tail.ncu-rep
ls -la gpurun_out/ | tail -12
