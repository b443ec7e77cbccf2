cd /root/repo
for cfg in "0 1" "0 4" "1 4"; do set -- $cfg; SMK_PDL=$1 SMK_DEC_STREAMS=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p$1_ds$2.log 2>&1; python scripts/summarize_bench.py gpurun_out/bench_p$1_ds$2.log 2>/dev/null | head -1; done
SMK_PDL=0 SMK_DEC_STREAMS=4 python scripts/step_timeline.py > gpurun_out/timeline_p0_ds4.log 2>&1; tail -1 gpurun_out/timeline_p0_ds4.log
