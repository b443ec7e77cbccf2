cd /root/repo
TAG=r01b bash scripts/gpu_profile.sh > gpurun_out/profile_run.log 2>&1
tail -8 gpurun_out/profile_run.log
python scripts/ncu_summary.py launches gpurun_out/r01b_launches.csv > gpurun_out/r01b_launch_table.md 2>&1
python scripts/ncu_summary.py full gpurun_out/r01b_enc.ncu-rep gpurun_out/r01b_kv.ncu-rep gpurun_out/r01b_dec.ncu-rep gpurun_out/r01b_tail.ncu-rep > gpurun_out/r01b_full_tables.md 2>&1
python scripts/ncu_summary.py traffic gpurun_out/r01b_enc.ncu-rep gpurun_out/r01b_kv.ncu-rep gpurun_out/r01b_dec.ncu-rep gpurun_out/r01b_tail.ncu-rep > gpurun_out/r01b_traffic.json 2>&1
rm -f gpurun_out/r01b_dec.ncu-rep gpurun_out/r01b_kv.ncu-rep
ls -la gpurun_out | head -30
