cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"query_iou_cells|mask_metrics_x4|mask_head" -c 3 -o gpurun_out/r01_tail2 -f python scripts/step_timeline.py > gpurun_out/ncu_tail2.log 2>&1; tail -3 gpurun_out/ncu_tail2.log
