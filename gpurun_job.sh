mkdir -p gpurun_out
(timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_filter|rerank|flat_scan|merge_topk|prep_rows' -s 7 -c 30 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_launch_final.log 2>&1); tail -n 2 gpurun_out/ncu_launch_final.log | cut -c1-300; wc -l gpurun_out/launches_final.csv
