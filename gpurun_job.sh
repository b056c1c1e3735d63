mkdir -p gpurun_out
(timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_filter|rerank|flat_scan|merge_topk' -c 28 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_launch_final.log 2>&1); wc -l gpurun_out/launches_final.csv
