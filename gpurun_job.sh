mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/pytest_gpu.log
(timeout 300 python bench.py --steps 20 --warmup 5 --math fast 2>&1 | tail -5) > gpurun_out/bench_fast.log
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_filter|rerank|flat_scan|merge_topk" -c 40 --csv --log-file gpurun_out/launches_fast.csv python bench.py --steps 3 --warmup 2 --math fast --no-cpu > gpurun_out/ncu_bench.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_filter -s 2 -c 1 -f -o gpurun_out/prof_gemm python bench.py --steps 2 --warmup 1 --math fast --no-cpu > gpurun_out/ncu_full.log 2>&1)
tail -n 12 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/bench_fast.log
