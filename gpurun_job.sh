mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fast.py -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_fast.log
(COLTT_DEBUG_PROF=1 timeout 300 python bench.py --steps 8 --warmup 3 --math fast --no-cpu 2>&1 | grep -E "coltt prof|ms_per_step" | cut -c1-300) > gpurun_out/prof_roles.log
(timeout 300 python bench.py --steps 30 --warmup 5 --math fast --no-cpu 2>&1 | tail -2 | cut -c1-1600) > gpurun_out/bench_fast.log
tail -n 4 gpurun_out/pytest_fast.log; cat gpurun_out/prof_roles.log; cat gpurun_out/bench_fast.log
