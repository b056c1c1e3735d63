mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_fast.py -m gpu -q -x 2>&1 | tail -30) > gpurun_out/pytest_fast.log
(COLTT_DEBUG_PROF=1 timeout 200 python bench.py --steps 8 --warmup 3 --math fast --no-cpu 2>&1 | grep -E "coltt prof|ms_per_step" | cut -c1-160) > gpurun_out/prof_pair.log
(COLTT_DEBUG_PROF=1 COLTT_DEBUG_FLAGS=1 timeout 200 python bench.py --steps 8 --warmup 3 --math fast --no-cpu 2>&1 | grep -E "mma_|prod_|ms_per_step" | cut -c1-160) > gpurun_out/prof_pair_noepi.log
(COLTT_FAST_PAIR=0 COLTT_DEBUG_PROF=1 timeout 200 python bench.py --steps 8 --warmup 3 --math fast --no-cpu 2>&1 | grep -E "coltt prof|ms_per_step" | cut -c1-160) > gpurun_out/prof_single.log
tail -n 12 gpurun_out/pytest_fast.log; cat gpurun_out/prof_pair.log; echo; cat gpurun_out/prof_pair_noepi.log; echo ==== single; cat gpurun_out/prof_single.log
