mkdir -p gpurun_out
(COLTT_DEBUG_PROF=1 timeout 300 python bench.py --steps 8 --warmup 3 --math fast --no-cpu 2>&1 | grep -E "coltt prof|ms_per_step" | cut -c1-300) > gpurun_out/prof_roles.log
(COLTT_DEBUG_PROF=1 COLTT_DEBUG_FLAGS=1 timeout 300 python bench.py --steps 8 --warmup 3 --math fast --no-cpu 2>&1 | grep -E "coltt prof|ms_per_step" | cut -c1-300) > gpurun_out/prof_roles_noepi.log
cat gpurun_out/prof_roles.log; echo ---; cat gpurun_out/prof_roles_noepi.log
