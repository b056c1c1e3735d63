mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_f8e.py tests/test_gpu_fast.py -x -q -s 2>&1 | tail -40) > gpurun_out/r2_pytest1.log; cat gpurun_out/r2_pytest1.log
for sb in 64 128; do
(COLTT_FAST_SB=$sb timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu --no-extras > gpurun_out/r2_c2_sb$sb.json 2> gpurun_out/r2_c2_sb$sb.err); tail -c 1800 gpurun_out/r2_c2_sb$sb.json; tail -3 gpurun_out/r2_c2_sb$sb.err
done
for sb in 64 128; do
(COLTT_FAST_SB=$sb timeout 300 python bench.py --workload c4 --rows 2000000 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2_c4_sb$sb.json 2> gpurun_out/r2_c4_sb$sb.err); tail -c 1800 gpurun_out/r2_c4_sb$sb.json; tail -3 gpurun_out/r2_c4_sb$sb.err
done
