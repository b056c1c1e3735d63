mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_f8e.py -x -q 2>&1 | tail -3)
(timeout 300 python bench.py --steps 500 --warmup 10 --no-cpu --no-extras > gpurun_out/r2_c2_fastloop.json 2> gpurun_out/r2_c2_fastloop.err)
(COLTT_FAST_SB=128 timeout 300 python bench.py --steps 500 --warmup 10 --no-cpu --no-extras > gpurun_out/r2_c2_sb128b.json 2> gpurun_out/r2_c2_sb128b.err)
(timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_c4_lock.json 2> gpurun_out/r2_c4_lock.err)
python - <<PY
import json
for nm in ("c2_fastloop","c2_sb128b","c4_lock"):
    try:
        j=json.load(open(f"gpurun_out/r2_{nm}.json")); print(nm, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "kernel ms", j["kernel_ms"], "frac", round(j["roofline"]["frac"],3), "fast", j["fast_path"], "clk", j["clocks"])
    except Exception as e: print(nm, "failed", e)
PY
tail -n 3 gpurun_out/r2_c2_fastloop.err gpurun_out/r2_c4_lock.err
(timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:gemm_filter_pair -s 4 -c 1 --csv --log-file gpurun_out/r2_c4_lock_ncu.csv python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu > /dev/null 2> gpurun_out/r2_c4_lock_ncu.err); grep -v "^==" gpurun_out/r2_c4_lock_ncu.csv | cut -d, -f13- | cut -c1-200
(timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:gemm_filter_pair -s 4 -c 1 --csv --log-file gpurun_out/r2_c2_fast_ncu.csv python bench.py --steps 3 --warmup 2 --no-cpu --no-extras > /dev/null 2> gpurun_out/r2_c2_fast_ncu.err); grep -v "^==" gpurun_out/r2_c2_fast_ncu.csv | cut -d, -f13- | cut -c1-200
