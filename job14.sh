mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_f8e.py -x -q 2>&1 | tail -3)
(timeout 300 python bench.py --steps 500 --warmup 10 --no-cpu --no-extras > gpurun_out/r2_c2_subsweep.json 2> gpurun_out/r2_c2_subsweep.err)
for w in 0 8 16 32 64; do
(COLTT_LOCK_WINDOW=$w timeout 300 python bench.py --workload c4 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_c4_w$w.json 2> gpurun_out/r2_c4_w$w.err)
done
python - <<PY
import json
for nm in ("c2_subsweep","c4_w0","c4_w8","c4_w16","c4_w32","c4_w64"):
    try:
        j=json.load(open(f"gpurun_out/r2_{nm}.json")); print(nm, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "kernel ms", j["kernel_ms"], "frac", round(j["roofline"]["frac"],3), "fast", j["fast_path"], "clk", j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
    except Exception as e: print(nm, "failed", e)
PY
for w in 16 32; do
(COLTT_LOCK_WINDOW=$w timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:gemm_filter_pair -s 4 -c 1 --csv --log-file gpurun_out/r2_c4_w${w}_ncu.csv python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu > /dev/null 2> gpurun_out/r2_c4_w_ncu.err); echo "window $w"; grep -v "^==" gpurun_out/r2_c4_w${w}_ncu.csv | cut -d, -f13- | cut -c1-120
done
