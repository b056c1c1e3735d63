mkdir -p gpurun_out
(timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu --no-extras > gpurun_out/r2_c2_fix.json 2> gpurun_out/r2_c2_fix.err)
(timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu --no-extras --quant none > gpurun_out/r2_c2_f32.json 2> gpurun_out/r2_c2_f32.err)
(timeout 300 python bench.py --workload c4 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_c4_fix.json 2> gpurun_out/r2_c4_fix.err)
python - <<PY
import json
for nm in ("c2_fix","c2_f32","c4_fix"):
    try:
        j=json.load(open(f"gpurun_out/r2_{nm}.json")); print(nm, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "kernel ms", j["kernel_ms"], "frac", round(j["roofline"]["frac"],3), "fast", j.get("fast_path"), "clk", j["clocks"])
    except Exception as e: print(nm, "failed", e)
PY
tail -3 gpurun_out/r2_c2_fix.err gpurun_out/r2_c2_f32.err gpurun_out/r2_c4_fix.err
(timeout 900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_f8e.py -x -q -s 2>&1 | tail -12)
