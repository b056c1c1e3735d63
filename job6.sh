mkdir -p gpurun_out
(timeout 300 python bench.py --steps 500 --warmup 10 --no-cpu --no-extras > gpurun_out/r2_c2_graph.json 2> gpurun_out/r2_c2_graph.err)
(COLTT_GRAPHS=0 timeout 300 python bench.py --steps 500 --warmup 10 --no-cpu --no-extras > gpurun_out/r2_c2_nograph.json 2> gpurun_out/r2_c2_nograph.err)
python - <<PY
import json
for nm in ("c2_graph","c2_nograph"):
    try:
        j=json.load(open(f"gpurun_out/r2_{nm}.json")); print(nm, "value", round(j["value"]), "ms/step", round(j["ms_per_step"],4), "e2e", round(j["e2e"]["value"]), "kernel ms", j["kernel_ms"], "launches", j["gpu_launches"], "clk", j["clocks"])
    except Exception as e: print(nm, "failed", e)
PY
tail -n 3 gpurun_out/r2_c2_graph.err
(COLTT_SCAN_CTAS=2 N=1000000 D=768 timeout 300 python tools/probe_shapes.py 2>&1 | grep "nq=   1\|nq=   8") > gpurun_out/r2_shapes_ctas2.log; cat gpurun_out/r2_shapes_ctas2.log
(COLTT_SCAN_CTAS=2 timeout 600 python -m pytest tests/test_gpu_flat.py -x -q 2>&1 | tail -3)
(timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_fast.py tests/test_gpu_f8e.py tests/test_gpu_multi.py -x -q 2>&1 | tail -3)
