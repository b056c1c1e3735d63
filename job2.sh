mkdir -p gpurun_out
for w in 1 2 4; do
(COLTT_TMA_WORD=$w timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu --no-extras > gpurun_out/r2_c2_w$w.json 2> gpurun_out/r2_c2_w$w.err); python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2_c2_w$w.json")); print("c2 word $w: value", round(j["value"]), "scan ms", round(j["kernel_ms"]["scan"],4), "clk", j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
except Exception as e: print("c2 word $w failed", e)
PY
tail -2 gpurun_out/r2_c2_w$w.err
done
for w in 1 4; do
(COLTT_TMA_WORD=$w timeout 300 python bench.py --workload c4 --rows 2000000 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2_c4_w$w.json 2> gpurun_out/r2_c4_w$w.err); python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2_c4_w$w.json")); print("c4(2M) word $w: value", round(j["value"]), "kernel ms", j["kernel_ms"], "frac", round(j["roofline"]["frac"],3))
except Exception as e: print("c4 word $w failed", e)
PY
tail -2 gpurun_out/r2_c4_w$w.err
done
(timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r2_c4_full.json 2> gpurun_out/r2_c4_full.err); tail -c 2500 gpurun_out/r2_c4_full.json; tail -3 gpurun_out/r2_c4_full.err
(timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r2_pytest_full.log; cat gpurun_out/r2_pytest_full.log
