mkdir -p gpurun_out
cd _r1 && (timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu > ../gpurun_out/r2_ab_old.json 2> ../gpurun_out/r2_ab_old.err); cd ..
python - <<PY
import json
for nm in ("old",):
    try:
        j=json.load(open(f"gpurun_out/r2_ab_{nm}.json")); print(nm, "value", round(j["value"]), "kernel ms", j["kernel_ms"], "clk", j["clocks"])
    except Exception as e: print(nm, "failed", e)
PY
tail -3 gpurun_out/r2_ab_old.err
(timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu --no-extras > gpurun_out/r2_ab_new.json 2> gpurun_out/r2_ab_new.err)
python - <<PY
import json
for nm in ("new",):
    try:
        j=json.load(open(f"gpurun_out/r2_ab_{nm}.json")); print(nm, "value", round(j["value"]), "kernel ms", j["kernel_ms"], "clk", j["clocks"])
    except Exception as e: print(nm, "failed", e)
PY
tail -3 gpurun_out/r2_ab_new.err
cd _r1 && (timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_filter_pair -s 3 -c 1 -o ../gpurun_out/r2_k2_old python bench.py --steps 3 --warmup 2 --no-cpu > /dev/null 2> ../gpurun_out/r2_ncu_old.err); cd ..
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_filter_pair -s 3 -c 1 -o gpurun_out/r2_k2_new python bench.py --steps 3 --warmup 2 --no-cpu --no-extras > /dev/null 2> gpurun_out/r2_ncu_new.err)
ls -la gpurun_out/*.ncu-rep
