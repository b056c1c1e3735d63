mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_flat.py tests/test_gpu_comm.py -x -q -m gpu 2>&1 | tail -5)
(timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras 2> gpurun_out/b19a.err | grep '^{' > gpurun_out/b19a.json); tail -3 gpurun_out/b19a.err
(timeout 300 python bench.py --no-extras 2> gpurun_out/b19b.err | grep '^{' > gpurun_out/b19b.json); tail -3 gpurun_out/b19b.err
python - <<PY
import json
for nm in ("a","b"):
    try:
        j=json.loads(open(f"gpurun_out/b19{nm}.json").read().strip().split("\n")[-1])
        print(nm, "value", round(j["value"]), "ms/step", round(j["ms_per_step"],4), "e2e", j["e2e"], j["clocks"])
    except Exception as e: print(nm, "failed", e)
PY
