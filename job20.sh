mkdir -p gpurun_out
(timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/b20a.err | grep '^{' > gpurun_out/b20a.json); tail -3 gpurun_out/b20a.err
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/b20b.err | grep '^{' > gpurun_out/b20b.json); grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/b20b.err | tail -3
(timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -4)
python - <<PY
import json
for nm in ("a","b"):
    try:
        j=json.loads(open(f"gpurun_out/b20{nm}.json").read().strip().split("\n")[-1])
        print(nm, "value", round(j["value"]), "ms/step", round(j["ms_per_step"],4), "e2e", j["e2e"], j["clocks"], j.get("merge_check"), j["config"].get("exchange"))
        if "c3" in j: print("  c3", j["c3"]["value"], j["c3"]["clocks"], "c4", j["c4"]["value"], j["c4"]["clocks"])
    except Exception as e: print(nm, "failed", e)
PY
