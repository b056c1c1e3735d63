mkdir -p gpurun_out
for m in 0 2 0 2; do
  (COLTT_FAST_PFMODE=$m timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu 2> gpurun_out/b22.err | grep '^{' > gpurun_out/b22_$m.json); tail -2 gpurun_out/b22.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/b22_$m.json").read().strip().split("\n")[-1])
print("pfmode $m c2: value", round(j["value"]), "ms/step", round(j["ms_per_step"],4), "K2", round(j["kernel_ms"]["scan"],4), "frac", round(j["roofline"]["frac"],3))
PY
done
