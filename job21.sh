mkdir -p gpurun_out
for m in 0 1 0 1; do
  (COLTT_FAST_PFMODE=$m timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu 2> gpurun_out/b21.err | grep '^{' > gpurun_out/b21_$m.json); tail -2 gpurun_out/b21.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/b21_$m.json").read().strip().split("\n")[-1])
print("pfmode $m c2: value", round(j["value"]), "ms/step", round(j["ms_per_step"],4), "K2", round(j["kernel_ms"]["scan"],4), "frac", round(j["roofline"]["frac"],3))
PY
done
for m in 0 1; do
  (COLTT_FAST_PFMODE=$m timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu 2> gpurun_out/b21.err | grep '^{' > gpurun_out/b21c4_$m.json); tail -2 gpurun_out/b21.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/b21c4_$m.json").read().strip().split("\n")[-1])
print("pfmode $m c4: value", round(j["value"]), "ms/step", round(j["ms_per_step"],4), "K2", round(j["kernel_ms"]["scan"],4), "frac", round(j["roofline"]["frac"],3), j["clocks"]["sm_mhz"])
PY
done
COLTT_FAST_PFMODE=1 timeout 300 python -m pytest tests/test_gpu_fast.py tests/test_gpu_f8e.py -x -q -m gpu 2>&1 | tail -3
