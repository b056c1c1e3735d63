mkdir -p gpurun_out
for flags in 0 4 8 12; do
echo "=== c2 flags=$flags"
(COLTT_B200_LIB=$PWD/coltt_b200/lib/libcoltt_b200_prof.so COLTT_DEBUG_FLAGS=$flags COLTT_DEBUG_PROF=1 timeout 300 python bench.py --steps 6 --warmup 4 --no-cpu --no-extras > gpurun_out/r2_probe_c2_$flags.json 2> gpurun_out/r2_probe_c2_$flags.err); grep "coltt prof" gpurun_out/r2_probe_c2_$flags.err | head -14; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2_probe_c2_$flags.json")); print("c2 flags=$flags scan ms", round(j["kernel_ms"]["scan"],4), "clk", j["clocks"]["sm_mhz"])
except Exception as e: print("failed", e)
PY
done
for flags in 0 4 8; do
echo "=== c4 flags=$flags"
(COLTT_B200_LIB=$PWD/coltt_b200/lib/libcoltt_b200_prof.so COLTT_DEBUG_FLAGS=$flags COLTT_DEBUG_PROF=1 timeout 300 python bench.py --workload c4 --rows 2000000 --steps 4 --warmup 4 --no-cpu > gpurun_out/r2_probe_c4_$flags.json 2> gpurun_out/r2_probe_c4_$flags.err); grep "coltt prof" gpurun_out/r2_probe_c4_$flags.err | head -14; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2_probe_c4_$flags.json")); print("c4 flags=$flags scan ms", round(j["kernel_ms"]["scan"],4), "clk", j["clocks"]["sm_mhz"])
except Exception as e: print("failed", e)
PY
done
