mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_comm.py "tests/test_gpu_flat.py::test_mutations_wait_for_an_asynchronous_search" -x -q 2>&1 | tail -6)
for flags in 0 2 8 10 12; do
(COLTT_B200_LIB=$PWD/coltt_b200/lib/libcoltt_b200_prof.so COLTT_DEBUG_FLAGS=$flags COLTT_DEBUG_PROF=1 timeout 150 python bench.py --steps 6 --warmup 4 --no-cpu --no-extras > gpurun_out/r2_probe2_c2_$flags.json 2> gpurun_out/r2_probe2_c2_$flags.err); echo "=== c2 flags=$flags"; grep "coltt prof" gpurun_out/r2_probe2_c2_$flags.err | grep "mma_\|epi_wait\|epi_total\|sweep" ; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2_probe2_c2_$flags.json")); print("c2 flags=$flags scan ms", round(j["kernel_ms"]["scan"],4), "clk", j["clocks"]["sm_mhz"])
except Exception as e: print("failed", e)
PY
done
