mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_hnsw_build.py -x -q 2>&1 | tail -3)
for ring in "512,2" "384,2" "256,2" "256,3" "256,4" "512,3" "384,3"; do
(COLTT_HNSW_RING=$ring timeout 400 python bench.py --workload hnsw --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_hnsw_ring2_$ring.json 2> gpurun_out/r2_hnsw_ring.err); python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2_hnsw_ring2_$ring.json")); print("hnsw ring='$ring' value", round(j["value"]), "kernel ms", round(j["roofline"]["kernel_ms"],3), "frac", round(j["roofline"]["frac"],3))
except Exception as e: print("ring $ring failed", e)
PY
done
(time timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err); tail -c 4000 gpurun_out/r2_bench_default.json; tail -n 5 gpurun_out/r2_bench_default.err
(time timeout 300 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err); tail -c 600 gpurun_out/r2_bench_ref.json
