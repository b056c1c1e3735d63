mkdir -p gpurun_out
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_filter_pair -s 4 -c 1 -o gpurun_out/r2_k2_c4_full python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu > /dev/null 2> gpurun_out/r2_ncu_c4.err); tail -n 2 gpurun_out/r2_ncu_c4.err
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_filter_pair -s 4 -c 1 -o gpurun_out/r2_k2_c2_full python bench.py --steps 3 --warmup 2 --no-cpu --no-extras > /dev/null 2> gpurun_out/r2_ncu_c2.err); tail -n 2 gpurun_out/r2_ncu_c2.err
(COLTT_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-extras > /dev/null 2> gpurun_out/r2_ncu_l2.err); tail -n 2 gpurun_out/r2_ncu_l2.err
(COLTT_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c4.csv python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > /dev/null 2> gpurun_out/r2_ncu_l4.err); tail -n 2 gpurun_out/r2_ncu_l4.err
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:rerank_kernel -s 4 -c 1 -o gpurun_out/r2_rerank_c4_full python bench.py --workload c4 --rows 2000000 --steps 3 --warmup 2 --no-cpu > /dev/null 2> gpurun_out/r2_ncu_rr.err); tail -n 2 gpurun_out/r2_ncu_rr.err
for ring in "" "768,2" "512,2" "1024,2"; do
echo "== RING=$ring"
(COLTT_HNSW_RING=$ring timeout 300 python -m pytest tests/test_gpu_hnsw.py -x -q 2>&1 | tail -2)
(COLTT_HNSW_RING=$ring timeout 400 python bench.py --workload hnsw --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_hnsw_ring_$ring.json 2> gpurun_out/r2_hnsw_ring.err); python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2_hnsw_ring_$ring.json")); print("hnsw ring='$ring' value", round(j["value"]), "kernel ms", round(j["roofline"]["kernel_ms"],3), "frac", round(j["roofline"]["frac"],3))
except Exception as e: print("failed", e)
PY
done
ls -la gpurun_out/*.ncu-rep
