mkdir -p gpurun_out
(timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err); tail -c 2500 gpurun_out/bench_final.json; tail -n 3 gpurun_out/bench_final.err
(timeout 240 python -m pytest tests/test_gpu_fast.py tests/test_gpu_hnsw.py tests/test_gpu_dist.py tests/test_gpu_flat.py -m gpu -q -x -k "not config2_shape_exact" 2>&1 | tail -8) > gpurun_out/pytest_final.log; cat gpurun_out/pytest_final.log
(timeout 150 python __graft_entry__.py smoke 2>&1 | tail -4) > gpurun_out/smoke_final.log; cat gpurun_out/smoke_final.log
(timeout 240 python bench.py --workload hnsw --steps 20 --warmup 3 > gpurun_out/bench_hnsw.json 2> gpurun_out/bench_hnsw.err); tail -c 2500 gpurun_out/bench_hnsw.json; tail -n 3 gpurun_out/bench_hnsw.err
(timeout 150 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err); tail -c 600 gpurun_out/bench_ref_final.json
