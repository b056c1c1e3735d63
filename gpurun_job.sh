# Scratch job for `gpurun -- 'bash gpurun_job.sh'`: the round's standard validation (tests, smoke, both bench arms, config-3 line).
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8) > gpurun_out/pytest_gpu_full.log; cat gpurun_out/pytest_gpu_full.log
(timeout 150 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/smoke.log; cat gpurun_out/smoke.log
(timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err); tail -c 1500 gpurun_out/bench.json
(timeout 240 python bench.py --workload hnsw --steps 20 --warmup 3 > gpurun_out/bench_hnsw.json 2> gpurun_out/bench_hnsw.err); tail -c 1200 gpurun_out/bench_hnsw.json
(timeout 150 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err); tail -c 400 gpurun_out/bench_ref.json
