# Scratch job for `gpurun -- 'bash gpurun_job.sh'`: the round's standard validation (tests, smoke, both bench arms).
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15) > gpurun_out/pytest_gpu_full.log; cat gpurun_out/pytest_gpu_full.log
(timeout 200 python __graft_entry__.py smoke 2>&1 | tail -6) > gpurun_out/smoke.log; cat gpurun_out/smoke.log
(timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err); tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
(timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err); tail -c 400 gpurun_out/bench_ref.json
