# Scratch job for `gpurun -- 'bash gpurun_job.sh'`: the round's standard validation (tests, smoke, both bench arms as the driver runs them).
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15) > gpurun_out/pytest_gpu_full.log; cat gpurun_out/pytest_gpu_full.log
(timeout 200 python __graft_entry__.py smoke 2>&1 | tail -6) > gpurun_out/smoke.log; cat gpurun_out/smoke.log
(timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench_driver.err | grep '^{' > gpurun_out/bench_driver.json); tail -c 1500 gpurun_out/bench_driver.json; tail -5 gpurun_out/bench_driver.err
(timeout 600 python bench.py 2> gpurun_out/bench.err | grep '^{' > gpurun_out/bench.json); tail -c 600 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
(timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench_ref.err | grep '^{' > gpurun_out/bench_ref.json); tail -c 400 gpurun_out/bench_ref.json
