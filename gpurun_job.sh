mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fast.py -m gpu -q 2>&1 | tail -8) > gpurun_out/pytest_fast.log
(timeout 300 python bench.py --steps 40 --warmup 5 --math fast --no-cpu 2>&1 | tail -1 | cut -c1-1500) > gpurun_out/bench_fast.log
(timeout 600 python tools/probe_shapes.py 2>&1 | tail -40) > gpurun_out/probe.log
tail -n 4 gpurun_out/pytest_fast.log; cat gpurun_out/bench_fast.log; cat gpurun_out/probe.log
