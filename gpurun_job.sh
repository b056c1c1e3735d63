mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_fast.py -m gpu -q 2>&1 | tail -8) > gpurun_out/pytest_flat.log
(timeout 600 python tools/probe_shapes.py 2>&1 | tail -16) > gpurun_out/probe.log
(timeout 600 python tools/probe_hnsw.py 2>&1 | tail -3) > gpurun_out/probe_hnsw.log
tail -n 4 gpurun_out/pytest_flat.log; cat gpurun_out/probe.log; cat gpurun_out/probe_hnsw.log
