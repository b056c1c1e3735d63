mkdir -p gpurun_out
(timeout 60 python -m pytest tests/test_gpu_flat.py -m gpu -q -x -k "config4" 2>&1 | tail -5) > gpurun_out/pytest_c4.log; cat gpurun_out/pytest_c4.log
