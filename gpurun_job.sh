mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_flat.py -m gpu -q -x -s 2>&1 | tail -30) > gpurun_out/pytest_multi.log
cat gpurun_out/pytest_multi.log
