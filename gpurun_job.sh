mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_hnsw_build.py tests/test_gpu_hnsw.py -m gpu -q -x 2>&1 | tail -40) > gpurun_out/pytest_hb.log
cat gpurun_out/pytest_hb.log
