mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_hnsw_build.py -m gpu -q -x 2>&1 | tail -30) > gpurun_out/pytest_hb.log
cat gpurun_out/pytest_hb.log
(DATA=latent N=100000 NQ=256,1024 timeout 300 python tools/probe_hnsw_1m.py > gpurun_out/hnsw_100k_v2.log 2>&1); cat gpurun_out/hnsw_100k_v2.log
(DATA=latent N=1000000 NQ=256,1024 timeout 600 python tools/probe_hnsw_1m.py > gpurun_out/hnsw_1m_v2.log 2>&1); cat gpurun_out/hnsw_1m_v2.log
