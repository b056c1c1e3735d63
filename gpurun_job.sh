mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_hnsw_build.py -m gpu -q -x 2>&1 | tail -30) > gpurun_out/pytest_hb.log
cat gpurun_out/pytest_hb.log
(COLTT_HNSW_DEBUG=1 DATA=latent N=1000000 NQ=256,1024,4096 timeout 600 python tools/probe_hnsw_1m.py > gpurun_out/hnsw_1m_v7.log 2>gpurun_out/hnsw_1m_v7.err); cat gpurun_out/hnsw_1m_v7.log | cut -c1-120,440-; sort gpurun_out/hnsw_1m_v7.err | uniq -c | head
