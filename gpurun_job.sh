mkdir -p gpurun_out
(DATA=latent N=100000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search -s 2 -c 1 -o gpurun_out/k4_full -f python tools/probe_hnsw_1m.py > gpurun_out/ncu_k4.log 2>&1)
(DATA=latent N=100000 NQ=1024 timeout 600 python tools/probe_hnsw_1m.py > gpurun_out/hnsw_100k_nq1024.log 2>&1)
tail -n 3 gpurun_out/ncu_k4.log; cat gpurun_out/hnsw_100k_nq1024.log
