mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_pq.py -x -q -s 2>&1 | tail -15) > gpurun_out/r2_pytest_pq.log; cat gpurun_out/r2_pytest_pq.log
(N=1000000 D=768 timeout 300 python tools/probe_shapes.py 2>&1 | grep "nq=   1\|nq=   8") > gpurun_out/r2_shapes.log; cat gpurun_out/r2_shapes.log
(timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r2_pytest_full2.log; cat gpurun_out/r2_pytest_full2.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6) > gpurun_out/r2_smoke.log; cat gpurun_out/r2_smoke.log
