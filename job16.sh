mkdir -p gpurun_out
(timeout 300 python -m pytest "tests/test_gpu_f8e.py::test_e4m3_edge_cases" -x -q 2>&1 | tail -12)
