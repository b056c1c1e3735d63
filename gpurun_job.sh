mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fast.py -m gpu -q -x 2>&1 | tail -4) > gpurun_out/pytest_fast.log
(timeout 300 python tools/probe_step.py > gpurun_out/step.log 2>&1)
(COLTT_DEBUG_NOTAIL=1 timeout 300 python tools/probe_step.py > gpurun_out/step_notail.log 2>&1)
cat gpurun_out/pytest_fast.log gpurun_out/step.log gpurun_out/step_notail.log
