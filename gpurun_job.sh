mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fast.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/pytest_fast.log
run() { echo "== $*"; env "$@" timeout 300 python tools/probe_k2.py 2>&1 | grep -E "K2 N=|rror" ; }
(
run A=1
run COLTT_PF_INNER=64
run COLTT_PF_INNER=256
run COLTT_FAST_NS=3
run COLTT_DEBUG_FLAGS=1
run COLTT_DEBUG_FLAGS=2
run NQ=128
run NQ=128 COLTT_PF_INNER=64
) > gpurun_out/k2_knobs8.log 2>&1
tail -3 gpurun_out/pytest_fast.log; cat gpurun_out/k2_knobs8.log
