mkdir -p gpurun_out
nvidia-smi -L | wc -l
(timeout 300 python -m pytest "tests/test_gpu_f8e.py::test_e4m3_edge_cases" -x -q 2>&1 | tail -8)
(timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_comm.py -x -q 2>&1 | tail -15) > gpurun_out/r2_pytest_dist_p2p.log; cat gpurun_out/r2_pytest_dist_p2p.log
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 300 --warmup 10 --no-extras > gpurun_out/r2_flat_2gpu_p2p.json 2> gpurun_out/r2_flat_2gpu_p2p.err); python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2_flat_2gpu_p2p.json")); print("p2p 2gpu: global_qps", round(j["global_qps"]), "ms/step", round(j["ms_per_step"],4), "e2e global", round(j["e2e"]["global_qps"]), "merge_check", j.get("merge_check"), "launches", j["gpu_launches"])
except Exception as e: print("failed", e)
PY
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_flat_2gpu_p2p.err | tail -n 5
(COLTT_P2P=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 300 --warmup 10 --no-extras > gpurun_out/r2_flat_2gpu_nccl.json 2> gpurun_out/r2_flat_2gpu_nccl.err); python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2_flat_2gpu_nccl.json")); print("nccl 2gpu: global_qps", round(j["global_qps"]), "ms/step", round(j["ms_per_step"],4), "e2e global", round(j["e2e"]["global_qps"]), "merge_check", j.get("merge_check"), "launches", j["gpu_launches"])
except Exception as e: print("failed", e)
PY
