mkdir -p gpurun_out
nvidia-smi -L | wc -l
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 500 --warmup 10 --no-extras 2> gpurun_out/r2_flat_8gpu_p2p.err | grep '^{' > gpurun_out/r2_flat_8gpu_p2p.json)
(COLTT_P2P=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 500 --warmup 10 --no-extras 2> gpurun_out/r2_flat_8gpu_nccl.err | grep '^{' > gpurun_out/r2_flat_8gpu_nccl.json)
python - <<PY
import json
for nm in ("p2p","nccl"):
    try:
        j=json.loads(open(f"gpurun_out/r2_flat_8gpu_{nm}.json").read().strip().split("\n")[-1])
        print(nm, "global_qps", round(j["global_qps"]), "ms/step", round(j["ms_per_step"],4), "e2e global", round(j["e2e"]["global_qps"]), "merge_check", j.get("merge_check"), "launches", j["gpu_launches"], j["clocks"]["sm_mhz"])
    except Exception as e: print(nm, "failed", e)
PY
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_flat_8gpu_p2p.err | tail -n 5
