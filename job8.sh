mkdir -p gpurun_out
nvidia-smi -L | wc -l
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload c4 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_c4_8gpu.json 2> gpurun_out/r2_c4_8gpu.err); tail -c 2600 gpurun_out/r2_c4_8gpu.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_c4_8gpu.err | tail -n 5
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 500 --warmup 10 --no-extras > gpurun_out/r2_flat_8gpu.json 2> gpurun_out/r2_flat_8gpu.err); tail -c 1500 gpurun_out/r2_flat_8gpu.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_flat_8gpu.err | tail -n 5
