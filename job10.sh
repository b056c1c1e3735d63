mkdir -p gpurun_out
nvidia-smi -L | wc -l
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --workload c5 --steps 10 --warmup 3 > gpurun_out/r2_c5_4gpu.json 2> gpurun_out/r2_c5_4gpu.err); tail -c 2600 gpurun_out/r2_c5_4gpu.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_c5_4gpu.err | tail -n 8
