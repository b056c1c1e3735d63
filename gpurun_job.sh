mkdir -p gpurun_out
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --math fast > gpurun_out/bench_2gpu_full.log 2>&1)
grep -E "metric" gpurun_out/bench_2gpu_full.log | cut -c1-1400; tail -n 3 gpurun_out/bench_2gpu_full.log | cut -c1-300
(COLTT_BENCH_BREAKDOWN=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --math fast > gpurun_out/bench_2gpu_brk.log 2>&1)
grep -E "breakdown|Error|error" gpurun_out/bench_2gpu_brk.log | head -5 | cut -c1-300
