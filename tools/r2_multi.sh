#!/bin/bash
# tools/r2_multi.sh N : the bench under torchrun on N GPUs of one box (gpurun --gpus N -- 'bash tools/r2_multi.sh N')
N=${1:-2}; mkdir -p gpurun_out/r2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/r2/bench_${N}gpu.json 2> gpurun_out/r2/bench_${N}gpu.err
tail -c 600 gpurun_out/r2/bench_${N}gpu.json; tail -3 gpurun_out/r2/bench_${N}gpu.err
nvidia-smi topo -m > gpurun_out/r2/topo_${N}gpu.txt 2>&1; lscpu | grep -E "NUMA|Model name|Socket|^CPU\(s\)" >> gpurun_out/r2/topo_${N}gpu.txt
