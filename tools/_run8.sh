mkdir -p gpurun_out/r2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2/bench_8gpu.json 2> gpurun_out/r2/bench_8gpu.err
tail -c 600 gpurun_out/r2/bench_8gpu.json; tail -3 gpurun_out/r2/bench_8gpu.err
nvidia-smi topo -m > gpurun_out/r2/topo_8gpu.txt 2>&1; lscpu | grep -E "NUMA|Model name|Socket|^CPU\(s\)" >> gpurun_out/r2/topo_8gpu.txt
