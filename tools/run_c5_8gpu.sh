#!/bin/bash
# BASELINE configs[4] at its shape: 64 AVI files (8 per GPU), 3840x2160, GOP-sharded over N GPUs of one box.
N=${1:-8}; TAG=${2:-r02_c5}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29810 \
    bench.py --gpus $N --workload c5 --files 8 --steps 2 --warmup 3 > $O/${TAG}_bench_c5_${N}gpu.json 2> $O/${TAG}_bench_c5_${N}gpu.err
grep "^{" $O/${TAG}_bench_c5_${N}gpu.json | cut -c1-600
tail -3 $O/${TAG}_bench_c5_${N}gpu.err
