#!/bin/bash
# Multi-GPU evidence in one gpurun call (charged N x box time, so everything that needs N GPUs runs here, once):
#   tools/run_multigpu.sh N [tag]   ->  gpurun_out/<tag>_*   (topology, D2H ceiling at 1..N ranks, the byte-equality test,
#                                        the C2 line at N GPUs and the C5 corpus line at N GPUs)
N=${1:-2}; TAG=${2:-r02_mg}; O=gpurun_out
mkdir -p $O
{ nproc; lscpu | grep -i "numa\|model name\|socket"; nvidia-smi topo -m; free -g | head -2; } > $O/${TAG}_topo.txt 2>&1
: > $O/${TAG}_d2h_ceiling.txt
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      tools/d2h_ceiling.py 2>> $O/${TAG}_d2h.err | grep '^{' >> $O/${TAG}_d2h_ceiling.txt
done
timeout 600 python -m pytest tests/test_corpus_gpu.py -m gpu -q -x -k "multi_gpu" > $O/${TAG}_pytest_multigpu.log 2>&1
tail -3 $O/${TAG}_pytest_multigpu.log
for n in $N; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 5 --warmup 3 > $O/${TAG}_bench_c2_${n}gpu.json 2> $O/${TAG}_bench_c2_${n}gpu.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      bench.py --gpus $n --steps 5 --warmup 3 --no-numa-bind > $O/${TAG}_bench_c2_${n}gpu_unbound.json 2> $O/${TAG}_bench_c2_${n}gpu_unbound.err
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) \
      bench.py --gpus $n --workload c5 --files ${C5_FILES:-8} --steps 2 --warmup 3 > $O/${TAG}_bench_c5_${n}gpu.json 2> $O/${TAG}_bench_c5_${n}gpu.err
done
cat $O/${TAG}_d2h_ceiling.txt
