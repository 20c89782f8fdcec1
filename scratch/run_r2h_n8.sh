#!/bin/bash
# 8-GPU box: configs[3] strong scaling at N = 1, 2, 4, 8 (same box), then the default bench at N = 8
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 python bench.py --recording-hours 1 --steps 5 > gpurun_out/bench_r2h_rec_n1.json 2> gpurun_out/bench_r2h_rec_n1.err
echo "rec n1 exit $?"
port=29600
for n in 2 4 8; do
  port=$((port+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 5 --recording-hours 1 > gpurun_out/bench_r2h_rec_n$n.json 2> gpurun_out/bench_r2h_rec_n$n.err
  echo "rec n$n exit $?"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2h_n8.json 2> gpurun_out/bench_r2h_n8.err
echo "bench n8 exit $?"
tail -c 400 gpurun_out/bench_r2h_n8.err
nproc; free -g | head -2
