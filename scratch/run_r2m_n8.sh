#!/bin/bash
# 8-GPU box: configs[3] strong scaling at N = 1, 2, 4, 8 (same box), then the default bench at N = 8
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 python bench.py --recording-hours 1 --steps 5 > gpurun_out/bench_r2m_rec_n1.json 2> gpurun_out/bench_r2m_rec_n1.err
echo "rec n1 exit $?"
port=29600
for n in 2 4 8; do
  port=$((port+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 5 --recording-hours 1 > gpurun_out/bench_r2m_rec_n$n.json 2> gpurun_out/bench_r2m_rec_n$n.err
  echo "rec n$n exit $?"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2m_n8.json 2> gpurun_out/bench_r2m_n8.err
echo "bench n8 exit $?"
tail -c 400 gpurun_out/bench_r2m_n8.err
nproc; free -g | head -2
python -m pytest tests/test_gpu_facade.py -k spread -x -q -m gpu 2>&1 | tail -3
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads([l for l in open(f'gpurun_out/bench_r2m_rec_n{n}.json') if l.startswith('{')][-1])
        r=d['recording']
        print(n, {k:round(r[k],3) if isinstance(r[k],float) else r[k] for k in ('ms_per_pass','decode_ms_this_rank','nccl_allgather_ms','exchange_ms_host','stitch_ms','parity_window')})
    except Exception as e: print(n,'ERR',e)
try:
    d=json.loads([l for l in open('gpurun_out/bench_r2m_n8.json') if l.startswith('{')][-1])
    print('n8 value',d['value']/1e9,'e2e',d['e2e']['value']/1e9,'rec',d['recording']['ms_per_pass'])
except Exception as e: print('n8 ERR',e)
PY
