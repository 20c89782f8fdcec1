#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2v_n2.json 2> gpurun_out/bench_r2v_n2.err
echo "bench n2 exit $?"; tail -c 300 gpurun_out/bench_r2v_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r2v_n2_reference.json 2> gpurun_out/bench_r2v_n2_reference.err
echo "reference n2 exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2v_n2.json') if l.startswith('{')][-1])
print('n2 value',d['value']/1e9,'e2e',d['e2e']['value']/1e9,'rec',d['recording']['ms_per_pass'], d['recording']['parity_window'], 'udp', d['online_udp'].get('rotation_p99_ms_worst_stream'), d['online_udp'].get('dropped_packets'), 'facade', d['e2e_facade']['frames_xyzi_meta'].get('value'))
r=json.loads([l for l in open('gpurun_out/bench_r2v_n2_reference.json') if l.startswith('{')][-1]); print('ref', r.get('value'), r.get('cpu_baseline',{}).get('cores'))
PY
