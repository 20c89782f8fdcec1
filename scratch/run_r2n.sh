#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_table_rows.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --recording-hours 1 --steps 5 --warmup 3 > gpurun_out/r2n_rec_n1.json 2> gpurun_out/r2n_rec_n1.err
echo "rec exit $?"; tail -c 400 gpurun_out/r2n_rec_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2n_rec_n1.json') if l.startswith('{')][-1])
r=d['recording']
print({k:r[k] for k in ('ms_per_pass','wall_ms_per_pass','decode_ms_this_rank','vs_wait_ms_host','exchange_ms_host','stitch_ms','global_frames','parity_window')})
PY
