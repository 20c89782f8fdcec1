#!/bin/bash
cd "$GRAFT_REPO_ROOT"
LEAN="--no-e2e --no-cpu --no-online --no-deskew --no-single-pass --no-parity --no-facade --no-hdl32 --recording-leg-hours 0 --online-udp-seconds 0"
for v in "1 -" "0 -" "1 1" "0 1" "1 -" "0 -" "1 1" "0 1"; do
  set -- $v
  if [ "$2" = "-" ]; then unset VELOSLAM_RESET_KERNEL; else export VELOSLAM_RESET_KERNEL=$2; fi
  VELOSLAM_DECODE_CHAIN=$1 python bench.py --steps 20 --warmup 3 $LEAN | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']
print('chain=$1 reset_kernel=$2', 'ms_per_step', round(d['ms_per_step'],4), 'k_decode_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), 'whole', round(r['frac_whole_step'],4), 'G/s', round(d['value']/1e9,2))"
done
