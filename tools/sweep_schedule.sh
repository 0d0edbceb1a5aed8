#!/bin/bash
# Schedule sweep on one B200 (under gpurun): chunk length, ramp, network SM share; device-resident default bench shape.
O=gpurun_out
for cfg in "8 1 64" "8 0 64" "10 0 64" "12 0 64" "16 0 64" "10 0 72" "10 1 64" "20 0 64"; do
  set -- $cfg
  PNB_CHUNK=$1 PNB_RAMP=$2 PNB_NET_SMS=$3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-int16-run --no-e2e > $O/sw_$1_$2_$3.json 2>/dev/null
  python - "$O/sw_$1_$2_$3.json" "$cfg" <<'P'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], round(d['value']/1e6,3), d['clocks']['sm_mhz'])
P
done
