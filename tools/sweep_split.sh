O=gpurun_out
for cfg in "0 64" "1 62" "1 60" "1 58"; do
  set -- $cfg
  PNB_SPLIT_FINE=$1 PNB_NET_SMS=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-int16-run --no-e2e > $O/sw2_$2.json 2>/dev/null
  python - "$O/sw2_$2.json" "$cfg" <<'P'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], round(d['value']/1e6,3), d['clocks']['sm_mhz'], d['schedule']['net_sms'], d['schedule']['dsp_sms'])
P
done
