# usage (GPU box): bash tools/stream_sweep.sh "<env>" "<bench args>" ...   (pairs)
run() { echo "== $1 | $2"; env $1 python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 30 --pool 2 $2 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(round(d['value'],1), round(d['ms_per_step'],3), d['config']['cuda_streams'], round(d['single_stream']['ms_per_scan'],4), d['roofline']['kernel_avg_launch_us_live'].get('k_spiral'))"; }
while [ $# -gt 0 ]; do run "$1" "$2"; shift 2; done
