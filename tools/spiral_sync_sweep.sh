# usage (GPU box): barrier vs point-to-point synchronisation of k_spiral_skew, batch and single stream
run() { python bench.py --no-cpu-baseline --no-e2e --steps 20 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$1', 'value', round(d['value'],1), 'spiral us', d['roofline']['per_kernel']['k_spiral']['avg_launch_us'], 'single ms/scan', round(d['single_stream']['ms_per_scan'],4))"; }
GG_SPIRAL_ASYNC=0 run barrier
GG_SPIRAL_SLEEP=0 run spin
GG_SPIRAL_SLEEP=20 run sleep20
GG_SPIRAL_SLEEP=200 run sleep200
