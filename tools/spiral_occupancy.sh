# usage (GPU box): how the batched spiral kernel scales with co-resident CTAs per SM (1, 2, 3 CTAs on each of the 148 SMs)
for s in 148 296 444; do
  python bench.py --no-cpu-baseline --no-e2e --steps 10 --pool 2 --streams $s 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); pk=d['roofline']['per_kernel']; print($s, 'streams: spiral', pk['k_spiral']['avg_launch_us'], 'us serialised;', {k:v['avg_launch_us'] for k,v in pk.items()})"
done
