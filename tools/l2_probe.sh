# usage (GPU box): per-kernel time per scan of a serialised pass at small batch sizes (does the inter-kernel data of a
# sub-batch that fits the 126 MB L2 make the bulk kernels faster than the 444-scan launches that stream through HBM?)
for s in 6 12 24 48 444; do
  GG_STREAMS=1 python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --pool 2 --streams $s 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); pk=d['roofline']['per_kernel']; print($s, 'scans/launch: us per scan', {k:round(v['avg_launch_us']/$s,2) for k,v in pk.items()}, 'sum', round(sum(v['avg_launch_us'] for v in pk.values())/$s,2))"
done
