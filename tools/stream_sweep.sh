run() { echo "== $1"; env $1 python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 40 --pool 4 $2 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(round(d['value'],1), round(d['ms_per_step'],3), d['config']['cuda_streams'], d['single_stream']['ms_per_scan'])"; }
run "GG_STREAMS=4 GG_STAGGER=0"
run "GG_STREAMS=8 GG_STAGGER=0"
run "GG_STREAMS=8 GG_STAGGER=1"
run "GG_STREAMS=4 GG_STAGGER=1"
run "GG_STREAMS=2 GG_STAGGER=1"
run "GG_STREAMS=8 GG_STAGGER=1" "--streams 444"
run "GG_STREAMS=4 GG_STAGGER=0" "--streams 444"
