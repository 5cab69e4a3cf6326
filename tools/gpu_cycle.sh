#!/bin/bash
# Builder-side helper (runs in the build container): one standard GPU cycle under gpurun with retries while the pod is busy.
#   tools/gpu_cycle.sh <tag> [extra shell appended on the GPU box]
# -> gpurun_out/<tag>_{tests.log,bench.json,bench.err,launches.csv}
tag=$1; extra=$2
cmd="python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_tests.log; tail -5 gpurun_out/${tag}_tests.log; python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; bash tools/profile_ncu.sh ${tag} ${FULL} > /dev/null 2>&1; tail -3 gpurun_out/${tag}_bench.err; ${extra}"
for attempt in 1 2 3 4 5 6 7 8 9 10; do
    /usr/local/graft/bin/gpurun --timeout 1500 -- "$cmd" > gpurun_out/${tag}.out 2>&1
    if grep -q "status=transient" gpurun_out/${tag}.out; then sleep 120; continue; fi
    break
done
tail -6 gpurun_out/${tag}.out
python tools/launch_table.py gpurun_out/${tag}_launches.csv
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1) if d.get('e2e') else None, 'path frac', round(d['roofline_path']['frac'],4))
print({k: v['avg_launch_us'] for k, v in d['roofline'].get('per_kernel', {}).items()})
print(d['single_stream'])
PY
