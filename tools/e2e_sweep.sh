# usage (on the GPU box): bash tools/e2e_sweep.sh  -- knob sweep, results under gpurun_out/
run() { # label, bench-args, env...
  label=$1; shift; bargs=$1; shift
  env "$@" timeout 200 python bench.py --steps 16 --no-cpu-baseline --no-e2e $bargs > gpurun_out/sw_$label.json 2>gpurun_out/sw_$label.err
  python -c "
import json;d=json.load(open('gpurun_out/sw_$label.json'));r=d['roofline']['kernel_avg_launch_us'];print('$label', 'value', round(d['value']), round(d['ms_per_step'],3), 'spiral', r['k_spiral'], 'detect', r['k_detect'], 'raster', r['k_rasterize'], 'cell_stats', r['k_cell_stats'], 'sc_lo', r['k_sort_scatter(lo)'], 'label', r['k_label'], 'single', round(d['single_stream']['ms_per_scan'],3))" || tail -5 gpurun_out/sw_$label.err
}
run s1 "" GG_STREAMS=1
run s4 "" GG_STREAMS=4
