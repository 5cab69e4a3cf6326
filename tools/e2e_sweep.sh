# usage (on the GPU box): bash tools/e2e_sweep.sh  -- e2e knob sweep, results under gpurun_out/
run() { # label, bench-args, env...
  label=$1; shift; bargs=$1; shift
  env "$@" timeout 200 python bench.py --steps 16 --no-cpu-baseline $bargs > gpurun_out/sw_$label.json 2>gpurun_out/sw_$label.err
  python -c "
import json;d=json.load(open('gpurun_out/sw_$label.json'));e=d['e2e'];y=e['synchronous_call'];print('$label', 'overlapped', round(e['value']), round(e['ms_per_step'],2), 'begin', round(e['begin_call_ms'],2), '| sync', round(y['value']), round(y['ms_per_step'],2), 'tail', round(y['ms_after_last_cloud_enqueued'],2), '|', e['scans_repacked_14B'], e['scans_raw_32B'], e['pcie_bytes_per_point'])" || tail -5 gpurun_out/sw_$label.err
}
run warm "" A=1
run base "" A=1
run unit8 "" GG_LAUNCH_UNIT=8
run unit16 "" GG_LAUNCH_UNIT=16
run unit64 "" GG_LAUNCH_UNIT=64
run s8 "" GG_STREAMS=8
run base2 "" A=1
