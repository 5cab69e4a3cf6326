# usage (on the GPU box): bash tools/e2e_sweep.sh  -- knob sweep, results under gpurun_out/
run() { # label, bench-args, env...
  label=$1; shift; bargs=$1; shift
  env "$@" timeout 200 python bench.py --steps 16 --no-cpu-baseline $bargs > gpurun_out/sw_$label.json 2>gpurun_out/sw_$label.err
  python -c "
import json;d=json.load(open('gpurun_out/sw_$label.json'));e=d['e2e'];y=e['synchronous_call'];b=e['begin_call_breakdown_ms'];print('$label', 'e2e', round(e['value']), round(e['ms_per_step'],2), 'begin', round(e['begin_call_ms'],2), 'pack', round(b['packer_threads_packing_sum'],1), 'slotwait', round(b['packer_threads_waiting_for_slot_sum'],1), 'idle', round(b['feeder_nothing_to_enqueue'],2), '| sync', round(y['value']), '|', e['scans_repacked_14B'], e['scans_raw_32B'], e['pcie_bytes_per_point'])" || tail -5 gpurun_out/sw_$label.err
}
run clwb "" GG_RAW_GATE=2
run cached "" GG_RAW_GATE=2 GG_PACK_STORE=1
run nt "" GG_RAW_GATE=2 GG_PACK_STORE=0
run nt_r64 "" GG_RAW_GATE=2 GG_PACK_STORE=0 GG_PACK_RING=64
run clwb_r16 "" GG_RAW_GATE=2 GG_PACK_RING=16
run clwb_c8 "" GG_RAW_GATE=2 GG_COPY_STREAMS=8
run clwb2 "" GG_RAW_GATE=2
