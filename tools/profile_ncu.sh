# usage (on the GPU box, one GPU): bash tools/profile_ncu.sh <tag>
# 1. launch list of bench.py's step (all kernels, gpu__time_duration only)   -> gpurun_out/<tag>_launches.csv
# 2. ncu --set full of one launch of every pipeline kernel (256 scans/launch) -> gpurun_out/<tag>_top.ncu-rep
tag=${1:-r01b}
export GG_STREAMS=1
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
# one step = 11 launches on one stream; init = 1 launch per slot (256) + table kernels; skip the warm-up steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    -k regex:"k_(rasterize|scan_lo_cells|sort_scatter|sort_scan_hi|cell_stats|detect|spiral|label|roll)" --launch-skip 44 --launch-count 33 \
    $BENCH > gpurun_out/${tag}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/${tag}_top \
    -k regex:"k_(rasterize|sort_scatter|cell_stats|detect|spiral_skew|label|roll_gather)" --launch-skip 32 --launch-count 8 \
    $BENCH > gpurun_out/${tag}_top_bench.log 2>&1
ls -la gpurun_out/${tag}_*
