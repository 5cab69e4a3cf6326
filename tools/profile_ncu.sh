# usage (on the GPU box, one GPU): bash tools/profile_ncu.sh <tag> [full]
# 1. launch list of bench.py's step (all kernels, gpu__time_duration only)   -> gpurun_out/<tag>_launches.csv
# 2. (with "full") ncu --set full of one whole step (same filter and window as the launch list: the 10 launches of the first timed step) -> gpurun_out/<tag>_top.ncu-rep
tag=${1:-r02a}
export GG_STREAMS=1
BENCH="python bench.py --steps 2 --warmup 3 --pool 2 --no-e2e --no-cpu-baseline --no-extras"
# one step = 10 launches on one stream (8 for the scan pipeline + 2 for the map roll); skip the warm-up steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    -k regex:"k_(rasterize|cell_tiles|cell_place|scatter|cell_stats|detect|spiral|label|roll)" --launch-skip 28 --launch-count 20 \
    $BENCH > gpurun_out/${tag}_launches_bench.log 2>&1
if [ "$2" = "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/${tag}_top \
    -k regex:"k_(rasterize|cell_tiles|cell_place|scatter|cell_stats|detect|spiral|label|roll)" --launch-skip 28 --launch-count 10 \
    $BENCH > gpurun_out/${tag}_top_bench.log 2>&1
fi
ls -la gpurun_out/${tag}_*
