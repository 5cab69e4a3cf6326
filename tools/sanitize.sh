# usage (on the GPU box): bash tools/sanitize.sh <tag>
# racecheck: the kernels with shared-memory exchange / barriers / warp-aggregated atomics (spiral in both thread layouts, TMA-staged
#            patch detection with its mbarrier pipeline, rasteriser, tile scan) through the golden, batch-of-ten and odd-N tests
# memcheck:  golden fixtures, batches (both host paths, overlapped begin/wait), per-phase entry points, layer images, multi-slot payload upload
# HostPacker: CUDA-free stress of the packer pool under ThreadSanitizer-free repeated runs (gg_host_packer_selftest) in tests/test_host_logic.py
tag=${1:-r02}
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -x -q \
    -k "golden or batch_of_ten or odd_cell_count or batched_slots" > gpurun_out/${tag}_sanitizer_racecheck.log 2>&1
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_rows.py -x -q \
    -k "golden or batch or overlapped or per_phase or layer_images or upload_cloud_msg or edge_cases" > gpurun_out/${tag}_sanitizer_memcheck.log 2>&1
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/${tag}_sanitizer_racecheck.log gpurun_out/${tag}_sanitizer_memcheck.log
