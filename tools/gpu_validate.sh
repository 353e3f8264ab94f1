#!/bin/bash
# One-call GPU validation of the current tree (run under gpurun on ONE B200):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_validate.sh [tag]'
# 1. the whole `-m gpu` suite; 2. the default bench line; 3. the C4 (--workload ddi_scaled) line with the per-entry
# kernel profile; 4. the late-round probes (streaming kernels at 2 M rows, hub-row SpMM at C4-like skew).
# Everything lands in gpurun_out/ with the tag in the name; nothing here runs under a profiler.
tag=${1:-val}
mkdir -p gpurun_out
BIGNN_RUN_UNVALIDATED=1 timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
timeout 240 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench exit $?"; cut -c1-260 gpurun_out/bench_$tag.json
timeout 240 python bench.py --workload ddi_scaled --steps 5 --warmup 3 --skip-cpu --skip-rooflines \
    > gpurun_out/bench_c4_n1_$tag.json 2> gpurun_out/bench_c4_n1_$tag.err
echo "c4 exit $?"; cut -c1-220 gpurun_out/bench_c4_n1_$tag.json
timeout 60 python profiles/r1b_probe.py > gpurun_out/probe_$tag.log 2>&1; head -6 gpurun_out/probe_$tag.log
timeout 90 python profiles/spmm_hub_probe.py 200000 8000000 > gpurun_out/hub_probe_$tag.log 2>&1; tail -2 gpurun_out/hub_probe_$tag.log
# candidate kernels that are NOT the default yet (written without GPU access at the end of round 1):
#   BIGNN_DW_BM=32  -> k_dw_tc_ring<32,4,3> (4-stage cp.async ring, double-buffered lo operand)
BIGNN_DW_BM=32 timeout 120 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_bwd.py -k "weight_gradient or linear_act or gin_mlp" -x -q \
    > gpurun_out/pytest_dwring_$tag.log 2>&1; echo "dw ring pytest exit $?"; tail -2 gpurun_out/pytest_dwring_$tag.log
BIGNN_DW_BM=32 timeout 60 python profiles/r1b_probe.py 2>&1 | head -1 | sed 's/^/ring: /'
