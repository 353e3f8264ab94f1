#!/bin/bash
# One-call GPU validation of the current tree (run under gpurun on ONE B200):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh [tag]'
# 1. the whole `-m gpu` suite (with -s: the parity reports are part of the log); 2. the flake hunt; 3. the default bench
# line (BASELINE config 4) with the CPU and GPU-eager baselines, the reference arm, the config-2 line; 4. the ncu launch
# list of two eager steps and one `--set full` capture of the fused layer kernel -- each only after the identical command
# has exited 0 without ncu.  Everything lands in gpurun_out/ with the tag in the name.
tag=${1:-val}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_$tag.log
grep -n "passed\|failed" gpurun_out/pytest_$tag.log | tail -3
for i in 1 2 3; do timeout 200 python tools/flake_hunt.py 300 2>&1 | grep -v Warn | tail -6; done > gpurun_out/flake_hunt_$tag.log; tail -2 gpurun_out/flake_hunt_$tag.log
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench exit $?"; cut -c1-260 gpurun_out/bench_$tag.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
echo "reference arm exit $?"; cut -c1-200 gpurun_out/bench_ref_$tag.json
timeout 600 python bench.py --workload drugcombo_shape --steps 30 --warmup 5 --skip-gpu-eager > gpurun_out/bench_c2_$tag.json 2> gpurun_out/bench_c2_$tag.err
echo "c2 exit $?"; cut -c1-220 gpurun_out/bench_c2_$tag.json
timeout 300 python bench.py --no-graph --steps 2 --warmup 3 --skip-cpu --skip-gpu-eager --skip-rooflines > gpurun_out/plain_eager_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --no-graph --steps 2 --warmup 3 --skip-cpu --skip-gpu-eager --skip-rooflines > gpurun_out/ncu_launches_$tag.log 2>&1
echo "ncu launch list exit $?"
timeout 120 python profiles/gin_layer_probe.py 2000000 > gpurun_out/probe_plain_$tag.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gin_layer -s 12 -c 2 -o gpurun_out/gin_layer_$tag \
    python profiles/gin_layer_probe.py 2000000 > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu full exit $?"; tail -1 gpurun_out/probe_plain_$tag.log | cut -c1-400
