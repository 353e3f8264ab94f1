#!/bin/bash
# Trimmed one-call GPU validation (see gpu_validate.sh): the `-m gpu` suite, the default bench line, the config-2 line,
# the fused-layer / weight-gradient / BatchNorm-backward probes, the ncu launch list and `--set full` captures of the
# fused layer kernel and the weight-gradient kernel.
tag=${1:-val}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_$tag.log
grep -n "passed\|failed" gpurun_out/pytest_$tag.log | tail -3
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench exit $?"; cut -c1-300 gpurun_out/bench_$tag.json
timeout 300 python bench.py --workload drugcombo_shape --steps 30 --warmup 5 --skip-gpu-eager --skip-cpu > gpurun_out/bench_c2_$tag.json 2> gpurun_out/bench_c2_$tag.err
echo "c2 exit $?"; cut -c1-220 gpurun_out/bench_c2_$tag.json
timeout 120 python profiles/gin_layer_probe.py 6000000 > gpurun_out/probe6m_$tag.log 2>&1; tail -1 gpurun_out/probe6m_$tag.log | cut -c1-500
timeout 100 python profiles/dw_big_probe.py > gpurun_out/probe_dw_$tag.log 2>&1; tail -1 gpurun_out/probe_dw_$tag.log
timeout 100 python profiles/bn_bwd_probe.py > gpurun_out/probe_bn_$tag.log 2>&1; tail -1 gpurun_out/probe_bn_$tag.log
timeout 300 python bench.py --no-graph --steps 2 --warmup 3 --skip-cpu --skip-gpu-eager --skip-rooflines > gpurun_out/plain_eager_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --no-graph --steps 2 --warmup 3 --skip-cpu --skip-gpu-eager --skip-rooflines > gpurun_out/ncu_launches_$tag.log 2>&1
echo "ncu launch list exit $?"
# the fused layer kernel at the in-step shape: 6 M rows keeping z and t (launch 25 of the probe = the first timed one of that case)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gin_layer -s 25 -c 1 -o gpurun_out/gin_layer_$tag \
    python profiles/gin_layer_probe.py 6000000 > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu full (fused layer) exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dw_tc -s 4 -c 1 -o gpurun_out/dw_tc_$tag \
    python profiles/dw_big_probe.py > gpurun_out/ncu_dw_$tag.log 2>&1
echo "ncu full (weight gradient) exit $?"
