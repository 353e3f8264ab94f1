mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "batch_norm" 2>&1 | tail -3
timeout 100 python profiles/bn_bwd_probe.py
timeout 600 python -m pytest tests/test_gpu_fused_bwd.py tests/test_gpu_fused_stack.py tests/test_gpu_engine.py tests/test_gpu_step.py -q 2>&1 | tail -3
timeout 900 python bench.py --skip-cpu --skip-gpu-eager > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench exit $?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r2e.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['kernel_profile']['launches_per_step'])
for t in d['kernel_profile']['top'][:9]: print(t['entry'], t['ms_per_call'], t['calls_per_step'], t.get('achieved_gbs'))
PY
timeout 120 python profiles/gin_layer_probe.py 6000000 > gpurun_out/probe6m_r2e.log 2>&1 && tail -1 gpurun_out/probe6m_r2e.log | cut -c1-400 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gin_layer -s 25 -c 1 -o gpurun_out/gin_layer_r2e python profiles/gin_layer_probe.py 6000000 > gpurun_out/ncu_full_r2e.log 2>&1; echo "ncu exit $?"
timeout 300 ncu --set full --clock-control none -k regex:k_bn_bwd_chunk -s 3 -c 1 -o gpurun_out/bn_chunk_r2e python profiles/bn_bwd_probe.py > gpurun_out/ncu_bn_r2e.log 2>&1; echo "ncu bn exit $?"
