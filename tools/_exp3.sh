mkdir -p gpurun_out
timeout 100 python profiles/dw_probe.py 2>&1 | grep -v "first row\|^  " | tail -9
timeout 100 python profiles/dw_big_probe.py
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_bwd.py tests/test_gpu_fused_stack.py tests/test_gpu_engine.py -q 2>&1 | tail -3
timeout 900 python bench.py --skip-cpu --skip-gpu-eager > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench exit $?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r2d.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])
for t in d['kernel_profile']['top'][:9]: print(t['entry'], t['ms_per_call'], t['calls_per_step'], t.get('achieved_gbs'))
PY
