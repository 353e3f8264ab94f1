mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_r2b.log 2>&1; echo "pytest exit $?"; grep -n "passed\|failed\|^FAILED" gpurun_out/pytest_r2b.log | tail
for v in "BIGNN_GL_THREADS=768" "BIGNN_GL_TMEM_A=0" "BIGNN_GL_THREADS=768 BIGNN_GL_TMEM_A=0" "BIGNN_GL_U=2"; do
  echo "=== $v"; env $v timeout 200 python -m pytest tests/test_gpu_fused_stack.py -q -s -k test_fused_stack_step_equals 2>&1 | grep -E "passed|failed|AssertionError|worst" | head -5
  cp gpurun_out/grad_errors_gin_gcn_fused_engine.txt "gpurun_out/grad_errors_fused_$(echo $v | tr ' =' '__').txt"
done
timeout 100 python tools/acc_diag2.py 2>&1 | tail -8
timeout 100 python tools/fwd_err_diag.py 2>&1 | tail -8
