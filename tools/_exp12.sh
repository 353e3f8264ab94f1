mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_stack.py tests/test_gpu_fused_bwd.py -q -x 2>&1 | tail -2
timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | cut -c60-420
BIGNN_GL_THREADS=832 timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | cut -c60-420
BIGNN_GL_THREADS=768 timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | cut -c60-420
timeout 100 python profiles/dw_big_probe.py
timeout 100 python profiles/gemm_tc_probe.py 2>&1 | tail -4
BIGNN_GL_TRACE=gpurun_out/gl_trace_lds.txt timeout 150 python profiles/gin_layer_probe.py 6000000 > /dev/null 2>&1; python profiles/gin_layer_trace.py gpurun_out/gl_trace_lds.txt 2>/dev/null | tail -4 | cut -c1-420
