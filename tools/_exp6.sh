mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_fused_stack.py -q -x 2>&1 | tail -3
for v in 1 0; do echo "EARLY=$v"; BIGNN_GL_EARLY=$v timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | cut -c1-420; done
BIGNN_GL_THREADS=768 timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | cut -c60-420
BIGNN_GL_THREADS=832 timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | cut -c60-420
BIGNN_GL_THREADS=768 BIGNN_GL_U=3 timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | cut -c60-420
BIGNN_GL_TRACE=gpurun_out/gl_trace_early.txt timeout 150 python profiles/gin_layer_probe.py 6000000 > /dev/null 2>&1; python profiles/gin_layer_trace.py gpurun_out/gl_trace_early.txt 2>/dev/null | tail -22 | cut -c1-200
