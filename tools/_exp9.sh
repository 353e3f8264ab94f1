mkdir -p gpurun_out
BIGNN_GL_TRACE=gpurun_out/gl_trace_e2.txt timeout 150 python profiles/gin_layer_probe.py 6000000 > /dev/null 2>&1; python profiles/gin_layer_trace.py gpurun_out/gl_trace_e2.txt 2>/dev/null | tail -4 | cut -c1-420
BIGNN_GL_DEBUG=2 BIGNN_GL_TRACE=gpurun_out/gl_trace_e2_nostore.txt timeout 150 python profiles/gin_layer_probe.py 6000000 > /dev/null 2>&1; python profiles/gin_layer_trace.py gpurun_out/gl_trace_e2_nostore.txt 2>/dev/null | tail -4 | cut -c1-420
