mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused_stack.py tests/test_gpu_fused_bwd.py -q -x 2>&1 | tail -2
run() { echo "=== $*"; env "$@" timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(' | '.join('%s %.3f' % (k[:28], v['ms']) for k, v in d.items() if isinstance(v, dict)))"; }
run X=0
run BIGNN_GL_TMA_OUT=0
run BIGNN_GL_THREADS=832
run BIGNN_GL_THREADS=768
BIGNN_GL_TRACE=gpurun_out/gl_trace_tma.txt timeout 150 python profiles/gin_layer_probe.py 6000000 > /dev/null 2>&1; python profiles/gin_layer_trace.py gpurun_out/gl_trace_tma.txt 2>/dev/null | tail -4 | cut -c1-420
