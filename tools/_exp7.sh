mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_fused_stack.py -q -x 2>&1 | tail -2
run() { echo "=== $*"; env "$@" timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(' | '.join('%s %.3f' % (k[:28], v['ms']) for k, v in d.items() if isinstance(v, dict)))"; }
run X=0
run BIGNN_GL_DEBUG=8
run BIGNN_GL_THREADS=1024
run BIGNN_GL_THREADS=960
run BIGNN_GL_THREADS=768
BIGNN_GL_TRACE=gpurun_out/gl_trace_pf.txt timeout 150 python profiles/gin_layer_probe.py 6000000 > /dev/null 2>&1; python profiles/gin_layer_trace.py gpurun_out/gl_trace_pf.txt 2>/dev/null | tail -3 | cut -c1-400
