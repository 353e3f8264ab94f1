mkdir -p gpurun_out
timeout 200 python tools/acc_diag3.py 2>&1 | grep -v Warn | tail -6
run() { echo "=== $*"; env "$@" timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(' | '.join('%s %.3f' % (k[:28], v['ms']) for k, v in d.items() if isinstance(v, dict)))"; }
run X=0
run BIGNN_GL_U=2
run BIGNN_GL_U=4
run BIGNN_GL_THREADS=1024 BIGNN_GL_U=2
run BIGNN_GL_DEBUG=4
run BIGNN_GL_STAGE=1 BIGNN_GL_U=2
run BIGNN_GL_DEBUG=1
run BIGNN_GL_DEBUG=2
run BIGNN_GL_DEBUG=3
BIGNN_GL_TRACE=gpurun_out/gl_trace.txt timeout 150 python profiles/gin_layer_probe.py 6000000 > /dev/null 2>&1; python profiles/gin_layer_trace.py gpurun_out/gl_trace.txt 2>&1 | tail -4
