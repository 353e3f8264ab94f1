run() { echo "=== $*"; env "$@" timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(' | '.join('%s %.3f' % (k[:28], v['ms']) for k, v in d.items() if isinstance(v, dict)))"; }
run BIGNN_GL_GATHER=0
run BIGNN_GL_GATHER=1
run BIGNN_GL_GATHER=2
