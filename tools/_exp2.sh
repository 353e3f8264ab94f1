mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_r2c.log 2>&1; echo "pytest exit $?"; grep -n "passed\|failed\|^FAILED" gpurun_out/pytest_r2c.log | tail
grep -n "ReLU masks\|fused vs layer" gpurun_out/pytest_r2c.log | head
run() { echo "=== $*"; env "$@" timeout 150 python profiles/gin_layer_probe.py 6000000 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(' | '.join('%s %.3f' % (k[:28], v['ms']) for k, v in d.items() if isinstance(v, dict)))"; }
run BIGNN_GL_U=2
run BIGNN_GL_THREADS=832 BIGNN_GL_U=2
run BIGNN_GL_THREADS=896 BIGNN_GL_U=2
run BIGNN_GL_THREADS=832 BIGNN_GL_U=3
timeout 100 python profiles/dw_big_probe.py && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dw_tc -s 4 -c 1 -o gpurun_out/dw_tc_r2c python profiles/dw_big_probe.py > gpurun_out/ncu_dw_r2c.log 2>&1; echo "ncu dw exit $?"
