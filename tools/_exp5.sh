for v in 0 1 2 3; do echo "variant $v"; BIGNN_BN_CL_VARIANT=$v timeout 100 python profiles/bn_bwd_probe.py; done
