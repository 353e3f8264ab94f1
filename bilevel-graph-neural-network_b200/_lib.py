"""ctypes binding of the C-ABI in include/bignn_b200.h.

The shared library (`_C/libbignn_b200.so`, built in-tree by `build()` /
`__graft_entry__.build()` with nvcc for sm_100a) is the ONLY compute backend of
this package: there is no CPU or eager-PyTorch fallback.  A missing library is
an ImportError at first use, a failed call a RuntimeError carrying the library's
own error string.
"""
import ctypes
import os
import subprocess
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, '_C')
LIB_PATH = os.path.join(LIB_DIR, 'libbignn_b200.so')
HEADER = os.path.join(ROOT, 'include', 'bignn_b200.h')

ABI_VERSION = 4          # include/bignn_b200.h BIGNN_ABI_VERSION (3: fused GIN layer, folded readout; 4: fused pair decoder)

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']

# name -> (restype, argument codes): p pointer, i int32, l int64, f float; a trailing
# 's' marks the stream argument, filled in from torch's current stream.
SIGNATURES = {
    'bignn_abi_version': ('i', ''),
    'bignn_error_string': ('z', 'i'),
    'bignn_launch_count': ('l', ''),
    'bignn_merge_build_workspace_bytes': ('l', 'i'),
    'bignn_merge_build': ('i', 'pppp' 'i' 'p' 'i' 'pp' 'ppp' 'ppp' 'ii' 'pl' 's'),
    'bignn_gcn_dinv': ('i', 'ppips'),
    'bignn_spmm_f32': ('i', 'pp' 'pl' 'pl' 'iiif' 'ppi' 's'),
    'bignn_spmm_rows_f32': ('i', 'pp' 'pl' 'pl' 'iiiif' 'ppi' 's'),
    'bignn_spmm_planned_rows_f32': ('i', 'pp' 'ppii' 'pii' 'pl' 'pl' 'iiiif' 'ppi' 'pl' 's'),
    'bignn_spmm_planned_workspace_bytes': ('l', 'ii'),
    'bignn_spmm_planned_f32': ('i', 'pp' 'ppii' 'pi' 'pl' 'pl' 'iiif' 'ppi' 'pl' 's'),
    'bignn_gemm_workspace_bytes': ('l', 'iiii'),
    'bignn_gemm_f32': ('i', 'iiiii' 'pl' 'pl' 'pl' 'pi' 'pl' 's'),
    'bignn_gemm_tc_supported': ('i', 'iii'),
    'bignn_gemm_tc_f32': ('i', 'iii' 'pl' 'pli' 'pl' 'pi' 's'),
    'bignn_gemm_tc_masked_f32': ('i', 'iii' 'pl' 'pli' 'pl' 'pi' 'pli' 's'),
    'bignn_dw_tc_supported': ('i', 'iii'),
    'bignn_dw_tc_workspace_bytes': ('l', 'iii'),
    'bignn_dw_tc_f32': ('i', 'iii' 'pl' 'pl' 'p' 'ip' 'pl' 's'),
    'bignn_colsum_workspace_bytes': ('l', 'ii'),
    'bignn_colsum_f32': ('i', 'pliip' 'pl' 's'),
    'bignn_act_bwd_f32': ('i', 'ppplis'),
    'bignn_bn_workspace_bytes': ('l', 'iii'),
    'bignn_bn_seg_fwd': ('i', 'plpl' 'piii' 'pp' 'ff' 'ppp' 'pp' 'p' 'pl' 's'),
    'bignn_bn_running_update': ('i', 'ppiifppp' 's'),
    'bignn_bn_eval_fwd': ('i', 'plpl' 'ii' 'pp' 'f' 'pp' 's'),
    'bignn_bn_seg_bwd': ('i', 'plplpl' 'piii' 'ppp' 'pp' 'i' 'pl' 's'),
    'bignn_bn_rows_workspace_bytes': ('l', 'ii'),
    'bignn_bn_rows_sums': ('i', 'plpl' 'iii' 'ppp' 'pl' 's'),
    'bignn_bn_rows_fwd_apply': ('i', 'plpl' 'iii' 'pl' 'pp' 'ff' 'ppp' 'pp' 's'),
    'bignn_bn_rows_bwd_apply': ('i', 'plplpl' 'iii' 'ppp' 'pl' 'i' 's'),
    'bignn_readout_fwd': ('i', 'pl' 'pii' 'i' 'p' 'pli' 's'),
    'bignn_readout_bwd': ('i', 'pli' 'p' 'pii' 'i' 'pl' 'i' 's'),
    'bignn_readout_gated_fwd': ('i', 'plpl' 'pii' 'pl' 's'),
    'bignn_readout_gated_bwd': ('i', 'plpl' 'pl' 'pii' 'plpl' 's'),
    'bignn_readout_fold_fwd': ('i', 'pl' 'pii' 'i' 'p' 'pppp' 'pli' 's'),
    'bignn_pair_decoder_supported': ('i', 'iiiii'),
    'bignn_pair_decoder_workspace_bytes': ('l', 'iiiiii'),
    'bignn_pair_decoder_fwd': ('i', 'pl' 'piii' 'ppi' 'ppi' 'ppi' 'i' 'pp' 'pl' 'ppp' 'p' 'pl' 's'),
    'bignn_pair_decoder_bwd': ('i', 'pl' 'piii' 'pi' 'pi' 'pi' 'i' 'pp' 'pl' 'ppp' 'p' 'pl' 'pppppp' 'pl' 's'),
    'bignn_gin_layer_supported': ('i', 'ii'),
    'bignn_gin_layer_stat_records': ('l', 'ii'),
    'bignn_gin_layer_fwd': ('i', 'iii' 'ppip' 'pl' 'ppp' 'pip' 'f' 'pppp' 'ii' 'pl' 'pl' 'pl' 'p' 's'),
    'bignn_gin_bn_finalize': ('i', 'ppii' 'f' 'p' 'pp' 'p' 'p' 's'),
    'bignn_pair_gather_norm_fwd': ('i', 'pl' 'pii' 'pl' 'p' 's'),
    'bignn_pair_gather_norm_bwd': ('i', 'pl' 'pii' 'pl' 'p' 'pl' 's'),
    'bignn_bce_fwd': ('i', 'ppips'),
    'bignn_bce_bwd': ('i', 'ppipps'),
    'bignn_bce_logits_fwd': ('i', 'ppips'),
    'bignn_bce_logits_bwd': ('i', 'ppipps'),
    'bignn_gat_fwd_workspace_bytes': ('l', 'ii'),
    'bignn_gat_fwd': ('i', 'pp' 'ppii' 'pi' 'iii' 'pl' 'pp' 'fi' 'pl' 'p' 'pl' 's'),
    'bignn_gat_bwd_workspace_bytes': ('l', 'iii'),
    'bignn_gat_bwd': ('i', 'pp' 'ppii' 'pi' 'iii' 'pl' 'pp' 'fi' 'pl' 'pl' 'p' 'pl' 'p' 'pl' 's'),
    'bignn_act_fwd_f32': ('i', 'pplis'),
    'bignn_add_f32': ('i', 'pppls'),
    'bignn_prelu_fwd_f32': ('i', 'pplipis'),
    'bignn_prelu_bwd_f32': ('i', 'pppplipis'),
    'bignn_rownorm_fwd_f32': ('i', 'plpliips'),
    'bignn_rownorm_bwd_f32': ('i', 'plplppliis'),
    'bignn_gate_mul_fwd_f32': ('i', 'pppls'),
    'bignn_gate_mul_bwd_f32': ('i', 'pppppls'),
    'bignn_pair_dot_fwd_f32': ('i', 'pliipis'),
    'bignn_pair_dot_bwd_f32': ('i', 'pliippipls'),
    'bignn_ce_fwd': ('i', 'plpiips'),
    'bignn_ce_bwd': ('i', 'plpiippls'),
}

_CT = {'p': ctypes.c_void_p, 'i': ctypes.c_int32, 'l': ctypes.c_int64, 'f': ctypes.c_float,
       's': ctypes.c_void_p, 'z': ctypes.c_char_p}

_lib = None
_lock = threading.Lock()
_backend_override = None     # tests only: an object with .call(name, *args)


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


STAMP_PATH = LIB_PATH + '.srchash'


def source_hash():
    """Content hash of everything the library is built from (kernels, shared headers, the C-ABI header, the flags):
    file times mean nothing after a snapshot copy to another box."""
    import hashlib
    h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    deps = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')) + [HEADER]
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    """nvcc cross-compiles every kernel for sm_100a into the in-tree .so (no GPU needed)."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    if not os.path.exists(nvcc):
        nvcc = 'nvcc'
    tmp = LIB_PATH + '.tmp.%d' % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + ['-o', tmp] + sources()
    if verbose:
        print(' '.join(cmd))
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout)
    os.replace(tmp, LIB_PATH)
    with open(STAMP_PATH, 'w') as f:
        f.write(source_hash())
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                'bignn_b200: CUDA library {} is missing -- run `python -c "import __graft_entry__ as g; '
                'g.build()"` (nvcc, sm_100a). There is no CPU fallback.'.format(LIB_PATH))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)        # AttributeError if the symbol is not exported
            fn.restype = _CT[res]
            fn.argtypes = [_CT[c] for c in args]
        if lib.bignn_abi_version() != ABI_VERSION:
            raise ImportError('bignn_b200: ABI version mismatch (library {}, binding {}): rebuild with '
                              '__graft_entry__.build()'.format(lib.bignn_abi_version(), ABI_VERSION))
        _lib = lib
    return _lib


def error_string(code):
    return load().bignn_error_string(int(code)).decode()


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr() if a.numel() > 0 else None
    return int(a)


def call(name, *args):
    """Invoke a C-ABI entry point.  Tensors become raw pointers; the stream argument
    is torch's current CUDA stream.  Non-zero status -> RuntimeError."""
    if _backend_override is not None:
        return _backend_override.call(name, *args)
    lib = load()
    res, codes = SIGNATURES[name]
    n = len(codes) - (1 if codes.endswith('s') else 0)
    if len(args) != n:
        raise TypeError('{} expects {} arguments, got {}'.format(name, n, len(args)))
    cargs = []
    for c, a in zip(codes, args):
        cargs.append(_ptr(a) if c == 'p' else a)
    if codes.endswith('s'):
        cargs.append(torch.cuda.current_stream().cuda_stream)
    out = getattr(lib, name)(*cargs)
    if res == 'i' and name not in ('bignn_abi_version', 'bignn_gemm_tc_supported', 'bignn_dw_tc_supported', 'bignn_gin_layer_supported',
                                  'bignn_pair_decoder_supported') and out != 0:
        raise RuntimeError('{} failed with status {}: {}'.format(name, out, error_string(out)))
    return out


def launch_count():
    if _backend_override is not None:
        return 0
    return int(load().bignn_launch_count())


def require_device(*tensors):
    """Every operand of the product path must live on a CUDA device."""
    if _backend_override is not None:
        return
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('bignn_b200 runs on CUDA devices only (got a {} tensor); '
                               'there is no CPU path'.format(t.device))
