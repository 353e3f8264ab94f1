"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every entry point that
include/bignn_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import bignn_b200 as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    src = open(os.path.join(ROOT, 'include', 'bignn_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(bignn_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_loads_and_exports_header():
    B._lib.build()
    lib = ctypes.CDLL(B._lib.LIB_PATH)
    names = declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(B._lib.SIGNATURES.keys())
    lib.bignn_abi_version.restype = ctypes.c_int
    assert lib.bignn_abi_version() == B._lib.ABI_VERSION == 4
    lib.bignn_error_string.restype = ctypes.c_char_p
    assert b'invalid' in lib.bignn_error_string(-1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'bilevel-graph-neural-network_b200')
    for f in os.listdir(pkg):
        if f.endswith('.py'):
            for line in open(os.path.join(pkg, f)):
                if re.match(r'\s*(from|import)\s', line):
                    assert 'oracle' not in line, (f, line)
