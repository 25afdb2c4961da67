"""CPU: the C-ABI library builds/loads and exports every symbol include/diffndm_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib_path():
    from diffndm_b200 import build
    return build.build()


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, 'include', 'diffndm_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(dndm_[a-z_0-9]+)\s*\(', hdr)))
    assert len(declared) >= 12
    lib = ctypes.CDLL(_lib_path())
    for s in declared:
        assert hasattr(lib, s), f'{s} declared in the header but not exported'
    from diffndm_b200.engine import EXPORTED_SYMBOLS
    assert sorted(EXPORTED_SYMBOLS) == declared


def test_version_and_argument_errors_without_gpu():
    lib = ctypes.CDLL(_lib_path())
    lib.dndm_version.restype = ctypes.c_char_p
    assert b'sm_100a' in lib.dndm_version()
    lib.dndm_last_error.restype = ctypes.c_char_p
    # NULL arguments are rejected before any CUDA call
    assert lib.dndm_engine_create(None, None) == -1
    assert b'null' in lib.dndm_last_error()
    assert lib.dndm_egnn_forward(None, None, None, None, 0, None, None, 0, 0, 0, None, None, None) == -1


def test_engine_refuses_cpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from diffndm_b200.engine import Engine
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Engine()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, 'diffndm_b200')
    for f in os.listdir(pkg):
        if f.endswith('.py'):
            src = open(os.path.join(pkg, f)).read()
            assert 'import oracle' not in src and 'from oracle' not in src, f
