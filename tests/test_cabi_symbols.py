"""CPU tier: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/alona_b200.h declares; no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope='module')
def lib():
    from alona_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared():
    src = open(os.path.join(ROOT, 'include', 'alona_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ab_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), 'missing export %s' % n


def test_ctypes_signatures_cover_header():
    from alona_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_version_and_clean_failure_without_gpu(lib, has_cuda):
    assert lib.ab_version() == 100
    if has_cuda:
        pytest.skip('GPU present')
    h = ctypes.c_void_p()
    rc = lib.ab_ctx_create(ctypes.byref(h), 0, 0, 1)
    assert rc != 0
    assert b'no CPU fallback' in lib.ab_last_error() or b'CUDA' in lib.ab_last_error()


def test_engine_refuses_to_run_without_gpu(has_cuda):
    if has_cuda:
        pytest.skip('GPU present')
    from alona_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'alona_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(root, f)).read()
                assert 'oracle' not in txt, '%s mentions the oracle' % f
                assert '/root/reference' not in txt


def test_sass_is_sm100a(lib):
    import subprocess
    from alona_b200 import _lib
    out = subprocess.run(['cuobjdump', '-lelf', _lib.lib_path()], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
