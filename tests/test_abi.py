"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/hsc_b200.h declares; the product refuses to run without a GPU (no CPU fallback);
the product package never touches oracle/."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_and_load():
    import __graft_entry__ as g
    lib_path = g.build()
    assert os.path.exists(lib_path)
    from hierarchical_sparse_coding_b200 import _native as N
    lib = N.load_library()
    assert lib.hsc_b200_abi_version() == 2


def test_every_declared_symbol_is_exported():
    from hierarchical_sparse_coding_b200 import _native as N
    lib = N.load_library()
    header = open(os.path.join(ROOT, 'include', 'hsc_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(hsc_b200_\w+)\s*\(', header)))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), 'header declares %s but the library does not export it' % name
    assert sorted(N.EXPORTED_SYMBOLS) == declared


def test_struct_layouts_match_header():
    from hierarchical_sparse_coding_b200 import _native as N
    assert ctypes.sizeof(N.MpOptions) == 64
    assert ctypes.sizeof(N.SignalState) == 80


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from hierarchical_sparse_coding_b200 import _native as N
    import hierarchical_sparse_coding_b200 as hsc
    lib = N.load_library()
    h = ctypes.c_void_p()
    assert lib.hsc_b200_create(0, ctypes.byref(h)) == N.HSC_E_CUDA       # the C ABI refuses
    with pytest.raises(RuntimeError):
        hsc.Engine()
    with pytest.raises(RuntimeError):
        hsc.ConvolutionalMatchingPursuit().computeCoefficients(np.zeros(64), np.ones((2, 4)))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'hierarchical_sparse_coding_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('no oracle', ''), '%s mentions the oracle' % f
                assert '/root/reference' not in src
    code = ('import sys; import hierarchical_sparse_coding_b200; '
            'assert not any(m == "oracle" or m.startswith("oracle.") for m in sys.modules), "oracle imported"')
    subprocess.run([sys.executable, '-c', code], check=True, cwd=ROOT)
