"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/plfb200.h declares,
fails loudly without a GPU, and the product never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        cuda = ctypes.CDLL('libcuda.so.1')
        n = ctypes.c_int()
        return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def test_library_exports_every_declared_symbol():
    from pylabfea_b200 import _lib
    L = _lib.load()
    hdr = open(os.path.join(ROOT, 'include', 'plfb200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    names = sorted(set(re.findall(r'\b(plf_[a-z0-9_]+)\s*\(', hdr)))
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert b'sm_100a' in L.plf_version()


def test_material_struct_layout_matches():
    from pylabfea_b200 import _lib
    L = _lib.load()
    assert L.plf_sizeof_material() == ctypes.sizeof(_lib.PlfMaterial)


@pytest.mark.skipif(_has_gpu(), reason='CUDA device present')
def test_no_cpu_fallback_without_gpu():
    from pylabfea_b200 import _lib
    with pytest.raises(_lib.PlfError) as ei:
        _lib.Context(0)
    assert 'no CPU fallback' in str(ei.value) or 'CUDA' in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'pylabfea_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(root, f)).read()
                assert 'oracle' not in src.lower().replace('no cpu fallback', ''), os.path.join(root, f)


def test_header_is_plain_c(tmp_path):
    """include/plfb200.h is the whole boundary: it must compile as C99 on its own (no C++ or CUDA types leak through)."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    src = tmp_path / 'use_header.c'
    src.write_text('#include "plfb200.h"\nint main(void) { plf_material m; (void)m; return PLF_MAT_SVC3 == 5 ? 0 : 1; }\n')
    r = subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-pedantic', '-fsyntax-only', '-I', os.path.join(ROOT, 'include'), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
