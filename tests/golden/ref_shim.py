"""Make the reference package importable in THIS container (test infrastructure only).

The reference (/root/reference/src/pylabfea) imports matplotlib at module top
(model.py:14,17; material.py:35,38) and asks importlib.metadata for its own
version (__init__.py:19).  Neither is available here, so we install stub
modules and patch the version lookup, then put the reference's src/ on
sys.path.  Only the golden-vector generator (make_golden.py) uses this; nothing
under tests/ that runs on the GPU box may import it (/root/reference does not
exist there).
"""
import sys
import types
import importlib.metadata as _md

REF_SRC = '/root/reference/src'


def _stub(name):
    m = types.ModuleType(name)
    m.__path__ = []

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, k):
            return _Any()

    def _ga(k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Any()
    m.__getattr__ = _ga
    sys.modules[name] = m
    return m


def import_reference():
    for n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'matplotlib.colorbar',
              'matplotlib.lines', 'matplotlib.cm', 'matplotlib.patches', 'ipympl'):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                _stub(n)
    _orig = _md.version

    def _version(name):
        if name == 'pylabfea':
            return '4.4.2'
        return _orig(name)
    _md.version = _version
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import pylabfea
    return pylabfea
