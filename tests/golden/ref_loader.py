"""Import the UNMODIFIED Python-2 reference (`/root/reference/hsc`) under Python 3.

TEST INFRASTRUCTURE ONLY.  Used by `tests/golden/make_golden.py` (to generate the committed
golden vectors), by the `not gpu` tests that pin `oracle/` against the live reference, by the
drop-in tests that drive the B200 approximator from the reference's own coder / learner classes,
and by `bench.py --impl reference` (the CPU arm runs `hsc.modeling` itself).  In the dev container
the reference is read from `/root/reference`; the GPU box has no such path and receives the
git-ignored copy `baseline/_ref/hsc` (baseline/install_reference.py) with the repo snapshot.
Nothing in the product package imports this file.

How it works: a meta-path finder loads `hsc.*` from the read-only reference tree and rewrites,
at AST level, every binary `a / b` into `__py2div__(a, b)` (floor division iff both operands are
integers / integer arrays, as Python 2 did).  The remaining py2-isms are satisfied by aliases
(cPickle, StringIO, itertools.izip, collections.Iterable, np.int/np.float/np.Inf,
np.unravel_index(dims=), np.issubdtype(x, float)) and a stub matplotlib.  No reference source is
copied or edited.
"""
import ast
import collections
import collections.abc
import importlib.abc
import importlib.machinery
import io
import itertools
import os
import pickle
import sys
import types

import numpy as np

def _default_root():
    """HSC_REFERENCE_ROOT, else the read-only source tree of the dev container, else the git-ignored copy the GPU box
    receives with the repo snapshot (baseline/_ref, written by baseline/install_reference.py)."""
    env = os.environ.get('HSC_REFERENCE_ROOT')
    if env:
        return env
    if os.path.isfile('/root/reference/hsc/modeling.py'):
        return '/root/reference'
    return os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'baseline', '_ref')


REFERENCE_ROOT = _default_root()


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'hsc', 'modeling.py'))


def _is_int_like(v):
    if isinstance(v, (bool, np.bool_)):
        return True
    if isinstance(v, (int, np.integer)):
        return True
    if isinstance(v, np.ndarray) and v.dtype.kind in 'iub':
        return True
    return False


def __py2div__(a, b):
    if _is_int_like(a) and _is_int_like(b):
        return a // b
    return a / b


class _DivRewriter(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            call = ast.Call(func=ast.Name(id='__py2div__', ctx=ast.Load()),
                            args=[node.left, node.right], keywords=[])
            return ast.copy_location(call, node)
        return node


class _Py2Loader(importlib.machinery.SourceFileLoader):
    def source_to_code(self, data, path, *, _optimize=-1):
        tree = ast.parse(data, filename=path)
        tree = _DivRewriter().visit(tree)
        ast.fix_missing_locations(tree)
        return compile(tree, path, 'exec', dont_inherit=True, optimize=_optimize)

    def exec_module(self, module):
        module.__dict__['__py2div__'] = __py2div__
        super().exec_module(module)

    # never write .pyc next to the read-only reference
    def set_data(self, path, data, *, _mode=0o666):
        return None

    def get_code(self, fullname):
        source_path = self.get_filename(fullname)
        source_bytes = self.get_data(source_path)
        return self.source_to_code(source_bytes, source_path)


class _Py2Finder(importlib.abc.MetaPathFinder):
    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path, target=None):
        if fullname != 'hsc' and not fullname.startswith('hsc.'):
            return None
        parts = fullname.split('.')
        base = os.path.join(self.root, *parts)
        if os.path.isdir(base):
            init = os.path.join(base, '__init__.py')
            return importlib.machinery.ModuleSpec(
                fullname, _Py2Loader(fullname, init), origin=init, is_package=True,
                loader_state=None) if os.path.isfile(init) else None
        src = base + '.py'
        if os.path.isfile(src):
            return importlib.machinery.ModuleSpec(fullname, _Py2Loader(fullname, src), origin=src)
        return None


def _install_aliases():
    sys.modules.setdefault('cPickle', pickle)
    if 'StringIO' not in sys.modules:
        m = types.ModuleType('StringIO')
        m.StringIO = io.StringIO
        sys.modules['StringIO'] = m
    if not hasattr(itertools, 'izip'):
        itertools.izip = zip
    for name in ('Iterable', 'Mapping', 'Sequence'):
        if not hasattr(collections, name):
            setattr(collections, name, getattr(collections.abc, name))
    for name, val in (('int', int), ('float', float), ('bool', bool), ('Inf', np.inf)):
        if name not in np.__dict__:
            setattr(np, name, val)
    if not getattr(np.unravel_index, '_py2shim', False):
        _orig_unravel = np.unravel_index

        def unravel_index(indices, shape=None, order='C', dims=None):
            if shape is None:
                shape = dims
            return _orig_unravel(indices, shape, order=order)
        unravel_index._py2shim = True
        np.unravel_index = unravel_index
    if not getattr(np.issubdtype, '_py2shim', False):
        _orig_issub = np.issubdtype

        def issubdtype(a, b):
            if b is float:
                b = np.floating
            elif b is int:
                b = np.integer
            return _orig_issub(a, b)
        issubdtype._py2shim = True
        np.issubdtype = issubdtype
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            class _Anything(types.ModuleType):
                def __getattr__(self, item):
                    if item.startswith('__'):
                        raise AttributeError(item)

                    def _f(*a, **k):
                        return None
                    return _f
            mpl = _Anything('matplotlib')
            plt = _Anything('matplotlib.pyplot')
            mpl.pyplot = plt
            mpl.rcParams = {}
            plt.rcParams = mpl.rcParams
            sys.modules['matplotlib'] = mpl
            sys.modules['matplotlib.pyplot'] = plt


_installed = False


def load_reference():
    """Returns the reference `hsc` package (modeling, utils, dataset, analysis importable)."""
    global _installed
    if not reference_available():
        raise ImportError('reference tree not found at %s' % REFERENCE_ROOT)
    if not _installed:
        _install_aliases()
        sys.meta_path.insert(0, _Py2Finder(REFERENCE_ROOT))
        _installed = True
    import hsc  # noqa: F401
    import hsc.utils  # noqa: F401
    import hsc.dataset  # noqa: F401
    import hsc.modeling  # noqa: F401
    return sys.modules['hsc']
