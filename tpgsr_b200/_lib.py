"""ctypes binding of libtpgsr_b200.so — signatures are generated from include/tpgsr_b200.h (single source of truth).

There is NO fallback: if the shared library is missing the import fails with instructions, and every call that
returns a non-zero status raises.  (The oracle under ``oracle/`` is test infrastructure and is never imported here.)
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libtpgsr_b200.so')
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'tpgsr_b200.h')

_ERRORS = {-1: 'TPGSR_ERR_BAD_ARG', -2: 'TPGSR_ERR_LAUNCH', -3: 'TPGSR_ERR_WORKSPACE', -4: 'TPGSR_ERR_UNSUPPORTED',
           -5: 'TPGSR_ERR_DRIVER'}

_CTYPES = {'int': ctypes.c_int, 'long': ctypes.c_long, 'long long': ctypes.c_longlong, 'size_t': ctypes.c_size_t, 'float': ctypes.c_float,
           'double': ctypes.c_double, 'cudaStream_t': ctypes.c_void_p}


class TpgsrError(RuntimeError):
    pass


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[str]]]:
    """{name: (return_type, [arg_type, ...])} for every ``tpgsr_*`` declaration in the header."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    decls = {}
    for m in re.finditer(r'\b(int|size_t)\s+(tpgsr_\w+)\s*\(([^)]*)\)\s*;', text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        types = []
        for a in args.split(','):
            a = a.strip()
            if not a or a == 'void':
                continue
            if '*' in a:
                types.append('ptr')
            else:
                types.append(a.rsplit(' ', 1)[0].replace('const ', '').strip())
        decls[name] = (ret, types)
    return decls


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} not found. tpgsr_b200 has no CPU/PyTorch fallback: build the CUDA library first with '
            f'`python tpgsr_b200/build.py` (or `python -c "import __graft_entry__ as g; g.build()"`).')
    lib = ctypes.CDLL(LIB_PATH)
    decls = parse_header()
    for name, (ret, types) in decls.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype = ctypes.c_size_t if ret == 'size_t' else ctypes.c_int
        fn.argtypes = [ctypes.c_void_p if t == 'ptr' else _CTYPES[t] for t in types]
    return lib, decls


LIB, DECLS = _load()


def call(name: str, *args):
    """Invoke an int-returning entry point; raise on a non-zero status."""
    rc = getattr(LIB, name)(*args)
    if rc != 0:
        raise TpgsrError(f'{name} failed: {_ERRORS.get(rc, rc)}')


def query(name: str, *args) -> int:
    """Invoke a size_t-returning *_workspace_bytes query."""
    return int(getattr(LIB, name)(*args))
