"""Shadow packages: run the reference's ``main.py`` / ``interfaces/*.py`` UNCHANGED on the tpgsr_b200 CUDA path.

``interfaces/base.py:19-22,30`` and ``interfaces/super_resolution.py:10,24-29`` import ``model``, ``loss`` and ``utils`` as top-level
packages.  Putting THIS directory on ``sys.path`` in front of the reference checkout,

    PYTHONPATH=<repo>/tpgsr_b200/shadow:<reference checkout>  python <reference checkout>/main.py --arch tsrn_tl_cascade ...

makes those imports resolve here first: the hot-path modules (``model.tsrn``, ``model.crnn``, ``model.srcnn``, ``model.srresnet``,
``model.stn_head``, ``model.tps_spatial_transformer``, ``loss.image_loss``, ``loss.semantic_loss``, ``utils.ssim_psnr``, ``utils.metrics``,
``utils.util``) are the drop-ins of ``tpgsr_b200``; every other submodule (``model.bicubic``, ``model.vdsr``, ``model.recognizer``,
``model.moran``, ``loss.gradient_loss``, ``utils.labelmaps``, ...) is found in the reference's own directory, which each shadow package
appends to its ``__path__`` — those files run as they are.  ``ptflops`` is shadowed as well: the reference calls it on the CPU module
(``interfaces/base.py:372-381``), and the drop-ins have no CPU forward.

The reference checkout is located through ``$TPGSR_REFERENCE_ROOT``, else the first ``sys.path`` entry (or the working directory)
that contains ``interfaces/super_resolution.py``.  Nothing here reads the reference at import time of ``tpgsr_b200`` itself.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from typing import List, Optional

HERE = os.path.dirname(os.path.abspath(__file__))


def path() -> str:
    """The directory to put in front of the reference checkout on ``sys.path`` / ``PYTHONPATH``."""
    return HERE


def reference_root() -> Optional[str]:
    cands = [os.environ.get('TPGSR_REFERENCE_ROOT')] + list(sys.path) + [os.getcwd()]
    for c in cands:
        if not c:
            continue
        c = os.path.abspath(c)
        if c != HERE and os.path.isfile(os.path.join(c, 'interfaces', 'super_resolution.py')) and os.path.isdir(os.path.join(c, 'model')):
            return c
    return None


def extend_with_reference(pkg_path: List[str], *rel: str) -> Optional[str]:
    """Append ``<reference>/<rel...>`` to a shadow package's ``__path__`` so that submodules that are not shadowed load from there."""
    root = reference_root()
    if root is None:
        return None
    d = os.path.join(root, *rel)
    if os.path.isdir(d) and d not in pkg_path:
        pkg_path.append(d)
    return d


def load_reference_module(alias: str, *rel: str):
    """Load ``<reference>/<rel...>`` under the module name ``alias`` (for names a shadow module passes through); None if unavailable."""
    root = reference_root()
    if root is None:
        return None
    fp = os.path.join(root, *rel)
    if not os.path.isfile(fp):
        return None
    if alias in sys.modules:
        return sys.modules[alias]
    spec = importlib.util.spec_from_file_location(alias, fp)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        sys.modules.pop(alias, None)
        raise
    return mod


def out_of_scope(module: str):
    """Module-level ``__getattr__`` for shadow modules: a clear error for reference names outside the hot path."""
    def __getattr__(name):
        raise AttributeError(f'{module}.{name} is outside the tpgsr_b200 hot path (SURVEY.md §8 marks it out of scope); only the classes '
                             f'the BASELINE.json configs use are shadowed')
    return __getattr__
