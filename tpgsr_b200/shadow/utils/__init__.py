"""Shadow of the reference's ``utils`` package (``utils/__init__.py:3-4`` exports ``labelmaps.*`` and ``str_filt``)."""
from tpgsr_b200.shadow import extend_with_reference

extend_with_reference(__path__, 'utils')
from .util import str_filt  # noqa: E402,F401

try:                                     # utils/labelmaps.py is host-side vocabulary code: the reference's own file
    from .labelmaps import *  # noqa: E402,F401,F403
except ImportError:
    pass
