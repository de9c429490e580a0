"""Shadow of the reference's ``model`` package (``interfaces/base.py:19-22``): hot-path submodules here, the rest from the reference."""
from tpgsr_b200.shadow import extend_with_reference

extend_with_reference(__path__, 'model')
