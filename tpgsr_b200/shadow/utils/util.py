"""``utils.util``: ``str_filt`` (reference ``utils/util.py:12-24``) is ours; the other helpers pass through to the reference's file."""
from tpgsr_b200.shadow import load_reference_module
from tpgsr_b200.utils.util import str_filt  # noqa: F401


def __getattr__(name):
    ref = load_reference_module('utils._reference_util', 'utils', 'util.py')
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError(f'utils.util.{name}: not part of the tpgsr_b200 hot path and no reference checkout to take it from')
