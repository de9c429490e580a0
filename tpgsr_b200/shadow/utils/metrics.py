"""``utils.metrics``: ``get_string_crnn`` (reference ``utils/metrics.py:71-88``) is the CUDA CTC decode; ``get_string_aster``, ``Accuracy``
and the lexicon helpers pass through to the reference's file."""
from tpgsr_b200.shadow import load_reference_module
from tpgsr_b200.utils.metrics import get_string_crnn  # noqa: F401


def __getattr__(name):
    ref = load_reference_module('utils._reference_metrics', 'utils', 'metrics.py')
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError(f'utils.metrics.{name}: not part of the tpgsr_b200 hot path and no reference checkout to take it from')
