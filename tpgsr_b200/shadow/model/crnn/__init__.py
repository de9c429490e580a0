"""``model.crnn`` (reference ``model/crnn/__init__.py``: ``CRNN, CRNN_ResNet18, Model``).  ``CRNN`` (``model/crnn/crnn.py:29-90``) is the
sm_100a drop-in; ``CRNN_ResNet18`` and the OPT ``Model`` come from the reference's files when a checkout is present."""
from tpgsr_b200.model.crnn import CRNN, BidirectionalLSTM  # noqa: F401
from tpgsr_b200.shadow import extend_with_reference

extend_with_reference(__path__, 'model', 'crnn')


def __getattr__(name):
    if name in ('CRNN_ResNet18', 'Model'):
        import importlib
        try:
            mod = importlib.import_module('.crnn' if name == 'CRNN_ResNet18' else '.model', __name__)
        except ImportError as e:
            raise AttributeError(f'model.crnn.{name} needs the reference checkout (out of the tpgsr_b200 hot path): {e}') from e
        return getattr(mod, name)
    raise AttributeError(f'module {__name__!r} has no attribute {name!r}')
