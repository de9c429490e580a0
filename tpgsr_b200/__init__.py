"""tpgsr_b200 — B200-native (sm_100a) implementation of the TPGSR hot path behind the reference's nn.Module API.

Importing this package loads ``libtpgsr_b200.so`` (built by ``python tpgsr_b200/build.py``); there is no CPU or
PyTorch fallback.  Public surface (mirrors the reference's import paths):

    tpgsr_b200.model.tsrn.TSRN / TSRN_TL         model/tsrn.py
    tpgsr_b200.model.crnn.CRNN                   model/crnn/crnn.py
    tpgsr_b200.loss.image_loss.ImageLoss         loss/image_loss.py
    tpgsr_b200.loss.semantic_loss.SemanticLoss   loss/semantic_loss.py
    tpgsr_b200.train.TrainStep                   the step glue of interfaces/super_resolution.py:297-424 as one CUDA graph
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)

__all__ = ['_lib']
