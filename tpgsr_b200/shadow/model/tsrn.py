"""``model.tsrn`` (reference ``model/tsrn.py:18-215,373-508``) -> the sm_100a drop-ins."""
from tpgsr_b200.model.tsrn import (TSRN, TSRN_TL, GruBlock, InfoGen, RecurrentResidualBlock, RecurrentResidualBlockTL,  # noqa: F401
                                   UpsampleBLock, mish)
from tpgsr_b200.shadow import out_of_scope

__getattr__ = out_of_scope('model.tsrn')
