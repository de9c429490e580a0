"""``model.srcnn`` (reference ``model/srcnn.py:109-145``) -> the sm_100a drop-in."""
from tpgsr_b200.model.srcnn import SRCNN  # noqa: F401
from tpgsr_b200.shadow import out_of_scope

__getattr__ = out_of_scope('model.srcnn')
