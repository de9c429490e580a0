"""``model.srresnet`` (reference ``model/srresnet.py:88-234``) -> the sm_100a drop-in."""
from tpgsr_b200.model.srresnet import SRResNet_TL  # noqa: F401
from tpgsr_b200.shadow import out_of_scope

__getattr__ = out_of_scope('model.srresnet')
