"""``model.stn_head`` (reference ``model/stn_head.py:25-106``) -> the parameter container executed inside TSRN / TSRN_TL."""
from tpgsr_b200.model.stn_head import STNHead  # noqa: F401
