"""``model.tps_spatial_transformer`` (reference ``model/tps_spatial_transformer.py:54-112``) -> the buffer container executed inside TSRN."""
from tpgsr_b200.model.tps_spatial_transformer import TPSSpatialTransformer  # noqa: F401
