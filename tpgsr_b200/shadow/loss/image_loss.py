"""``loss.image_loss`` (reference ``loss/image_loss.py:10-51``) -> the fused CUDA losses."""
from tpgsr_b200.loss.image_loss import GradientPriorLoss, ImageLoss  # noqa: F401
