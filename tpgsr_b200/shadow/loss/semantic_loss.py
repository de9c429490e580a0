"""``loss.semantic_loss`` (reference ``loss/semantic_loss.py:10-38``) -> the fused CUDA loss."""
from tpgsr_b200.loss.semantic_loss import SemanticLoss  # noqa: F401
