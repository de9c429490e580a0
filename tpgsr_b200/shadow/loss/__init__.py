"""Shadow of the reference's ``loss`` package (``interfaces/base.py:30``): ``image_loss`` / ``semantic_loss`` here, the rest from the reference."""
from tpgsr_b200.shadow import extend_with_reference

extend_with_reference(__path__, 'loss')
