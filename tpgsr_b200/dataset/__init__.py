from .dataset import resizeNormalize, stage  # noqa: F401
from .loader import BatchPrefetcher, batches  # noqa: F401
