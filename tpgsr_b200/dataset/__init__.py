from .collate import encode_labels, label_onehot_collate  # noqa: F401
from .dataset import resizeNormalize, stage  # noqa: F401
from .loader import BatchPrefetcher, batches  # noqa: F401
