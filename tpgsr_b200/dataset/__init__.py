from .dataset import resizeNormalize  # noqa: F401
