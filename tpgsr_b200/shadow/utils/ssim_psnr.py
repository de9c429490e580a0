"""``utils.ssim_psnr`` (reference ``utils/ssim_psnr.py:9-15,53-78``) -> the fused CUDA metrics."""
from tpgsr_b200.utils.ssim_psnr import SSIM, calculate_psnr  # noqa: F401
