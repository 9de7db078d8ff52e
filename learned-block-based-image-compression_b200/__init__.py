"""B200-native closed-loop encode/decode of the block-based masked-convolution codec (v9).

Drop-in for the reference's compress/decompress path (graphs/models/BlockBasedImgCompLossy_net.py:
319-452 behind agents/blkbsdimgcomp_agent.py:560-641).  Host code is Python/PyTorch; all compute is
hand-written sm_100a CUDA behind the C ABI declared in include/lbic.h.  There is no CPU fallback.
"""
from . import weights  # noqa: F401
from .config import load_config  # noqa: F401
from .net import BlkBasedPostProcessing, BlockBasedImgCompLossyNetv9, get_lru, get_scale_table  # noqa: F401,E402
from .layout import arrange_block_pixels_to_channel_dim, arrange_channel_dim_to_block_pixels  # noqa: F401,E402
from .codec import ImageCodec, ms_ssim, pack_container, psnr, unpack_container  # noqa: F401,E402
