from .rle import rle_decode, rle_encode, rle_encode_ids  # noqa: F401
from .utils import cuda_cast, force_fp32  # noqa: F401
