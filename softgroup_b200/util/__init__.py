from .rle import rle_decode, rle_encode, rle_encode_ids, rle_encode_many  # noqa: F401
from .utils import cuda_cast, force_fp32  # noqa: F401
