from .blocks import MLP, Custom1x1Subm3d, ResidualBlock, UBlock  # noqa: F401
from .softgroup import SoftGroup  # noqa: F401
