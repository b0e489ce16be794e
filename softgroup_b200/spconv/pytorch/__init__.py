"""`import spconv.pytorch as spconv` compatible namespace (see softgroup_b200.install_as_reference_backends)."""
from ..core import (SparseConv3d, SparseConvTensor, SparseInverseConv3d, SparseModule,  # noqa: F401
                    SparseSequential, SubMConv3d)
from . import modules  # noqa: F401
