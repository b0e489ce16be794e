from .core import (SparseConv3d, SparseConvTensor, SparseInverseConv3d, SparseModule,  # noqa: F401
                   SparseSequential, SubMConv3d, bn_relu_rows, build_down_map, build_subm_map, check_overflow, conv_forward,
                   fold_bn)
