from ..core import SparseModule, SparseSequential  # noqa: F401
