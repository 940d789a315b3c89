"""spconv.pytorch drop-in (subset used by the reference's SECOND / PV-RCNN / PartA2 model code)."""
from . import conv, modules  # noqa: F401
from .conv import SparseConv3d, SparseConvolution, SparseInverseConv3d, SubMConv3d  # noqa: F401
from .core import SparseConvTensor  # noqa: F401
from .modules import RemoveGrid, SparseModule, SparseSequential, ToDense  # noqa: F401
