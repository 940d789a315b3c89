"""crb3d - B200-native (sm_100a) hot path of CRB-active-3Ddet: host-side Python over the C ABI in include/crb3d.h.

Add `<repo>/crb-active-3ddet_b200` to sys.path, then `import crb3d` (kernels), `import spconv.pytorch` /
`cumm.tensorview` (drop-in shims) or `crb3d.dropin.install()` (pcdet extension-module shims).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
