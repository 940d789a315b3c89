"""Python modules that stand in for the reference's pybind11 CUDA extensions, function for function
(pcdet/ops/*/src/*_api.cpp). `crb3d.dropin.install()` registers them under the reference's import names so the unmodified
pcdet/ops/*_utils.py wrappers (`from . import iou3d_nms_cuda` ...) resolve to these."""
