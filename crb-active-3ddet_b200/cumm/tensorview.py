"""cumm.tensorview shim: the reference only uses `tv.from_numpy(arr)` and `.numpy()` on the results of
Point2VoxelCPU3d.point_to_voxel (pcdet/datasets/processor/data_processor.py:10,54-59)."""
import numpy as np


class Tensor(object):
    def __init__(self, array):
        self._array = np.asarray(array)

    def numpy(self):
        return self._array

    @property
    def shape(self):
        return self._array.shape

    @property
    def dtype(self):
        return self._array.dtype


def from_numpy(array):
    return Tensor(array)
