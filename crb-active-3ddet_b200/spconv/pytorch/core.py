"""SparseConvTensor: the container spconv.pytorch exposes (reference call sites:
pcdet/models/backbones_3d/spconv_backbone.py:141-146, pcdet/utils/spconv_utils.py:29-33,
pcdet/models/backbones_2d/map_to_bev/height_compression.py:21, pcdet/models/backbones_3d/pfe/voxel_set_abstraction.py:385-386).
"""
import torch

from crb3d import ops


class IndiceData(object):
    """Cached rulebook of one `indice_key`: B200 layout = output-stationary neighbour tables (see csrc/rulebook.cu)."""

    def __init__(self, out_indices, indices, nbr, nbr_t, out_spatial_shape, spatial_shape, ksize, stride, padding,
                 dilation, is_subm):
        self.out_indices = out_indices
        self.indices = indices
        self.nbr = nbr            # [K, n_out]  input row per (offset, output row) or -1
        self.nbr_t = nbr_t        # [K, n_in]   output row per (offset, input row) or -1 (None for SubM: nbr flipped)
        self.out_spatial_shape = out_spatial_shape
        self.spatial_shape = spatial_shape
        self.ksize = ksize
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.is_subm = is_subm
        self._pairs = None

    @property
    def indice_pairs(self):
        """spconv-format (indice_pairs [2,K,N], indice_pair_num [K]) derived on demand."""
        if self._pairs is None:
            self._pairs = ops.compact_pairs(self.nbr)
        return self._pairs


class SparseConvTensor(object):
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None,
                 benchmark=False, permanent_thrust_allocator=False, enable_timer=False, force_algo=None):
        self._features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self.grid = grid
        self.voxel_num = voxel_num
        self.benchmark = benchmark
        self.benchmark_record = {}

    @property
    def features(self):
        return self._features

    @features.setter
    def features(self, val):
        # spconv 2.x forbids assignment; pcdet's replace_feature helper falls back to it only for spconv 1.x
        self._features = val

    def replace_feature(self, feature):
        new = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid, self.voxel_num,
                               self.indice_dict, self.benchmark)
        new.benchmark_record = self.benchmark_record
        return new

    @property
    def spatial_size(self):
        n = 1
        for s in self.spatial_shape:
            n *= s
        return n

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key, None)

    def dense(self, channels_first=True):
        """(B, C, D, H, W) zero-filled dense tensor (channels_first=False: (B, D, H, W, C))."""
        out = _DenseFunction.apply(self._features, self.indices, self.batch_size, tuple(self.spatial_shape))
        return out if channels_first else out.permute(0, 2, 3, 4, 1).contiguous()

    def dense_bev_channels_last(self):
        """(B, C*D, H, W) view whose memory is NHWC - equal to dense().view(B, C*D, H, W) (height_compression.py:21-23)."""
        d = ops.sparse_to_dense(self._features, self.indices, self.batch_size, self.spatial_shape, channels_last_bev=True)
        return d.permute(0, 3, 1, 2)

    @property
    def sparity(self):
        return self.indices.shape[0] / max(self.spatial_size * self.batch_size, 1)


class _DenseFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, indices, batch_size, spatial_shape):
        ctx.save_for_backward(indices)
        ctx.meta = (features.shape[1], spatial_shape)
        return ops.sparse_to_dense(features, indices, batch_size, spatial_shape)

    @staticmethod
    def backward(ctx, grad):
        (indices,) = ctx.saved_tensors
        C, spatial_shape = ctx.meta
        return ops.dense_to_sparse(grad.contiguous(), indices, C, spatial_shape), None, None, None
