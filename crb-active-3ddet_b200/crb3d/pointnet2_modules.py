"""Host-side mirror of the reference's stacked PointNet++ interface on the crb3d kernels (same names, argument meaning
and autograd behaviour as pcdet/ops/pointnet2/pointnet2_stack/pointnet2_utils.py:8-260 and pointnet2_modules.py:30-112),
i.e. what PV-RCNN's VoxelSetAbstraction (voxel_set_abstraction.py:334-411) and RoI-grid pooling (pvrcnn_head.py:68-114)
call. The MLPs are torch.nn (1x1 convs = GEMMs, as in the reference); ball query / grouping / FPS / 3-NN are the kernels.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt):
        """xyz (N1+N2.., 3), new_xyz (M1+M2.., 3) -> idx (M, nsample) int32 (first hits in index order, padded with the
        first hit), empty_ball_mask (M,) bool (rows without any hit; their idx is zeroed)."""
        assert new_xyz.is_contiguous() and new_xyz_batch_cnt.is_contiguous() and xyz.is_contiguous() and xyz_batch_cnt.is_contiguous()
        B, M = xyz_batch_cnt.shape[0], new_xyz.shape[0]
        idx = torch.zeros((M, nsample), dtype=torch.int32, device=xyz.device)
        ops.ball_query(B, M, radius, nsample, new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, idx)
        empty = idx[:, 0] == -1
        idx[empty] = 0
        ctx.mark_non_differentiable(idx, empty)
        return idx, empty

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None, None, None


ball_query = BallQuery.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, features_batch_cnt, idx, idx_batch_cnt):
        """features (N1+N2.., C), idx (M, nsample) -> (M, C, nsample)."""
        assert features.is_contiguous() and idx.is_contiguous()
        M, ns = idx.shape
        N, C = features.shape
        B = idx_batch_cnt.shape[0]
        out = torch.empty((M, C, ns), dtype=torch.float32, device=features.device)
        ops.group_points(B, M, C, ns, features, features_batch_cnt, idx, idx_batch_cnt, out)
        ctx.for_backwards = (B, N, idx, features_batch_cnt, idx_batch_cnt)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        B, N, idx, features_batch_cnt, idx_batch_cnt = ctx.for_backwards
        M, C, ns = grad_out.shape
        grad = torch.zeros((N, C), dtype=torch.float32, device=grad_out.device)
        ops.group_points_grad(B, M, C, N, ns, grad_out.contiguous(), idx, idx_batch_cnt, features_batch_cnt, grad)
        return grad, None, None, None


grouping_operation = GroupingOperation.apply


class QueryAndGroup(nn.Module):
    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt, features=None):
        """Returns new_features (M, 3 + C, nsample) (xyz offsets to the query first) and idx."""
        assert xyz.shape[0] == int(xyz_batch_cnt.sum()) and new_xyz.shape[0] == int(new_xyz_batch_cnt.sum())
        idx, empty = ball_query(self.radius, self.nsample, xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt)
        grouped_xyz = grouping_operation(xyz, xyz_batch_cnt, idx, new_xyz_batch_cnt)
        grouped_xyz = grouped_xyz - new_xyz.unsqueeze(-1)
        grouped_xyz[empty] = 0
        if features is not None:
            grouped = grouping_operation(features, xyz_batch_cnt, idx, new_xyz_batch_cnt)
            grouped[empty] = 0
            new_features = torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
        else:
            assert self.use_xyz, "cannot have no features and not use xyz"
            new_features = grouped_xyz
        return new_features, idx


class FarthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        """xyz (B, N, 3) -> (B, npoint) int32."""
        assert xyz.is_contiguous()
        B, N, _ = xyz.shape
        out = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        ops.farthest_point_sampling(B, N, npoint, xyz, temp, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


farthest_point_sample = furthest_point_sample = FarthestPointSampling.apply


class StackFarthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, xyz_batch_cnt, npoint):
        """xyz (N1+N2.., 3); npoint int or per-batch list -> global indices (sum npoint,) int32."""
        assert xyz.is_contiguous() and xyz.shape[1] == 3
        B = xyz_batch_cnt.numel()
        if not isinstance(npoint, torch.Tensor):
            npoint = [npoint] * B if not isinstance(npoint, (list, tuple)) else list(npoint)
            npoint = torch.tensor(npoint, device=xyz.device).int()
        temp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)
        out = torch.empty((int(npoint.sum()),), dtype=torch.int32, device=xyz.device)
        ops.stack_farthest_point_sampling(xyz, temp, xyz_batch_cnt.int().contiguous(), out, npoint.int().contiguous())
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


stack_farthest_point_sample = StackFarthestPointSampling.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, unknown_batch_cnt, known, known_batch_cnt):
        """-> (dist (N,3) euclidean, idx (N,3) int32 global indices into `known`)."""
        assert unknown.is_contiguous() and known.is_contiguous()
        N = unknown.shape[0]
        dist2 = torch.empty((N, 3), dtype=torch.float32, device=unknown.device)
        idx = torch.empty((N, 3), dtype=torch.int32, device=unknown.device)
        ops.three_nn(unknown_batch_cnt.numel(), N, known.shape[0], unknown, unknown_batch_cnt.int().contiguous(), known,
                     known_batch_cnt.int().contiguous(), dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        """features (M, C), idx/weight (N, 3) -> (N, C)."""
        assert idx.shape == weight.shape and idx.shape[1] == 3
        ctx.for_backwards = (idx, weight, features.shape[0])
        out = torch.empty((idx.shape[0], features.shape[1]), dtype=torch.float32, device=features.device)
        ops.three_interpolate(idx.shape[0], features.shape[1], features.contiguous(), idx.contiguous(), weight.contiguous(), out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, M = ctx.for_backwards
        grad = torch.zeros((M, grad_out.shape[1]), dtype=torch.float32, device=grad_out.device)
        ops.three_interpolate_grad(grad_out.shape[0], grad_out.shape[1], grad_out.contiguous(), idx.contiguous(),
                                   weight.contiguous(), grad)
        return grad, None, None


three_interpolate = ThreeInterpolate.apply


class StackSAModuleMSG(nn.Module):
    """Multi-scale set abstraction on stacked batches (pointnet2_modules.py:30-112): per radius, ball query + grouping
    -> shared MLP (1x1 Conv2d + BN + ReLU) -> max (or avg) pool over the samples; scales concatenated on channels."""

    def __init__(self, radii, nsamples, mlps, use_xyz=True, pool_method="max_pool"):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.groupers, self.mlps = nn.ModuleList(), nn.ModuleList()
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            self.groupers.append(QueryAndGroup(radius, nsample, use_xyz=use_xyz))
            spec = list(spec)
            if use_xyz:
                spec[0] += 3
            layers = []
            for k in range(len(spec) - 1):
                layers += [nn.Conv2d(spec[k], spec[k + 1], kernel_size=1, bias=False), nn.BatchNorm2d(spec[k + 1]), nn.ReLU()]
            self.mlps.append(nn.Sequential(*layers))
        self.pool_method = pool_method
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            if isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0)

    def forward(self, xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt, features=None, empty_voxel_set_zeros=True):
        outs = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            new_features, _ = grouper(xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt, features)   # (M, C, ns)
            x = mlp(new_features.permute(1, 0, 2).unsqueeze(0))                                     # (1, C', M, ns)
            x = x.max(dim=3)[0] if self.pool_method == "max_pool" else x.mean(dim=3)
            outs.append(x.squeeze(0).permute(1, 0))                                                 # (M, C')
        return new_xyz, torch.cat(outs, dim=1)
