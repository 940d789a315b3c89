"""SparseConvolution modules (SubMConv3d / SparseConv3d / SparseInverseConv3d) over the crb3d C ABI.

Drop-in for spconv-cu113==2.1.21 as used by pcdet/models/backbones_3d/spconv_backbone.py:12-17,77-117 and
pcdet/utils/spconv_utils.py:11-25 (find_all_spconv_keys relies on `spconv.conv.SparseConvolution`).
Weight layout [C_out, kz, ky, kx, C_in] with kaiming_uniform(a=sqrt(5)) init, so reference checkpoints load unchanged
(pcdet/models/detectors/detector3d_template.py:455-484).
"""
import math

import torch
from torch import nn
from torch.nn import init

from crb3d import ops

from .core import IndiceData, SparseConvTensor
from .modules import SparseModule


def _triple(v, ndim=3):
    if isinstance(v, (list, tuple)):
        assert len(v) == ndim
        return [int(x) for x in v]
    return [int(v)] * ndim


class _SparseConvFunction(torch.autograd.Function):
    """out = conv(feat; nbr, W) (+bias). backward: dX through the transposed table, dW by the pair-wise outer products."""

    @staticmethod
    def forward(ctx, feat, weight, bias, nbr, nbr_t, subm):
        out = ops.spconv_forward(feat, nbr, weight, shift=bias)
        ctx.save_for_backward(feat, weight, nbr, nbr_t if nbr_t is not None else nbr)
        ctx.subm = subm
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        feat, weight, nbr, nbr_t = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        K = nbr.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if ctx.subm:
                # SubM: table^T[k] == table[K-1-k]; reuse the forward table with reversed weight slices
                kmap = torch.arange(K - 1, -1, -1, dtype=torch.int32, device=feat.device)
                dx = ops.spconv_forward(grad_out, nbr, weight, transpose=True, kmap=kmap)
            else:
                dx = ops.spconv_forward(grad_out, nbr_t, weight, transpose=True)
        if ctx.needs_input_grad[1]:
            dw = ops.spconv_wgrad(feat, grad_out, nbr, weight.shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = grad_out.sum(0)
        return dx, dw, db, None, None, None


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 algo=None, fp32_accum=None, name=None):
        super().__init__()
        if ndim != 3:
            raise NotImplementedError("crb3d spconv shim implements 3-D sparse convolutions only")
        if groups != 1 or transposed:
            raise NotImplementedError("groups != 1 / transposed sparse conv are not on the CRB hot path")
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _triple(kernel_size)
        self.conv1x1 = all(k == 1 for k in self.kernel_size)
        self.stride = _triple(stride)
        self.padding = _triple(padding)
        self.dilation = _triple(dilation)
        self.output_padding = _triple(output_padding)
        self.groups = groups
        self.subm = subm
        self.inverse = inverse
        self.transposed = transposed
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def extra_repr(self):
        return "{in_channels}, {out_channels}, kernel_size={kernel_size}, stride={stride}, padding={padding}, " \
               "subm={subm}, indice_key={indice_key}".format(**self.__dict__)

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels
            for k in self.kernel_size:
                fan_in *= k
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    # -- rulebook ---------------------------------------------------------------------------------------------
    def _rulebook(self, x):
        datas = x.find_indice_pair(self.indice_key)
        if self.inverse:
            if datas is None:
                raise ValueError("SparseInverseConv3d needs the rulebook of a previous conv with indice_key=%r" % self.indice_key)
            return datas
        if datas is not None and self.indice_key is not None:
            if self.subm and datas.is_subm and datas.indices.shape[0] == x.indices.shape[0]:
                return datas
            if not self.subm and not datas.is_subm:
                return datas
        if self.subm:
            # rows ranked by a strided rulebook of this forward pass: read the table off its cell -> row map (no hash table)
            cm = x.indice_dict.get("_crb3d_cellmaps", {}).get(x.indices.data_ptr())
            nbr = ops.subm_rulebook(x.indices, x.spatial_shape, self.kernel_size, self.dilation, cellmap=cm)
            datas = IndiceData(x.indices, x.indices, nbr, None, x.spatial_shape, x.spatial_shape, self.kernel_size,
                               [1, 1, 1], [(k // 2) * d for k, d in zip(self.kernel_size, self.dilation)],
                               self.dilation, True)
        else:
            oc, oshape, nbr, nbr_t, cm = ops.sparse_rulebook(x.indices, x.batch_size, x.spatial_shape, self.kernel_size,
                                                             self.stride, self.padding, self.dilation, want_cellmap=True)
            x.indice_dict.setdefault("_crb3d_cellmaps", {})[oc.data_ptr()] = cm
            datas = IndiceData(oc, x.indices, nbr, nbr_t, oshape, x.spatial_shape, self.kernel_size, self.stride,
                               self.padding, self.dilation, False)
        if self.indice_key is not None:
            x.indice_dict[self.indice_key] = datas
        return datas

    def forward(self, x, fused_scale=None, fused_shift=None, fused_relu=False):
        assert isinstance(x, SparseConvTensor)
        feat = x.features
        if feat.dtype != torch.float32:
            feat = feat.float()
        datas = self._rulebook(x)
        if self.inverse:
            nbr, nbr_t, out_idx, out_shape = datas.nbr_t, datas.nbr, datas.indices, datas.spatial_shape
            if nbr is None:
                raise ValueError("inverse conv over a submanifold rulebook is not defined")
        else:
            nbr, nbr_t, out_idx, out_shape = datas.nbr, datas.nbr_t, datas.out_indices, datas.out_spatial_shape
        if fused_scale is not None and not torch.is_grad_enabled():
            shift = fused_shift if self.bias is None else fused_shift + fused_scale * self.bias
            # inference with the TF32 tensor-core kernel: weights rounded (not truncated) to TF32 once, outputs rounded in the
            # epilogue - the same arithmetic as the captured static step (crb3d/second.py: _static_step)
            tc = ops.SPCONV_TF32
            out = ops.spconv_forward(feat, nbr, ops.tf32_weight(self) if tc else self.weight, scale=fused_scale, shift=shift,
                                     relu=fused_relu, round_out=tc)
        else:
            out = _SparseConvFunction.apply(feat, self.weight, self.bias, nbr, nbr_t, self.subm and not self.inverse)
        res = SparseConvTensor(out, out_idx, out_shape, x.batch_size, x.grid, x.voxel_num, x.indice_dict, x.benchmark)
        res.benchmark_record = x.benchmark_record
        return res


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, fp32_accum=None, name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, fp32_accum=None, name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key=indice_key)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True, algo=None, fp32_accum=None,
                 name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key)
