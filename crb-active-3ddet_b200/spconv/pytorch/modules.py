"""SparseModule / SparseSequential (spconv.pytorch.modules): spconv_backbone.py:21-25,30,77-117 wraps each conv with
nn.BatchNorm1d + nn.ReLU inside a SparseSequential; dense modules run on the (N, C) feature matrix."""
from collections import OrderedDict

import torch
from torch import nn


def is_spconv_module(module):
    return isinstance(module, SparseModule)


def is_sparse_conv(module):
    from .conv import SparseConvolution
    return isinstance(module, SparseConvolution)


class SparseModule(nn.Module):
    """Marker base class: forward() takes and returns a SparseConvTensor."""
    pass


class SparseSequential(SparseModule):
    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists")
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        it = iter(self._modules.values())
        for _ in range(idx):
            next(it)
        return next(it)

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    @staticmethod
    def _bn_affine(bn):
        """eval-mode BatchNorm1d as y = x*scale + shift. Cached on the module (5 tiny kernels otherwise, per layer per
        call); the cache key is the version counter of every tensor involved, so load_state_dict / optimizer steps /
        running-stat updates invalidate it."""
        key = (bn.running_mean._version, bn.running_var._version, bn.weight._version if bn.affine else 0,
               bn.bias._version if bn.affine else 0, bn.running_mean.data_ptr())
        cached = getattr(bn, "_crb3d_affine", None)
        if cached is not None and cached[0] == key:
            return cached[1], cached[2]
        with torch.no_grad():
            inv = 1.0 / torch.sqrt(bn.running_var + bn.eps)
            scale = bn.weight * inv if bn.affine else inv
            shift = (bn.bias if bn.affine else 0.0) - bn.running_mean * scale
            scale, shift = scale.float().contiguous(), shift.float().contiguous()
        bn._crb3d_affine = (key, scale, shift)
        return scale, shift

    def forward(self, input):
        from .core import SparseConvTensor
        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            module = mods[i]
            if is_spconv_module(module):
                # conv -> BatchNorm1d(eval) -> ReLU is folded into the conv kernel's epilogue when no graph is recorded
                if (is_sparse_conv(module) and isinstance(input, SparseConvTensor) and not torch.is_grad_enabled()
                        and i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d) and not mods[i + 1].training
                        and mods[i + 1].track_running_stats):
                    scale, shift = self._bn_affine(mods[i + 1])
                    relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                    input = module(input, fused_scale=scale, fused_shift=shift, fused_relu=relu)
                    i += 3 if relu else 2
                    continue
                input = module(input)
            else:
                if isinstance(input, SparseConvTensor):
                    if input.indices.shape[0] != 0:
                        input = input.replace_feature(module(input.features))
                else:
                    input = module(input)
            i += 1
        return input


class ToDense(SparseModule):
    def forward(self, x):
        return x.dense()


class RemoveGrid(SparseModule):
    def forward(self, x):
        x.grid = None
        return x
