"""oracle/spconv_cpu.py -- a CPU stand-in for the `spconv.pytorch` surface the reference model touches, built on
oracle/spconv_oracle.py. TEST INFRASTRUCTURE ONLY (same import rule as the rest of oracle/): it exists so that the
UNMODIFIED reference model code (softgroup/model/blocks.py, softgroup.py) can be executed in the build container,
on the CPU, to generate whole-model golden vectors (tests/golden/make_forward_golden.py).

Parity note: the arithmetic is the spconv restatement of spconv_oracle.py (parity unpinned by the reference, see
there); what THIS module adds is only the module plumbing the reference uses -- SparseConvTensor with
`.features .indices .spatial_shape .batch_size .indice_dict .grid .replace_feature()`, SparseSequential applying
plain nn.Modules to `.features`, SubMConv3d / SparseConv3d / SparseInverseConv3d with `[out,k,k,k,in]` weights and
`indice_key` pairing (blocks.py:33-41, 50-70, 96-129; softgroup.py:60-65, 307, 388, 516, 720)."""
import sys
import types
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import spconv_oracle as so


class SparseConvTensor(object):

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, indice_dict=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.grid = grid
        self.indice_dict = indice_dict if indice_dict is not None else {}

    def replace_feature(self, feature):
        out = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid, self.indice_dict)
        return out


class SparseModule(nn.Module):
    pass


class SparseSequential(SparseModule):

    def __init__(self, *args):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)

    def forward(self, input):
        for module in self._modules.values():
            if isinstance(module, SparseModule):
                input = module(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input = input.replace_feature(module(input.features))
            else:
                input = module(input)
        return input


def _np_idx(t):
    return t.detach().cpu().numpy().astype(np.int32)


class _Conv(SparseModule):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, indice_key=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = [kernel_size] * 3 if isinstance(kernel_size, int) else list(kernel_size)
        self.stride = [stride] * 3 if isinstance(stride, int) else list(stride)
        self.padding = [padding] * 3 if isinstance(padding, int) else list(padding)
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        nn.init.kaiming_uniform_(self.weight.view(out_channels, -1), a=5**0.5)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def _apply(self, input, mp, indices, spatial_shape):
        w = self.weight.detach().numpy()
        out = so.conv_from_map(input.features.detach().numpy(), mp, w, acc64=True)
        out = torch.from_numpy(out)
        if self.bias is not None:
            out = out + self.bias.detach()
        res = SparseConvTensor(out, indices, spatial_shape, input.batch_size, input.grid, input.indice_dict)
        return res


class SubMConv3d(_Conv):

    def forward(self, input):
        assert self.kernel_size == [3, 3, 3] and self.padding == [1, 1, 1] and self.stride == [1, 1, 1]
        key = ('subm', self.indice_key)
        cached = input.indice_dict.get(key) if self.indice_key is not None else None
        if cached is None:
            cached = so.subm_map(_np_idx(input.indices))
            if self.indice_key is not None:
                input.indice_dict[key] = cached
        return self._apply(input, cached, input.indices, input.spatial_shape)


class SparseConv3d(_Conv):

    def forward(self, input):
        assert self.kernel_size == [2, 2, 2] and self.stride == [2, 2, 2] and self.padding == [0, 0, 0]
        out_indices, mp, inv, out_shape = so.down_map(_np_idx(input.indices), input.spatial_shape)
        if self.indice_key is not None:
            input.indice_dict[('down', self.indice_key)] = (inv, input.indices, input.spatial_shape)
        return self._apply(input, mp, torch.from_numpy(out_indices), out_shape)


class SparseInverseConv3d(_Conv):

    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, indice_key=indice_key)

    def forward(self, input):
        inv, indices, spatial_shape = input.indice_dict[('down', self.indice_key)]
        return self._apply(input, inv, indices, spatial_shape)


def install(as_names=('spconv', 'spconv.pytorch', 'spconv.pytorch.modules')):
    """Register this module under the names the reference imports (blocks.py:3-5, softgroup.py:5)."""
    me = sys.modules[__name__]
    root = types.ModuleType('spconv')
    root.pytorch = me
    me.modules = types.ModuleType('spconv.pytorch.modules')
    me.modules.SparseModule = SparseModule
    sys.modules['spconv'] = root
    sys.modules['spconv.pytorch'] = me
    sys.modules['spconv.pytorch.modules'] = me.modules
    return me
