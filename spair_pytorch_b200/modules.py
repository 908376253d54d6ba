"""Building blocks of the SPAIR drop-in, under the reference's ``spair.modules`` API.

Same public names, argument meaning and error behaviour as the reference's
``spair/modules.py`` so that ``from spair.modules import *`` keeps working:

  Backbone, compute_backbone_feature_shape, build_MLP, SequentialMultipleOutput,
  latent_to_mean_std, clamped_sigmoid, exponential_decay, stn, to_C_H_W, to_H_W_C, safe_log

The dense contractions (``Backbone`` convs, ``build_MLP`` linears) stay on cuDNN / cuBLAS.
``stn`` — the operator-level entry of the spatial transformer (reference modules.py:216-273) —
runs on the hand-written sm_100a kernels through ``spair_pytorch_b200.ops`` (forward and
backward, no CPU implementation).  The small scalar helpers are plain tensor expressions.
"""
from __future__ import annotations

import copy
from collections import OrderedDict

import numpy as np
import torch
from torch import nn
from torch.nn import Conv2d, Linear, Module, ModuleList, ReLU, Sequential

from . import config as cfg

__all__ = ["Backbone", "compute_backbone_feature_shape", "build_MLP", "SequentialMultipleOutput",
           "latent_to_mean_std", "clamped_sigmoid", "exponential_decay", "stn", "to_C_H_W", "to_H_W_C",
           "safe_log", "receptive_field_geometry", "cfg", "OrderedDict", "np", "torch", "nn", "Sequential", "Conv2d",
           "ReLU", "Linear", "Module", "ModuleList"]


def receptive_field_geometry(topology, image_hw):
    """Receptive-field arithmetic of the conv stack (reference modules.py:68-105).

    Returns (padding (left, right, top, bottom), n_grid_cells [Hc, Wc], grid_cell_size [py, px]) such
    that output cell (h, w) is centred on input pixels [h*py, (h+1)*py) x [w*px, (w+1)*px).
    Known answer (reference test_notebook.ipynb:350): default topology on 128x128 ->
    rf 31, cell 12, grid 11x11, padding 9 before / 14 after."""
    stride_acc = np.array([1, 1])
    rf = np.array([1, 1])
    for layer in topology:
        rf = rf + (np.array(layer['kernel_size']) - 1) * stride_acc
        stride_acc = stride_acc * np.array(layer['stride'])
    cell = stride_acc
    before = np.floor(rf / 2 - cell / 2).astype('i')
    hw = np.array(image_hw)
    n_cells = np.ceil(hw / cell).astype('i')
    after = rf + (n_cells - 1) * cell - hw - before
    return (int(before[1]), int(after[1]), int(before[0]), int(after[0])), n_cells, cell


class Backbone(Module):
    """Feature extractor: ZeroPad2d + conv/ReLU stack + 1x1 output conv, no output activation
    (reference modules.py:12-111).  ``topology`` entries are dicts with ``filters`` (or
    ``out_channels``), ``kernel_size`` and ``stride``; unlike the reference the caller's list is
    never mutated."""

    def __init__(self, input_shape, n_out_channels, topology=None, internal_activation=ReLU):
        super().__init__()
        self.topology = copy.deepcopy(cfg.DEFAULT_BACKBONE_TOPOLOGY if topology is None else topology)
        self.input_shape = input_shape
        self.net = self._build_backbone(input_shape[0], n_out_channels, internal_activation)
        pad, self.n_grid_cells, self.grid_cell_size = receptive_field_geometry(self.topology, input_shape[-2:])
        self.padding = nn.ZeroPad2d(pad)

    def _build_backbone(self, n_in_channels, n_out_channels, activation):
        layers = OrderedDict()
        width = n_in_channels
        for i, spec in enumerate(self.topology):
            spec['in_channels'] = width
            if 'filters' in spec:
                spec['out_channels'] = spec.pop('filters')
            width = spec['out_channels']
            layers['conv_%d' % i] = Conv2d(**spec)
            layers['act_%d' % i] = activation()
        layers['conv_out'] = Conv2d(in_channels=width, out_channels=n_out_channels, kernel_size=1, stride=1)
        return Sequential(layers)

    def compute_output_shape(self):
        """[F, Hc, Wc] for one image of ``cfg.INPUT_IMAGE_SHAPE`` (reference modules.py:32-41; the
        random probe image is kept so the RNG stream of model construction matches the reference)."""
        probe = torch.rand(1, *cfg.INPUT_IMAGE_SHAPE)
        with torch.no_grad():
            return self(probe.to(next(self.parameters()).device)).shape[1:]

    def _fused_stem(self, x, channels_last=False):
        """ZeroPad2d + conv_0 + act_0 as one sm_100a kernel each way (csrc/stem.cu) when the layer has the shape the
        kernel covers and the image needs no gradient; None otherwise (library path).  ``channels_last``: the map is
        written [B,Ho,Wo,Cout], the layout the GEMM tail consumes."""
        conv = self.net[0]
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and not x.requires_grad
                and isinstance(conv, Conv2d) and isinstance(self.net[1], ReLU) and conv.bias is not None
                and conv.padding == (0, 0) and conv.dilation == (1, 1) and conv.groups == 1
                and conv.stride[0] == conv.stride[1] and conv.weight.dtype == torch.float32):
            return None
        from . import kernels as K, ops
        if not K.stem_supported(conv.in_channels, conv.out_channels, conv.kernel_size) or x.shape[1] != conv.in_channels:
            return None
        pl, pr, pt, pb = self.padding.padding
        stride, k = conv.stride[0], conv.kernel_size[0]
        Ho = (x.shape[2] + pt + pb - k) // stride + 1
        Wo = (x.shape[3] + pl + pr - k) // stride + 1
        if Ho <= 0 or Wo <= 0:
            return None
        return ops.StemConvFunction.apply(x, conv.weight, conv.bias, stride, pt, pl, Ho, Wo, channels_last)

    def _gemm_tail_plan(self, x):
        """(specs, params) of conv_1 .. conv_out for the tcgen05 GEMM path (ops.ConvTailFunction) when every layer has the
        shape it covers; None otherwise (cuDNN)."""
        from . import ops
        layers = list(self.net)[2:]
        if not (ops.USE_TENSOR_CORE_GEMM and x.is_cuda and x.dtype == torch.float32 and isinstance(self.net[0], Conv2d)):
            return None
        specs, params, i, cin = [], [], 0, self.net[0].out_channels
        while i < len(layers):
            conv = layers[i]
            if not (isinstance(conv, Conv2d) and conv.padding == (0, 0) and conv.dilation == (1, 1) and conv.groups == 1
                    and conv.kernel_size[0] == conv.kernel_size[1] and conv.stride[0] == conv.stride[1]
                    and conv.bias is not None and conv.in_channels == cin and cin % 4 == 0):
                return None
            relu = i + 1 < len(layers) and isinstance(layers[i + 1], ReLU)
            if i + 1 < len(layers) and not relu:
                return None
            specs.append((conv.kernel_size[0], conv.stride[0], relu))
            params += [conv.weight, conv.bias]
            cin = conv.out_channels
            i += 2 if relu else 1
        return (tuple(specs), params) if specs else None

    def forward(self, x):
        from . import ops
        tail = self._gemm_tail_plan(x)
        y = self._fused_stem(x, channels_last=tail is not None)
        if y is not None:
            if tail is not None:      # the stem wrote channels-last: conv_1 reads it as it is
                return ops.ConvTailFunction.apply(y, tail[0], True, *tail[1])
            for layer in list(self.net)[2:]:
                y = layer(y)
            return y
        return self.net(self.padding(x))


def compute_backbone_feature_shape(backbone):
    """Reference modules.py:113-122 (feeds an un-batched image through the backbone)."""
    return backbone(torch.randn(cfg.INPUT_IMAGE_SHAPE)).shape


class SequentialMultipleOutput(Module):
    """Shared body + several linear heads; ``forward`` returns a GENERATOR over the head outputs,
    as the reference does (modules.py:276-284)."""

    def __init__(self, input, outputs):
        super().__init__()
        self.body = Sequential(input)
        self.output_layers = ModuleList(list(outputs.values()))

    def forward(self, x):
        hidden = self.body(x)
        return (head(hidden) for head in self.output_layers)


def build_MLP(n_in, output=None, multiple_output=None, hidden_layers=None, activation=None,
              internal_activation=ReLU):
    """MLP builder (reference modules.py:124-165): ``dense{i}``/``relu{i}`` hidden layers, then
    either one ``out`` layer (+ optional ``act``) or ``SequentialMultipleOutput`` heads."""
    hidden_layers = cfg.DEFAULT_MLP_TOPOLOGY if hidden_layers is None else hidden_layers
    body = OrderedDict()
    width = n_in
    for i, h in enumerate(hidden_layers):
        body['dense%d' % i] = Linear(width, h)
        body['relu%d' % i] = internal_activation()
        width = h
    if output is not None:
        body['out'] = Linear(width, output)
        if activation is not None:
            body['act'] = activation()
        return Sequential(body)
    if multiple_output is not None:
        heads = OrderedDict(('out_%d' % i, Linear(width, n)) for i, n in enumerate(multiple_output))
        return SequentialMultipleOutput(body, heads)
    raise AssertionError('Unknown output type')


def latent_to_mean_std(latent_var):
    """Split a latent into (mean, std) with std = 2*sigmoid(clamp(log_std, -10, 10)) (modules.py:167-176)."""
    mean, log_std = torch.chunk(latent_var, 2, dim=-1)
    return mean, 2 * torch.sigmoid(log_std.clamp(-10, 10))


def clamped_sigmoid(logit, use_analytical=False):
    """sigmoid(clamp(logit, -10, 10)), or the un-clamped 1/(exp(-x)+1) (modules.py:178-189)."""
    if use_analytical:
        return 1 / ((-logit).exp() + 1)
    return torch.sigmoid(torch.clamp(logit, -10, 10))


def exponential_decay(global_step, device, start, end, decay_rate, decay_step, staircase=False, log_space=False):
    """(start-end) * decay_rate ** (step/decay_step) + end, optionally log(value + 1e-6)
    (modules.py:191-213); evaluated in fp32 like the reference."""
    step = torch.tensor(global_step, dtype=torch.float32).to(device)
    exponent = step // decay_step if staircase else step / decay_step
    value = (start - end) * (decay_rate ** exponent) + end
    return (value + 1e-6).log() if log_space else value


def stn(image, z_where, output_dims, device=None, inverse=False):
    """Spatial transformer (reference modules.py:216-273).

    ``z_where`` rows are (xt, yt, xs, ys), normalised to the image.  ``inverse=False`` cuts an
    ``output_dims`` glimpse out of ``image`` (bilinear, border padding); ``inverse=True`` pastes
    ``image`` back onto an ``output_dims`` canvas (zeros padding).  ``align_corners`` is False (the
    behaviour of the reference on torch >= 1.3).  Differentiable wrt ``z_where`` and ``image``.
    Runs on the sm_100a kernels; CUDA fp32 tensors only."""
    from . import ops
    z_where = z_where.reshape(-1, 4)
    if inverse:
        return ops.PasteFunction.apply(image, z_where, int(output_dims[0]), int(output_dims[1]))
    return ops.GlimpseFunction.apply(image, z_where, int(output_dims[0]), int(output_dims[1]))


def to_C_H_W(t: torch.Tensor):
    """[B, H, W, C] -> [B, C, H, W] (modules.py:286-289)."""
    assert t.shape[1] == t.shape[2] and t.shape[3] != t.shape[2], 'are you sure this tensor is in [B, H, W, C] format?'
    return t.permute(0, 3, 1, 2)


def to_H_W_C(t: torch.Tensor):
    """[B, C, H, W] -> [B, H, W, C] (modules.py:291-294)."""
    assert t.shape[2] == t.shape[3] and t.shape[1] != t.shape[2], 'are you sure this tensor is in [B, C, H, W] format?'
    return t.permute(0, 2, 3, 1)


def safe_log(t):
    """log(t + 1e-9) (modules.py:296-297)."""
    return torch.log(t + 1e-9)
