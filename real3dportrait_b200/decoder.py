"""OSGDecoder / FullyConnectedLayer — host mirror of modules/img2plane/triplane.py:122-146 (duplicate in
modules/eg3ds/models/triplane.py:166-189) and modules/eg3ds/models/networks_stylegan2.py:99-131.
Parameter names/shapes equal the reference's so released checkpoints load with strict=True."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi as capi


class FullyConnectedLayer(torch.nn.Module):
    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.activation = activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,in] -> [B,out] = x @ (W*gain)^T + b   (linear activation only on this path)."""
        if self.activation != 'linear' or self.bias is None or self.bias_gain != 1:
            raise NotImplementedError('only linear FullyConnectedLayer with bias and lr_multiplier=1 is on the render/SR path')
        x = capi.f32(x)
        assert x.ndim == 2 and x.shape[1] == self.in_features
        y = torch.empty(x.shape[0], self.out_features, device=x.device, dtype=torch.float32)
        capi.check(capi.lib().r3dp_sr_styles(capi.ptr(x), capi.ptr(capi.f32(self.weight)), capi.ptr(capi.f32(self.bias)), x.shape[0],
                                             self.in_features, self.out_features, C.c_float(1.0), capi.ptr(y), capi.stream()))
        return y

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


class OSGDecoder(torch.nn.Module):
    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.net = torch.nn.Sequential(
            FullyConnectedLayer(n_features, self.hidden_dim, lr_multiplier=options['decoder_lr_mul']),
            torch.nn.Softplus(),
            FullyConnectedLayer(self.hidden_dim, 1 + options['decoder_output_dim'], lr_multiplier=options['decoder_lr_mul']),
        )

    def mlp_struct(self) -> capi.MlpStruct:
        """C-ABI view of the four parameter tensors (kept alive by the module)."""
        l0, l2 = self.net[0], self.net[2]
        if l0.bias_gain != 1 or l2.bias_gain != 1:
            raise NotImplementedError('decoder_lr_mul != 1 is not on the inference path')
        self._keep = [capi.f32(l0.weight), capi.f32(l0.bias), capi.f32(l2.weight), capi.f32(l2.bias)]
        return capi.mlp_struct(*self._keep)

    def forward(self, sampled_features: torch.Tensor, ray_directions=None, **kwargs):
        """sampled_features [N,3,P,C] (mean over planes taken here) or [N,P,C] -> {'rgb': [N,P,32], 'sigma': [N,P,1]}."""
        f = capi.f32(sampled_features)
        if f.ndim == 3:
            f = f.unsqueeze(1)
        N, K, P, Cf = f.shape
        rgb = torch.empty(N, P, self.net[2].out_features - 1, device=f.device, dtype=torch.float32)
        sigma = torch.empty(N, P, 1, device=f.device, dtype=torch.float32)
        m = self.mlp_struct()
        capi.check(capi.lib().r3dp_decode(capi.ptr(f), N, K, P, Cf, C.byref(m), capi.ptr(rgb), capi.ptr(sigma), capi.stream()))
        return {'rgb': rgb, 'sigma': sigma}
