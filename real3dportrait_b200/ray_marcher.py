"""MipRayMarcher2 — host mirror of modules/eg3ds/volumetric_rendering/ray_marcher.py:20-63."""
from __future__ import annotations

import torch

from . import _capi as capi


class MipRayMarcher2(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def run_forward(self, colors, densities, depths, rendering_options):
        """colors [N,M,S,C], densities [N,M,S,1], depths [N,M,S,1] -> rgb [N,M,C], depth [N,M,1], weights [N,M,S-1,1]."""
        assert rendering_options['clamp_mode'] == 'softplus', 'MipRayMarcher only supports `clamp_mode`=`softplus`!'
        colors, densities, depths = capi.f32(colors), capi.f32(densities), capi.f32(depths)
        N, M, S, Cc = colors.shape
        dev = colors.device
        rgb = torch.empty(N, M, Cc, device=dev)
        depth = torch.empty(N, M, 1, device=dev)
        weights = torch.empty(N, M, S - 1, 1, device=dev)
        ws = torch.empty(8, device=dev, dtype=torch.int32)
        capi.check(capi.lib().r3dp_ray_march(capi.ptr(colors), capi.ptr(densities), capi.ptr(depths), N, M, S, Cc,
                                             int(bool(rendering_options.get('white_back', False))), capi.ptr(rgb), capi.ptr(depth),
                                             capi.ptr(weights), capi.ptr(ws, torch.int32), capi.stream()))
        return rgb, depth, weights

    def forward(self, colors, densities, depths, rendering_options):
        return self.run_forward(colors, densities, depths, rendering_options)
