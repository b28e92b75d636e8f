"""RaySampler — host mirror of modules/eg3ds/volumetric_rendering/ray_sampler.py:18-63."""
from __future__ import annotations

import torch

from . import _capi as capi


class RaySampler(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, cam2world_matrix: torch.Tensor, intrinsics: torch.Tensor, resolution: int):
        """cam2world_matrix (N,4,4), intrinsics (N,3,3), resolution int -> ray_origins (N,M,3), ray_dirs (N,M,3);
        M = resolution**2, ray m = row*resolution + col (ray_sampler.py:43-44)."""
        assert cam2world_matrix.ndim == 3 and cam2world_matrix.shape[1:] == (4, 4), cam2world_matrix.shape
        assert intrinsics.ndim == 3 and intrinsics.shape[1:] == (3, 3), intrinsics.shape
        N, M = cam2world_matrix.shape[0], int(resolution) ** 2
        c2w, K = capi.f32(cam2world_matrix), capi.f32(intrinsics)
        ray_o = torch.empty(N, M, 3, device=c2w.device, dtype=torch.float32)
        ray_d = torch.empty_like(ray_o)
        capi.check(capi.lib().r3dp_gen_rays(capi.ptr(c2w), capi.ptr(K), N, int(resolution), capi.ptr(ray_o), capi.ptr(ray_d),
                                            capi.stream()))
        return ray_o, ray_d
