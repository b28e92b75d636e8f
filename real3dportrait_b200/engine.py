"""FrameEngine — the per-GPU frame loop of a clip (the caller side of the hot path, inference/real3d_infer.py:480-492):
a batch of tri-planes + cameras + jitter in, 512^2 frames out, one process per GPU; with world > 1 every step ends with
an NCCL all-gather of that step's frames (the only collective of the path: frames are independent)."""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from . import _capi as capi
from .synthesis import RenderHead


def default_sr_mode() -> str:
    try:
        from . import sr_tc  # noqa: F401
        return 'tc' if sr_tc.available() else 'fp32'
    except ImportError:
        return 'fp32'


class FrameEngine:
    def __init__(self, batch: int = 4, sr_mode: str = 'fp32', device=None, world: int = 1, rank: int = 0, dist=None, hp: Optional[dict] = None):
        self.batch, self.world, self.rank, self.dist = batch, world, rank, dist
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.head = RenderHead(hp=hp, sr_mode=sr_mode).to(self.device).eval()
        self.gathered = torch.empty(world * batch, 3, 512, 512, device=self.device) if world > 1 else None

    def load_params(self, decoder_params: Dict[str, torch.Tensor], sr_params: Dict[str, torch.Tensor]) -> None:
        sd = {'decoder.' + k: v for k, v in decoder_params.items()}
        sd.update({'superresolution.' + k: v for k, v in sr_params.items()})
        self.head.load_state_dict(sd, strict=True)

    @torch.no_grad()
    def step(self, planes: torch.Tensor, cameras: torch.Tensor, u_coarse: Optional[torch.Tensor] = None, u_fine=None) -> torch.Tensor:
        """planes [B,3,32,256,256], cameras [B,25] -> frames [B,3,512,512] (world == 1) or the all-gathered
        [world*B,3,512,512] (rank-major)."""
        over = {}
        if u_coarse is not None:
            over['u_coarse'] = u_coarse
        if u_fine is not None:
            over['u_fine'] = u_fine
        out = self.head.synthesis(planes, cameras, **over)['image']
        if self.world > 1:
            with capi.region('allgather'):
                self.dist.all_gather_into_tensor(self.gathered, out.contiguous())
            return self.gathered
        return out

    def profile_steps(self, inputs: Callable[[int], tuple], first: int, steps: int) -> Dict:
        """Re-run `steps` steps with CUDA events around every stage (on the launching stream)."""
        capi.PROF = capi.Profiler()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            self.step(*inputs(first + i))
        b.record()
        torch.cuda.synchronize()
        stages = capi.PROF.totals()
        capi.PROF = None
        kernel = 'sr_tc_conv (tcgen05 implicit GEMM)' if self.head.superresolution.sr_mode == 'tc' else 'conv_taps_kernel (fp32 CUDA-core direct conv)'
        return {'stages': stages, 'sr_conv_ms': stages.get('sr_conv', float('nan')), 'total_ms': a.elapsed_time(b), 'sr_kernel': kernel}
