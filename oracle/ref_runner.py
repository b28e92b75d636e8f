"""Runs the reference's OWN modules (staged byte-for-byte under oracle/_ref/ by oracle/make_ref.py) on the CPU for the timed baseline:
RaySampler -> ImportanceRenderer -> OSGDecoder -> SuperresolutionHybrid8XDC [/_Warp], wired as OSAvatarSECC_Img2plane.synthesis does
after plane production (modules/real3d/secc_img2plane.py:93-137, img2plane_baseline.py:140-147).  The full OSAvatar* model classes need
timm/mmcv/pytorch3d (absent), so the wiring of those ~20 lines is restated here; every operator is the reference's.
Test infrastructure: imported only by bench.py's CPU-baseline legs and tests/."""
import os
import sys
import types
from contextlib import contextmanager

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, 'MANIFEST.json'))


_mods = None


def modules():
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError('oracle/_ref is not staged: run `python oracle/make_ref.py` where /root/reference exists')
        sys.path.insert(0, REF_DIR)
        sys.modules.setdefault('imageio', types.ModuleType('imageio'))
        from utils.commons.hparams import hparams
        hparams.update(dict(enable_rescale_plane_regulation=False, triplane_feature_type='triplane', triplane_depth=1))
        from modules.eg3ds.volumetric_rendering.renderer import ImportanceRenderer
        from modules.eg3ds.volumetric_rendering.ray_sampler import RaySampler
        from modules.eg3ds.models.triplane import OSGDecoder
        from modules.eg3ds.models.superresolution import SuperresolutionHybrid8XDC
        _mods = dict(hparams=hparams, ImportanceRenderer=ImportanceRenderer, RaySampler=RaySampler, OSGDecoder=OSGDecoder,
                     SuperresolutionHybrid8XDC=SuperresolutionHybrid8XDC)
    return _mods


@contextmanager
def supplied_uniforms(u_coarse, u_fine):
    """torch.rand_like / torch.rand return the supplied tensors in the order the renderer draws them (renderer.py:226,281)."""
    orig_like, orig_rand = torch.rand_like, torch.rand
    torch.rand_like = lambda t, *a, **k: u_coarse.clone()
    torch.rand = lambda *size, **k: u_fine.clone()
    try:
        yield
    finally:
        torch.rand_like, torch.rand = orig_like, orig_rand


class Head:
    """decoder + renderer + SR of the reference with synthetic parameters (real3dportrait_b200.synthetic)."""

    def __init__(self, mlp, srp, S=48, S_imp=0, torso=False, warp_hparams=None, torso_model=None):
        m = modules()
        self.S, self.S_imp = S, S_imp
        self.dec = m['OSGDecoder'](32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32}).eval()
        self.dec.load_state_dict(mlp, strict=True)
        self.ren = m['ImportanceRenderer'](hp=m['hparams'])
        self.rays = m['RaySampler']()
        if torso:
            m['hparams'].update(warp_hparams or {})
            from modules.real3d.super_resolution.sr_with_ref import SuperresolutionHybrid8XDC_Warp
            self.sr = SuperresolutionHybrid8XDC_Warp(channels=32, img_resolution=512, sr_num_fp16_res=0, sr_antialias=True, channel_base=32768,
                                                     channel_max=512, fused_modconv_default='inference_only').eval()
            self.sr.torso_model = torso_model
        else:
            self.sr = m['SuperresolutionHybrid8XDC'](channels=32, img_resolution=512, sr_num_fp16_res=0, sr_antialias=True, channel_base=32768,
                                                     channel_max=512, fused_modconv_default='inference_only').eval()
        self.sr.load_state_dict(srp, strict=True)
        self.torso = torso
        self.opts = {'ray_start': 'auto', 'ray_end': 'auto', 'box_warp': 1.0, 'depth_resolution': S, 'depth_resolution_importance': S_imp,
                     'disparity_space_sampling': False, 'clamp_mode': 'softplus', 'white_back': False}

    @torch.no_grad()
    def render(self, planes, camera, u_coarse, u_fine=None, res=64):
        c2w, K = camera[:, :16].reshape(-1, 4, 4), camera[:, 16:25].reshape(-1, 3, 3)
        o, d = self.rays(c2w, K, res)
        with supplied_uniforms(u_coarse, u_fine):
            return self.ren(planes, self.dec, o, d, self.opts)

    @torch.no_grad()
    def frame(self, planes, camera, u_coarse, u_fine=None, cond=None, res=64, sr=True):
        feat, depth, wsum, valid = self.render(planes, camera, u_coarse, u_fine, res)
        N = feat.shape[0]
        fimg = feat.permute(0, 2, 1).reshape(N, feat.shape[-1], res, res).contiguous()       # secc_img2plane.py:119
        if not sr:
            return fimg
        ws = torch.ones(N, 14, 512)
        if self.torso:
            wimg = wsum.permute(0, 2, 1).reshape(N, 1, res, res)
            img, _ = self.sr(fimg[:, :3], fimg, ws, cond['ref_torso_rgb'], cond['ref_bg_rgb'], wimg, cond['segmap'], cond['kp_s'], cond['kp_d'],
                             noise_mode='none')
            return img
        return self.sr(fimg[:, :3], fimg, ws, noise_mode='none')
