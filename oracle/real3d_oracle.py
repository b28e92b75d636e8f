"""CPU ORACLE for the Real3D-Portrait render + super-resolution hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product path
(`real3dportrait_b200/`) never imports anything from `oracle/` and fails loudly without its CUDA library.

What it is: a from-scratch fp32 restatement (torch CPU tensors used as the array library, the same one the
reference itself computes with) of the reference algorithm, function by function, each citing the reference
file:line it follows (paths relative to the reference tree).  Written from SURVEY.md Appendix A, not copied.

Pinning: the reference ships NO tests/golden vectors for this path ("parity unpinned" by the
reference's own tests).  The oracle is therefore pinned against outputs of the reference modules themselves,
executed in the authoring container by `tests/golden/make_golden.py` (committed) and stored as fixtures in
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every fixture.

Two gather implementations are kept on purpose: `sample_planes` is an index-arithmetic restatement of
`grid_sample(bilinear, zeros, align_corners=False)`; `sample_planes_lib` uses the library op the reference calls
(used only so the timed CPU baseline has the reference's performance characteristics).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------------------------------------
# A.1  ray generation                      modules/eg3ds/volumetric_rendering/ray_sampler.py:24-63
# ----------------------------------------------------------------------------------------------------------------------

def gen_rays(cam2world: Tensor, intrinsics: Tensor, res: int) -> Tuple[Tensor, Tensor]:
    """c2w[N,4,4], K[N,3,3] -> origins[N,res^2,3], dirs[N,res^2,3].  Ray m = i*res + j: row i = y, col j = x
    (ray_sampler.py:43-44 flips the meshgrid so x varies fastest)."""
    N = cam2world.shape[0]
    fx, fy = intrinsics[:, 0, 0:1], intrinsics[:, 1, 1:2]
    cx, cy = intrinsics[:, 0, 2:3], intrinsics[:, 1, 2:3]
    sk = intrinsics[:, 0, 1:2]
    idx = torch.arange(res, dtype=torch.float32) * (1.0 / res) + (0.5 / res)          # ray_sampler.py:43
    x_cam = idx.repeat(res).unsqueeze(0).expand(N, -1)                                   # j fastest
    y_cam = idx.repeat_interleave(res).unsqueeze(0).expand(N, -1)
    x_l = (x_cam - cx + cy * sk / fy - sk * y_cam / fy) / fx                            # ray_sampler.py:51
    y_l = (y_cam - cy) / fy                                                             # ray_sampler.py:52
    pts = torch.stack([x_l, y_l, torch.ones_like(x_l), torch.ones_like(x_l)], dim=-1)    # [N,M,4]
    world = torch.einsum('nij,nmj->nmi', cam2world, pts)[..., :3]                       # ray_sampler.py:56
    origin = cam2world[:, :3, 3]
    d = world - origin[:, None, :]
    d = d / d.norm(dim=2, keepdim=True).clamp_min(1e-12)                                # F.normalize, :59
    return origin[:, None, :].expand(-1, res * res, -1).contiguous(), d


# ----------------------------------------------------------------------------------------------------------------------
# A.2  ray / box limits                    modules/eg3ds/volumetric_rendering/math_utils.py:46-98, renderer.py:121-126
# ----------------------------------------------------------------------------------------------------------------------

def ray_limits_box(o: Tensor, d: Tensor, box: float) -> Tuple[Tensor, Tensor]:
    """Slab test against the cube [-box/2, box/2]^3; invalid rays get (t0,t1)=(-1,-2) (math_utils.py:94-96)."""
    shape = o.shape[:-1]
    o = o.reshape(-1, 3)
    d = d.reshape(-1, 3)
    lo, hi = -box / 2, box / 2
    inv = 1.0 / d
    neg = inv < 0
    near = torch.where(neg, torch.full_like(o, hi), torch.full_like(o, lo))
    far = torch.where(neg, torch.full_like(o, lo), torch.full_like(o, hi))
    tn = (near - o) * inv                                                               # per-axis entry
    tf = (far - o) * inv                                                                # per-axis exit
    valid = torch.ones(o.shape[0], dtype=torch.bool)
    tmin, tmax = tn[:, 0], tf[:, 0]
    valid &= ~((tmin > tf[:, 1]) | (tn[:, 1] > tmax))                                    # math_utils.py:77
    tmin, tmax = torch.maximum(tmin, tn[:, 1]), torch.minimum(tmax, tf[:, 1])
    valid &= ~((tmin > tf[:, 2]) | (tn[:, 2] > tmax))                                    # math_utils.py:88
    tmin, tmax = torch.maximum(tmin, tn[:, 2]), torch.minimum(tmax, tf[:, 2])
    tmin = torch.where(valid, tmin, torch.full_like(tmin, -1.0))
    tmax = torch.where(valid, tmax, torch.full_like(tmax, -2.0))
    return tmin.reshape(*shape, 1), tmax.reshape(*shape, 1)


def auto_limits(o: Tensor, d: Tensor, box: float) -> Tuple[Tensor, Tensor, Tensor]:
    """renderer.py:121-126: is_ray_valid = t1 > t0; invalid rays get t0 = min(valid t0), t1 = max(valid **t0**)
    over the WHOLE call (batch-global, and the far end is filled from ray_start — sic)."""
    t0, t1 = ray_limits_box(o, d, box)
    valid = t1 > t0
    if bool(valid.any()):
        s_min, s_max = t0[valid].min(), t0[valid].max()
        t0 = torch.where(valid, t0, s_min)
        t1 = torch.where(valid, t1, s_max)
    return t0, t1, valid


# ----------------------------------------------------------------------------------------------------------------------
# A.3  stratified depths                   renderer.py:223-226, math_utils.py:101-118
# ----------------------------------------------------------------------------------------------------------------------

def stratified_depths(t0: Tensor, t1: Tensor, S: int, u: Tensor) -> Tensor:
    """t0,t1[N,M,1], u[N,M,S,1] in [0,1) -> depths[N,M,S,1]: d_k = t0 + k/(S-1)*(t1-t0) + u_k*(t1-t0)/(S-1)."""
    steps = (torch.arange(S, dtype=torch.float32) / (S - 1)).view(1, 1, S, 1)
    span = (t1 - t0).unsqueeze(-2)                                                       # [N,M,1,1]
    base = t0.unsqueeze(-2) + steps * span
    return base + u * ((t1 - t0) / (S - 1)).unsqueeze(-2)


# ----------------------------------------------------------------------------------------------------------------------
# A.4  tri-plane gather                    renderer.py:30-75
# ----------------------------------------------------------------------------------------------------------------------

#: which world axes feed (u, v) of each plane: plane0 (x,y), plane1 (x,z), plane2 (z,x)   (renderer.py:30-63;
#: the inverse of generate_planes()' axes, verified against the reference in tests/golden/make_golden.py)
PLANE_UV = ((0, 1), (0, 2), (2, 0))


def sample_planes(planes: Tensor, coords: Tensor, box_warp: float) -> Tensor:
    """planes[N,3,C,H,W], coords[N,P,3] -> [N,3,P,C].  Bilinear, zero padding, align_corners=False:
    pixel = ((g+1)*size-1)/2 with g = 2/box_warp * x  (renderer.py:71-74)."""
    N, n_planes, C, H, W = planes.shape
    P = coords.shape[1]
    g = (2.0 / box_warp) * coords
    out = torch.empty(N, n_planes, P, C, dtype=planes.dtype)
    flat = planes.reshape(N, n_planes, C, H * W)
    for p, (au, av) in enumerate(PLANE_UV):
        px = ((g[..., au] + 1) * W - 1) / 2                                              # [N,P]
        py = ((g[..., av] + 1) * H - 1) / 2
        x0, y0 = torch.floor(px), torch.floor(py)
        wx1, wy1 = px - x0, py - y0
        wx0, wy0 = (x0 + 1) - px, (y0 + 1) - py
        acc = torch.zeros(N, C, P, dtype=planes.dtype)
        for dy, wy in ((0, wy0), (1, wy1)):
            for dx, wx in ((0, wx0), (1, wx1)):
                xi, yi = (x0 + dx).long(), (y0 + dy).long()
                inb = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
                lin = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1))                      # [N,P]
                tex = torch.gather(flat[:, p], 2, lin[:, None, :].expand(-1, C, -1))     # [N,C,P]
                acc = acc + tex * (wx * wy * inb)[:, None, :]
        out[:, p] = acc.permute(0, 2, 1)
    return out


def sample_planes_lib(planes: Tensor, coords: Tensor, box_warp: float) -> Tensor:
    """Same result through the library op the reference calls (renderer.py:74); used for the timed CPU baseline."""
    N, n_planes, C, H, W = planes.shape
    P = coords.shape[1]
    g = (2.0 / box_warp) * coords
    grids = torch.stack([g[..., [au, av]] for au, av in PLANE_UV], dim=1).reshape(N * n_planes, 1, P, 2)
    o = F.grid_sample(planes.reshape(N * n_planes, C, H, W), grids, mode='bilinear', padding_mode='zeros',
                      align_corners=False)
    return o.permute(0, 3, 2, 1).reshape(N, n_planes, P, C)


#: third projected coordinate of each plane (the tri-grid depth axis): coords @ inv(plane_axes) = (x,y,z), (x,z,y), (z,x,y)
PLANE_W = (2, 1, 1)


def sample_trigrids(planes: Tensor, coords: Tensor, box_warp: float, depth: int) -> Tensor:
    """sample_from_trigrids (renderer.py:78-89; `triplane_feature_type: trigrid_v2`, egs/os_avatar/img2plane.yaml:65-66):
    planes[N,3,C*D,H,W] viewed as [N*3,C,D,H,W] (channel index = c*D + d), trilinear, zero padding, align_corners=False -> [N,3,P,C]."""
    N, n_planes, CD, H, W = planes.shape
    C, D = CD // depth, depth
    P = coords.shape[1]
    g = (2.0 / box_warp) * coords
    grid = planes.reshape(N, n_planes, C, D, H * W)
    out = torch.empty(N, n_planes, P, C, dtype=planes.dtype)
    for p, ((au, av), aw) in enumerate(zip(PLANE_UV, PLANE_W)):
        px = ((g[..., au] + 1) * W - 1) / 2
        py = ((g[..., av] + 1) * H - 1) / 2
        pz = ((g[..., aw] + 1) * D - 1) / 2
        x0, y0, z0 = torch.floor(px), torch.floor(py), torch.floor(pz)
        wx, wy, wz = (((x0 + 1) - px, px - x0), ((y0 + 1) - py, py - y0), ((z0 + 1) - pz, pz - z0))
        acc = torch.zeros(N, C, P, dtype=planes.dtype)
        gp = grid[:, p].reshape(N, C, D * H * W)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    xi, yi, zi = (x0 + dx).long(), (y0 + dy).long(), (z0 + dz).long()
                    inb = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H) & (zi >= 0) & (zi < D)
                    lin = (zi.clamp(0, D - 1) * H + yi.clamp(0, H - 1)) * W + xi.clamp(0, W - 1)       # [N,P]
                    tex = torch.gather(gp, 2, lin[:, None, :].expand(-1, C, -1))                      # [N,C,P]
                    acc = acc + tex * (wx[dx] * wy[dy] * wz[dz] * inb)[:, None, :]
        out[:, p] = acc.permute(0, 2, 1)
    return out


def sample_trigrids_lib(planes: Tensor, coords: Tensor, box_warp: float, depth: int) -> Tensor:
    """Same through F.grid_sample (5-D), as the reference calls it."""
    N, n_planes, CD, H, W = planes.shape
    C, D = CD // depth, depth
    P = coords.shape[1]
    g = (2.0 / box_warp) * coords
    grids = torch.stack([g[..., [au, av, aw]] for (au, av), aw in zip(PLANE_UV, PLANE_W)], dim=1).reshape(N * n_planes, 1, 1, P, 3)
    o = F.grid_sample(planes.reshape(N * n_planes, C, D, H, W), grids, mode='bilinear', padding_mode='zeros', align_corners=False)
    return o.permute(0, 4, 3, 2, 1).reshape(N, n_planes, P, C)


# ----------------------------------------------------------------------------------------------------------------------
# A.5  OSG decoder                         modules/img2plane/triplane.py:122-146, networks_stylegan2.py:99-131
# ----------------------------------------------------------------------------------------------------------------------

def softplus(x: Tensor) -> Tensor:
    """torch.nn.Softplus(beta=1, threshold=20)."""
    return torch.where(x > 20, x, torch.log1p(torch.exp(torch.clamp(x, max=20.0))))


def decode(feat3: Tensor, mlp: Dict[str, Tensor]) -> Tuple[Tensor, Tensor]:
    """feat3[N,3,P,C] -> rgb[N,P,32], sigma[N,P,1].  mean over planes (triplane.py:136); FC gains 1/sqrt(fan_in)
    (networks_stylegan2.py:113-114); rgb = sigmoid*1.002-0.001 (triplane.py:144)."""
    x = feat3.mean(1)
    w1, b1, w2, b2 = mlp['net.0.weight'], mlp['net.0.bias'], mlp['net.2.weight'], mlp['net.2.bias']
    h = softplus(x @ (w1 * (1.0 / math.sqrt(w1.shape[1]))).t() + b1)
    y = h @ (w2 * (1.0 / math.sqrt(w2.shape[1]))).t() + b2
    return torch.sigmoid(y[..., 1:]) * 1.002 - 0.001, y[..., 0:1]


def run_model(planes: Tensor, mlp: Dict[str, Tensor], coords: Tensor, box_warp: float, lib: bool = False, depth: int = 0):
    """renderer.py:169-188 (inference branch: no plane rescale, no density noise).  depth > 0: tri-grids (trigrid / trigrid_v2)."""
    if depth > 0:
        f = (sample_trigrids_lib if lib else sample_trigrids)(planes, coords, box_warp, depth)
    else:
        f = (sample_planes_lib if lib else sample_planes)(planes, coords, box_warp)
    return decode(f, mlp)


# ----------------------------------------------------------------------------------------------------------------------
# A.6  ray marcher                         modules/eg3ds/volumetric_rendering/ray_marcher.py:25-57
# ----------------------------------------------------------------------------------------------------------------------

def ray_march(colors: Tensor, sigmas: Tensor, depths: Tensor, white_back: bool = False):
    """colors[N,M,S,C], sigmas[N,M,S,1], depths[N,M,S,1] -> rgb[N,M,C], depth[N,M,1], weights[N,M,S-1,1]."""
    delta = depths[:, :, 1:] - depths[:, :, :-1]
    c_mid = (colors[:, :, :-1] + colors[:, :, 1:]) / 2
    s_mid = softplus((sigmas[:, :, :-1] + sigmas[:, :, 1:]) / 2 - 1)                      # ray_marcher.py:33
    d_mid = (depths[:, :, :-1] + depths[:, :, 1:]) / 2
    alpha = 1 - torch.exp(-(s_mid * delta))
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :, :1]), 1 - alpha + 1e-10], -2), -2)[:, :, :-1]
    w = alpha * trans
    rgb = (w * c_mid).sum(-2)
    wsum = w.sum(2)
    depth = (w * d_mid).sum(-2) / wsum
    depth = torch.nan_to_num(depth, float('inf'))
    depth = torch.clamp(depth, depths.min(), depths.max())                               # batch-global, :50
    if white_back:
        rgb = rgb + 1 - wsum
    return rgb * 2 - 1, depth, w


# ----------------------------------------------------------------------------------------------------------------------
# A.7  importance sampling + merge         renderer.py:197-207, 234-297
# ----------------------------------------------------------------------------------------------------------------------

def importance_depths(depths: Tensor, weights: Tensor, u: Tensor) -> Tensor:
    """depths[N,M,S,1], weights[N,M,S-1,1], u[N*M,Ni] -> fine depths [N,M,Ni,1] (unsorted)."""
    N, M, S, _ = depths.shape
    z = depths.reshape(N * M, S)
    w = weights.reshape(N * M, S - 1)
    ninf = torch.full_like(w[:, :1], float('-inf'))
    wp = torch.cat([ninf, w, ninf], 1)
    m = torch.maximum(wp[:, :-1], wp[:, 1:])                                             # max_pool1d(2,1,pad 1): S values
    a = 0.5 * (m[:, :-1] + m[:, 1:]) + 0.01                                              # avg_pool1d(2,1) + .01: S-1
    bins = 0.5 * (z[:, :-1] + z[:, 1:])                                                  # S-1 midpoints
    p = a[:, 1:-1] + 1e-5                                                                # S-3 weights
    pdf = p / p.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)           # S-2
    idx = torch.searchsorted(cdf, u.contiguous(), right=True)
    lo = (idx - 1).clamp_min(0)
    hi = idx.clamp_max(p.shape[1])
    c_lo, c_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    b_lo, b_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    den = c_hi - c_lo
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    return (b_lo + (u - c_lo) / den * (b_hi - b_lo)).reshape(N, M, -1, 1)


def unify(d1, c1, s1, d2, c2, s2):
    d = torch.cat([d1, d2], -2)
    c = torch.cat([c1, c2], -2)
    s = torch.cat([s1, s2], -2)
    _, order = torch.sort(d, dim=-2)
    return (torch.gather(d, -2, order), torch.gather(c, -2, order.expand(-1, -1, -1, c.shape[-1])),
            torch.gather(s, -2, order))


def render(planes: Tensor, mlp: Dict[str, Tensor], ray_o: Tensor, ray_d: Tensor, *, S: int, S_imp: int = 0,
           box_warp: float = 1.0, white_back: bool = False, u_coarse: Tensor, u_fine: Optional[Tensor] = None,
           lib: bool = False, trigrid_depth: int = 0):
    """ImportanceRenderer.forward with 'auto' limits (renderer.py:118-167); jitter supplied by the caller.
    Returns rgb[N,M,C], depth[N,M,1], weights_sum[N,M,1], is_ray_valid[N,M,1]."""
    N, M, _ = ray_o.shape
    t0, t1, valid = auto_limits(ray_o, ray_d, box_warp)
    d_c = stratified_depths(t0, t1, S, u_coarse)
    xyz = (ray_o.unsqueeze(-2) + d_c * ray_d.unsqueeze(-2)).reshape(N, -1, 3)
    col, sig = run_model(planes, mlp, xyz, box_warp, lib, trigrid_depth)
    col, sig = col.reshape(N, M, S, -1), sig.reshape(N, M, S, 1)
    if S_imp > 0:
        _, _, w = ray_march(col, sig, d_c, white_back)
        d_f = importance_depths(d_c, w, u_fine)
        xyz = (ray_o.unsqueeze(-2) + d_f * ray_d.unsqueeze(-2)).reshape(N, -1, 3)
        col_f, sig_f = run_model(planes, mlp, xyz, box_warp, lib, trigrid_depth)
        d_a, c_a, s_a = unify(d_c, col, sig, d_f, col_f.reshape(N, M, S_imp, -1), sig_f.reshape(N, M, S_imp, 1))
        rgb, depth, w = ray_march(c_a, s_a, d_a, white_back)
    else:
        rgb, depth, w = ray_march(col, sig, d_c, white_back)
    return rgb, depth, w.sum(2), valid


# ----------------------------------------------------------------------------------------------------------------------
# A.8  super-resolution                    superresolution.py:331-359, networks_stylegan2.py:37-94,286-473,
#                                          torch_utils/ops/{conv2d_resample,upfirdn2d,bias_act}.py
# ----------------------------------------------------------------------------------------------------------------------

def fir_kernel() -> Tensor:
    """upfirdn2d.setup_filter([1,3,3,1]) (upfirdn2d.py:72-116): outer product, normalised to sum 1."""
    f = torch.tensor([1.0, 3.0, 3.0, 1.0])
    f = torch.outer(f, f)
    return f / f.sum()


def affine_styles(w_lat: Tensor, aff_w: Tensor, aff_b: Tensor) -> Tensor:
    """FullyConnectedLayer(512, Cin, bias_init=1) with linear act (networks_stylegan2.py:113-127)."""
    return w_lat @ (aff_w * (1.0 / math.sqrt(aff_w.shape[1]))).t() + aff_b


def fold_weight(weight: Tensor, styles: Tensor, demod: bool) -> Tensor:
    """modulated_conv2d's per-sample weights (networks_stylegan2.py:63-70): [N,O,I,kh,kw]."""
    w = weight.unsqueeze(0) * styles[:, None, :, None, None]
    if demod:
        w = w * (w.square().sum(dim=[2, 3, 4], keepdim=True) + 1e-8).rsqrt()
    return w


def fir_pad(y: Tensor, pad: Tuple[int, int, int, int], gain: float) -> Tensor:
    """upfirdn2d with up=down=1 (upfirdn2d.py:171-215): zero-pad, true convolution with the 4x4 filter, *gain.
    Written as 16 shifted adds (independent of the library conv)."""
    f = fir_kernel().flip(0, 1) * gain
    yp = F.pad(y, pad)
    Ho, Wo = yp.shape[-2] - 3, yp.shape[-1] - 3
    out = torch.zeros(*y.shape[:-2], Ho, Wo, dtype=y.dtype)
    for a in range(4):
        for b in range(4):
            out = out + f[a, b] * yp[..., a:a + Ho, b:b + Wo]
    return out


def upsample2x(img: Tensor) -> Tensor:
    """upfirdn2d.upsample2d(img, f) (upfirdn2d.py:317-354): zero-insert x2, pad (2,1,2,1), FIR, gain 4."""
    N, C, H, W = img.shape
    z = torch.zeros(N, C, H * 2, W * 2, dtype=img.dtype)
    z[:, :, ::2, ::2] = img
    return fir_pad(z, (2, 1, 2, 1), 4.0)


def lrelu_gain(x: Tensor, bias: Tensor) -> Tensor:
    """bias_act(act='lrelu', alpha 0.2, gain sqrt 2) (bias_act.py:54-122)."""
    return F.leaky_relu(x + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2.0)


def mod_conv(x: Tensor, wf: Tensor, up: int) -> Tensor:
    """Per-sample convolution with folded weights wf[N,O,I,3,3].
    up=1: correlation, pad 1 (conv2d_resample.py:136-138).
    up=2: conv_transpose2d(stride 2, pad 0) with the UNflipped weight, then FIR pad 1 gain 4
          (conv2d_resample.py:116-133 with flip_weight=False from networks_stylegan2.py:334)."""
    outs = []
    for n in range(x.shape[0]):
        if up == 1:
            outs.append(F.conv2d(x[n:n + 1], wf[n], padding=wf.shape[-1] // 2))
        else:
            y = F.conv_transpose2d(x[n:n + 1], wf[n].transpose(0, 1), stride=2)
            outs.append(fir_pad(y, (1, 1, 1, 1), 4.0))
    return torch.cat(outs, 0)


def synthesis_layer(x, w_lat, p: Dict[str, Tensor], prefix: str, up: int) -> Tensor:
    """SynthesisLayer.forward with noise_mode='none' (networks_stylegan2.py:322-342)."""
    s = affine_styles(w_lat, p[prefix + 'affine.weight'], p[prefix + 'affine.bias'])
    wf = fold_weight(p[prefix + 'weight'], s, True)
    return lrelu_gain(mod_conv(x, wf, up), p[prefix + 'bias'])


def to_rgb(x, w_lat, p: Dict[str, Tensor], prefix: str) -> Tensor:
    """ToRGBLayer.forward (networks_stylegan2.py:365-370): styles*1/sqrt(Cin), no demod, linear."""
    cin = p[prefix + 'weight'].shape[1]
    s = affine_styles(w_lat, p[prefix + 'affine.weight'], p[prefix + 'affine.bias']) * (1.0 / math.sqrt(cin))
    wf = fold_weight(p[prefix + 'weight'], s, False)
    y = torch.cat([F.conv2d(x[n:n + 1], wf[n]) for n in range(x.shape[0])], 0)
    return y + p[prefix + 'bias'].view(1, -1, 1, 1)


def synthesis_block(x, img, ws3, p, prefix) -> Tuple[Tensor, Tensor]:
    """SynthesisBlock.forward, architecture 'skip', in_channels != 0 (networks_stylegan2.py:429-473)."""
    x = synthesis_layer(x, ws3[:, 0], p, prefix + 'conv0.', up=2)
    x = synthesis_layer(x, ws3[:, 1], p, prefix + 'conv1.', up=1)
    img = upsample2x(img) + to_rgb(x, ws3[:, 2], p, prefix + 'torgb.')
    return x, img


def resize_bilinear(x: Tensor, size: int) -> Tensor:
    """F.interpolate(bilinear, align_corners=False, antialias=True) for UP-scaling (antialias is a no-op there;
    superresolution.py:351-355).  src = (dst+0.5)*in/out-0.5, clamped at 0 below; taps clamped at the border."""
    N, C, H, W = x.shape
    def taps(n_in, n_out):
        src = (torch.arange(n_out, dtype=torch.float32) + 0.5) * (n_in / n_out) - 0.5
        src = src.clamp_min(0)
        i0 = src.floor().long().clamp_max(n_in - 1)
        i1 = (i0 + 1).clamp_max(n_in - 1)
        t = src - i0.float()
        return i0, i1, t
    y0, y1, ty = taps(H, size)
    x0, x1, tx = taps(W, size)
    rows = x[:, :, y0] * (1 - ty).view(1, 1, -1, 1) + x[:, :, y1] * ty.view(1, 1, -1, 1)
    return rows[..., x0] * (1 - tx) + rows[..., x1] * tx


def superres(rgb: Tensor, x: Tensor, ws: Tensor, p: Dict[str, Tensor]) -> Tensor:
    """SuperresolutionHybrid8XDC.forward (superresolution.py:348-359), noise_mode='none', fp32."""
    ws3 = ws[:, -1:, :].repeat(1, 3, 1)
    if x.shape[-1] != 128:
        x, rgb = resize_bilinear(x, 128), resize_bilinear(rgb, 128)
    x, rgb = synthesis_block(x, rgb, ws3, p, 'block0.')
    x, rgb = synthesis_block(x, rgb, ws3, p, 'block1.')
    return rgb


def feature_image(rgb_feat: Tensor, res: int) -> Tensor:
    """[N,M,C] -> [N,C,res,res] (secc_img2plane.py:118-119)."""
    N, M, C = rgb_feat.shape
    return rgb_feat.permute(0, 2, 1).reshape(N, C, res, res).contiguous()


def frame(planes, mlp, sr_params, cam2world, intrinsics, *, res=64, S=48, S_imp=0, box_warp=1.0, u_coarse, u_fine=None,
          lib=False) -> Dict[str, Tensor]:
    """The synthesis() render head (secc_img2plane.py:93-137) from tri-planes to the 512^2 image."""
    o, d = gen_rays(cam2world, intrinsics, res)
    feat, depth, wsum, valid = render(planes, mlp, o, d, S=S, S_imp=S_imp, box_warp=box_warp, u_coarse=u_coarse,
                                      u_fine=u_fine, lib=lib)
    fimg = feature_image(feat, res)
    ws = torch.ones(planes.shape[0], 14, 512)
    sr = superres(fimg[:, :3], fimg, ws, sr_params)
    return {'image_raw': fimg[:, :3].clamp(-1, 1), 'image': sr.clamp(-1, 1), 'image_feature': fimg,
            'image_depth': feature_image(depth, res), 'weights_img': feature_image(wsum, res), 'is_ray_valid': valid}


# ----------------------------------------------------------------------------------------------------------------------
# torso head                               modules/real3d/super_resolution/sr_with_ref.py:16-162 (fuse mode 'v2')
# ----------------------------------------------------------------------------------------------------------------------

def aa_down2(x: Tensor) -> Tensor:
    """F.interpolate(scale 1/2, bilinear, align_corners=False, antialias=True) restated: triangle filter of support 2 around
    centre 2(i+0.5): taps 2i-1..2i+2 with weights [1,3,3,1]/8, clipped to the image and renormalised (sr_with_ref.py:79-82)."""
    def axis(t, dim):
        n = t.shape[dim]
        k = torch.tensor([0.25, 0.75, 0.75, 0.25])
        idx = torch.arange(n // 2)[:, None] * 2 - 1 + torch.arange(4)[None, :]              # [n/2,4]
        ok = (idx >= 0) & (idx < n)
        w = (k[None, :] * ok) / (k[None, :] * ok).sum(1, keepdim=True)
        g = t.index_select(dim, idx.clamp(0, n - 1).reshape(-1))
        shp = list(t.shape); shp[dim:dim + 1] = [n // 2, 4]
        g = g.reshape(shp)
        wshape = [1] * g.ndim; wshape[dim] = n // 2; wshape[dim + 1] = 4
        return (g * w.reshape(wshape)).sum(dim + 1)
    return axis(axis(x, 2), 3)


def conv_plain(x: Tensor, p: Dict[str, Tensor], name: str, act: Optional[float] = None) -> Tensor:
    """nn.Conv2d(stride 1, 'same' padding) [+ nn.LeakyReLU(act)]."""
    w = p[name + '.weight']
    y = F.conv2d(x, w, p[name + '.bias'], padding=w.shape[-1] // 2)
    return F.leaky_relu(y, act) if act is not None else y


def synthesis_block_noup(x, img, ws3, p, prefix):
    """SynthesisBlockNoUp.forward (superresolution.py:209-258): two non-upsampling layers, img += torgb (no upsampling of img)."""
    x = synthesis_layer(x, ws3[:, 0], p, prefix + 'conv0.', up=1)
    x = synthesis_layer(x, ws3[:, 1], p, prefix + 'conv1.', up=1)
    return x, img + to_rgb(x, ws3[:, 2], p, prefix + 'torgb.')


def superres_warp(rgb, x, ws, ref_torso_rgb, ref_bg_rgb, weights_img, segmap, kp_s, kp_d, p: Dict[str, Tensor], torso_model, head_threshold=0.9,
                  mode: str = 'v2'):
    """SuperresolutionHybrid8XDC_Warp.forward, torso_model_version 'v2', eval mode; `mode` = htbsr_head_weight_fuse_mode:
    'v1' sr_with_ref.py:96-104, 'v2' (the released configuration) :106-124, 'v3' :126-152."""
    ws3 = ws[:, -1:, :].expand(rgb.shape[0], 3, -1)
    if x.shape[-1] != 128:
        x, rgb = resize_bilinear(x, 128), resize_bilinear(rgb, 128)
    rgb_256 = resize_bilinear(rgb, 256)
    weights_256 = resize_bilinear(weights_img, 256)
    ref_torso_256, ref_bg_256 = aa_down2(ref_torso_rgb), aa_down2(ref_bg_rgb)
    x, rgb = synthesis_block(x, rgb, ws3, p, 'block0.')
    rgb_torso, ret = torso_model(ref_torso_256, segmap, kp_s, kp_d, rgb_256, weights_256, cal_loss=True, target_torso_mask=None)
    x_torso = conv_plain(ret['deformed_torso_hid'], p, 'torso_encoder.0')
    x_bg = conv_plain(conv_plain(conv_plain(ref_bg_256, p, 'bg_encoder.0', 0.01), p, 'bg_encoder.2', 0.01), p, 'bg_encoder.4')
    torso_occ = ret['occlusion_2'] if ret['occlusion_2'].shape[-1] == 256 else resize_bilinear(ret['occlusion_2'], 256)
    if mode == 'v1':                                                                       # :96-104: plain alpha blend of rgb AND features, no head/torso convs
        alpha = weights_256
        rgb = rgb * alpha + rgb_torso * (1 - alpha)
        x = x * alpha + x_torso * (1 - alpha)
    else:
        if mode == 'v3':                                                                   # :129-132: a small conv net post-processes the head mask, capped by the weights
            inp = torch.cat([rgb.clamp(-1, 1) / 2 + 0.5, weights_256, rgb_torso.clamp(-1, 1) / 2 + 0.5], dim=1)
            a_ = conv_plain(conv_plain(conv_plain(inp, p, 'head_torso_alpha_predictor.0', 0.01), p, 'head_torso_alpha_predictor.2', 0.01), p,
                            'head_torso_alpha_predictor.4')
            alpha = torch.minimum(torch.sigmoid(a_), weights_256)
        else:
            alpha = weights_256                                                            # :108-109 (the masked assignment is a no-op)
        rgb = rgb * alpha + rgb_torso * (1 - alpha)
        x = torch.cat([x * alpha, x_torso * (1 - alpha)], dim=1)
        x = conv_plain(conv_plain(x, p, 'fuse_head_torso_convs.0', 0.01), p, 'fuse_head_torso_convs.2')
        x, rgb = synthesis_block_noup(x, rgb, ws3, p, 'head_torso_block.')
    thr = head_threshold
    if mode == 'v3':                                                                       # :141-143 (eval): batch-wide 5 % quantile of the mask values above 0.05
        sel = alpha[alpha > 0.05]
        thr = max(float(sel.quantile(0.05)), head_threshold)
    head_occ = torch.where(alpha > thr, torch.ones_like(alpha), alpha)
    person = (torso_occ + head_occ).clamp(0, 1)
    rgb = rgb * person + ref_bg_256 * (1 - person)
    x = torch.cat([x * person, x_bg * (1 - person)], dim=1)
    x = conv_plain(conv_plain(conv_plain(x, p, 'fuse_fg_bg_convs.0', 0.01), p, 'fuse_fg_bg_convs.2', 0.01), p, 'fuse_fg_bg_convs.4')
    x, rgb = synthesis_block(x, rgb, ws3, p, 'block1.')
    return rgb, ret


# ----------------------------------------------------------------------------------------------------------------------
# large_sr                                 modules/eg3ds/models/superresolution.py:263-345 (LargeSynthesisBlock0/1, ResBlock2d)
# ----------------------------------------------------------------------------------------------------------------------

def resblock(x: Tensor, p: Dict[str, Tensor], prefix: str) -> Tensor:
    """ResBlock2d.forward (superresolution.py:283-288): relu(conv2(relu(conv1(x)))) + x."""
    out = F.relu(conv_plain(x, p, prefix + 'conv1'))
    out = F.relu(conv_plain(out, p, prefix + 'conv2'))
    return out + x


def superres_large(rgb: Tensor, x: Tensor, ws: Tensor, p: Dict[str, Tensor], n_res: int) -> Tensor:
    """SuperresolutionHybrid8XDC.forward with large_sr=True: each LargeSynthesisBlock = SynthesisBlock -> n_res ResBlock2d ->
    rgb = rgb + to_rgb(x) (superresolution.py:309-312,325-329)."""
    ws3 = ws[:, -1:, :].repeat(1, 3, 1)
    if x.shape[-1] != 128:
        x, rgb = resize_bilinear(x, 128), resize_bilinear(rgb, 128)
    for blk in ('block0', 'block1'):
        pre = {k[len(blk) + 7:]: v for k, v in p.items() if k.startswith(blk + '.block.')}
        x, rgb = synthesis_block(x, rgb, ws3, {f'{blk}.{k}': v for k, v in pre.items()}, blk + '.')
        for i in range(n_res):
            x = resblock(x, p, f'{blk}.resblocks.{i}.')
        rgb = rgb + conv_plain(x, p, f'{blk}.to_rgb')
    return rgb
