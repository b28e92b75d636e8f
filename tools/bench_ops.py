"""Per-op timings with CUDA events (L2 flushed between iterations) for the roofline table in DESIGN.md:
sample_from_planes (the HBM-bound gather), channels-last repack, fused render.  `python tools/bench_ops.py [--op all]`."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import real3dportrait_b200 as r3  # noqa: E402
from real3dportrait_b200 import synthetic as syn  # noqa: E402
from oracle import real3d_oracle as orc  # noqa: E402  (only to build ray-march sample coordinates on the host)


def timeit(fn, iters=20, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                                   # evict L2 (126 MB) between timed iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--op', default='all')
    ap.add_argument('--frames', type=int, default=4)
    args = ap.parse_args()
    N = args.frames
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
    planes = syn.make_planes(N, seed=0).cuda()
    cam = syn.make_cameras(N, seed=1)
    u_c = syn.make_jitter(N, 4096, 48, 0, seed=2)[0]
    c2w, K = syn.split_camera(cam)
    o, d = orc.gen_rays(c2w, K, 64)
    t0, t1, _ = orc.auto_limits(o, d, 1.0)
    depths = orc.stratified_depths(t0, t1, 48, u_c)
    coords = (o.unsqueeze(-2) + depths * d.unsqueeze(-2)).reshape(N, -1, 3).cuda()          # [N,196608,3] = the render's samples
    pcl = r3.planes_to_channels_last(planes)
    dec = r3.OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32})
    dec.load_state_dict(syn.make_decoder_params(seed=4)); dec = dec.cuda()
    ren = r3.ImportanceRenderer()
    opts = dict(syn.RENDERING_OPTIONS, u_coarse=u_c.cuda())
    o_d, d_d = o.cuda(), d.cuda()
    out = {}
    if args.op in ('all', 'sample'):
        t = timeit(lambda: r3.sample_from_planes(None, pcl, coords, box_warp=1.0))
        byt = 103.0e6 * N                                # 25.17 MB planes + 2.36 MB coords + 75.5 MB out per frame
        out['sample_from_planes'] = {'s': t, 'algorithmic_GBps': byt / t / 1e9, 'frac_of_hbm_peak': byt / t / 1e9 / peaks['hbm_gbs']}
    if args.op in ('all', 'repack'):
        t = timeit(lambda: r3.planes_to_channels_last(planes, out=pcl.data))
        byt = 2 * 25165824.0 * N
        out['planes_to_channels_last'] = {'s': t, 'algorithmic_GBps': byt / t / 1e9, 'frac_of_hbm_peak': byt / t / 1e9 / peaks['hbm_gbs']}
    if args.op in ('all', 'render'):
        t = timeit(lambda: ren(pcl, dec, o_d, d_d, opts))
        byt = 26.51e6 * N
        out['render_fused_48'] = {'s': t, 'us_per_frame': t / N * 1e6, 'algorithmic_GBps': byt / t / 1e9, 'mlp_TFLOPs': 1.636e9 * N / t / 1e12}
        u_f = syn.make_jitter(N, 4096, 48, 48, seed=2)[1].cuda()
        opts2 = dict(opts, depth_resolution_importance=48, u_fine=u_f)
        t = timeit(lambda: ren(pcl, dec, o_d, d_d, opts2))
        out['render_fused_48_48'] = {'s': t, 'us_per_frame': t / N * 1e6, 'mlp_TFLOPs': 3.27e9 * N / t / 1e12}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
