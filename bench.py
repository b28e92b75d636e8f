"""bench.py — headline benchmark of the render + SR hot path (BASELINE.json: rendered frames/sec @512^2, 64^2 NeRF,
48 samples/ray).  One "step" = one batch of `--batch` frames per GPU: tri-planes (resident in HBM, reference NCHW
layout) -> channels-last repack -> fused render -> SR -> 512^2 fp32 frames [-> NCCL all-gather of the step's frames when
N > 1].  Prints ONE JSON line on rank 0.  See DESIGN.md §Measurement for every field.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR_GFLOP_PER_FRAME = 197.63          # SURVEY.md §8d (FlopCounter probe of the reference SR, exact)
SAMPLE_BYTES_PER_FRAME = 103.0e6     # stand-alone sample_from_planes op, S=48


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'src': 'fallback'}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe).  Read through NVML in-process (what nvidia-smi
    itself calls) every 20 ms: spawning `nvidia-smi` several times a second from a side thread stalled the driver long enough to cost
    the PCIe-bound e2e loop a quarter of its H2D rate (39 vs 55 GB/s on the same box).  Falls back to the nvidia-smi query if NVML
    cannot be loaded."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    BITS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows, self.source = index, threading.Event(), [], 'nvml'
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(index)
            self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(f'{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0')
            self.max_sm = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml, self.source = None, 'nvidia-smi'

    def _sample_nvml(self):
        n = self.nvml
        sm = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.rows.append([str(sm), str(self.max_sm)] + [('Active' if mask & bit else 'Not Active') for bit in (0x8, 0x40, 0x20, 0x4)])

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml is not None else 0.5)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_min_mhz': sm[0] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.rows), 'source': self.source}


def host_threads():
    """Threads the CPU baseline may really use: min(cpu_count, affinity mask, cgroup cpu quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, per = open('/sys/fs/cgroup/cpu.max').read().split()
        if q != 'max':
            n = max(1, min(n, int(float(q) / float(per))))
    except Exception:
        pass
    return n


def cpu_frame_fn(n_frames):
    """The oracle (port of the reference's PyTorch CPU path: same library ops) on `n_frames` frames of the workload."""
    from oracle import real3d_oracle as orc
    from real3dportrait_b200 import synthetic as syn
    planes, cam = syn.make_planes(n_frames, seed=0), syn.make_cameras(n_frames, seed=1)
    u_c, _ = syn.make_jitter(n_frames, 4096, 48, 0, seed=2)
    mlp, srp = syn.make_decoder_params(seed=4), syn.make_sr_params(seed=5)
    c2w, K = syn.split_camera(cam)
    def run():
        with torch.no_grad():
            return orc.frame(planes, mlp, srp, c2w, K, u_coarse=u_c, lib=True)['image']
    return run


def time_cpu(n_frames, reps, warm=1, budget_s=45.0):
    """Median frames/s of the oracle on the host.  Tries the full thread count and (if that is > 32) 32 threads, because
    torch's CPU convolutions often run SLOWER when heavily over-threaded; reports the best, with the count used."""
    fn = cpu_frame_fn(n_frames)
    best = None
    cands = [host_threads()] + ([32] if host_threads() > 32 else [])
    for nt in cands:
        torch.set_num_threads(nt)
        t_start = time.perf_counter()
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
            if time.perf_counter() - t_start > budget_s / len(cands):
                break
        ts.sort()
        fps = n_frames / ts[len(ts) // 2]
        if best is None or fps > best[0]:
            best = (fps, ts, nt)
    return best


def config_of(args, extra=None):
    c = {'workload': f'BASELINE configs[2]: batch={args.batch} frames/step/GPU, 64x64 rays x 48 samples on 3x32x256x256 fp32 '
                     f'tri-planes + SR 64^2->512^2 (197.63 GFLOP/frame), random-init weights',
         'frames_per_step_per_gpu': args.batch, 'samples_per_ray': 48, 'render_res': 64, 'out_res': 512,
         'parallelism': f'frames sharded over {args.gpus} GPU(s), no data-path collective except the frame all-gather'}
    c.update(extra or {})
    return c


def run_reference(args):
    """--impl reference: the reference's CPU path for this workload (oracle port; /root/reference cannot travel to the box)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    fps, ts, nthr = time_cpu(1, steps, warm)
    line = {'impl': 'reference', 'metric': 'rendered frames/sec @512^2 (64^2 NeRF, 48 samples/ray)', 'value': fps, 'unit': 'frames/s',
            'n_gpus': args.gpus, 'steps': steps, 'warmup': warm, 'ms_per_step': 1e3 * ts[len(ts) // 2], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config_of(args, {'note': 'each step = ONE frame of the batch on the host CPU (bounded sample)'}),
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': nthr, 'kind': 'port',
                             'sample': f'{len(ts)} x 1 frame (render 64^2x48 + SR), torch CPU threads={nthr} of {host_threads()} usable'},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=32)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--pool', type=int, default=32, help='distinct resident frames per GPU (32 x 25 MB = 805 MB > L2)')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--sr-mode', default=None, choices=[None, 'fp32', 'tc'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of replaying one CUDA graph per step')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        return run_reference(args)

    import ctypes as C
    import real3dportrait_b200 as r3
    from real3dportrait_b200 import _capi, synthetic as syn, engine

    rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    L = _capi.lib()
    _capi.check(L.r3dp_device_info(None, None, None))

    sr_mode = args.sr_mode or engine.default_sr_mode()
    eng = engine.FrameEngine(batch=args.batch, sr_mode=sr_mode, device=dev, world=world, rank=rank, dist=dist, use_graph=not args.no_graph,
                             hp={'num_samples_fine': 0})
    eng.load_params(syn.make_decoder_params(seed=4), syn.make_sr_params(seed=5))
    B, P = args.batch, max(args.pool, args.batch)
    # resident inputs: every rank owns its own shard of the clip (different seeds per rank)
    planes = syn.make_planes(P, seed=100 + rank).to(dev)
    cams = syn.make_cameras(P, seed=200 + rank).to(dev)
    u_c = syn.make_jitter(P, 4096, 48, 0, seed=300 + rank)[0].to(dev)
    sl = lambda i: slice((i % (P // B)) * B, (i % (P // B)) * B + B)        # P/B distinct resident batches, cycled

    def step(i):
        s = sl(i)
        return eng.step(planes[s], cams[s], u_c[s])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    eng.prepare([(planes[sl(i)], cams[sl(i)], u_c[sl(i)]) for i in range(P // B)])     # one step graph per resident batch: inputs are read in place
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = L.r3dp_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    eng.wait_gather()                              # the last steps' frame exchanges (side stream) are inside the timed region
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = L.r3dp_launch_count() - launches0
    if eng.graph is not None or eng.inplace:        # kernels replayed from the captured graph are not re-counted by the library
        launches += eng.launches_per_step * args.steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    fps = world * B * args.steps / (ms / 1e3)

    # ---- dominant kernel live timing (SR convolutions) over the same steps, CUDA events on the launching stream
    prof = eng.profile_steps(lambda i: (planes[sl(i)], cams[sl(i)], u_c[sl(i)]), args.warmup, args.steps)
    barrier()

    # ---- end-to-end through the public call with HOST buffers (pinned), H2D + D2H inside the timed region
    with engine.gpu_local_cpus(torch.cuda.current_device()):                    # pinned staging buffers on the GPU's NUMA node
        h_planes = planes[:B].cpu().pin_memory(); h_cams = cams[:B].cpu().pin_memory(); h_u = u_c[:B].cpu().pin_memory()
        h_out = torch.empty(B, 3, 512, 512, dtype=torch.float32).pin_memory()
    def e2e_step():
        eng.step_host(h_planes, h_cams, h_u, h_out)                 # public host-buffer call: H2D -> step -> D2H, pipelined over 3 streams
    for _ in range(3):
        e2e_step()
    eng.sync_host()
    barrier()
    ksteps = max(8, min(args.steps, 32))
    t_host0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ksteps):
        e2e_step()
    eng.sync_host()                                                  # the last D2H has landed in h_out
    e1.record()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t_host0) * 1e3
    clocks = sampler.summary()                    # sampled from the start of the timed region to the end of the e2e region (GPU busy throughout)
    te = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_fps = world * B * ksteps / (float(te.item()) / 1e3)
    h2d = h_planes.numel() * 4 + h_cams.numel() * 4 + h_u.numel() * 4
    d2h = h_out.numel() * 4

    # ---- north_star's HBM roofline: the stand-alone sample_from_planes op on one step's render samples (L2 flushed between iterations)
    hbm = None
    if rank == 0:
        from real3dportrait_b200 import renderer as _ren
        pcl = _ren.planes_to_channels_last(planes[:B])
        ro, rd = r3.RaySampler()(cams[:B, :16].reshape(-1, 4, 4), cams[:B, 16:25].reshape(-1, 3, 3), 64)
        depth_mid = 2.15 + 1.1 * (torch.arange(48, device=dev).view(1, 1, 48, 1) + u_c[:B]) / 47.0            # samples spread through the box
        coords = (ro.unsqueeze(-2) + depth_mid * rd.unsqueeze(-2)).reshape(B, -1, 3).contiguous()
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        ts = []
        for it in range(13):
            flush.zero_()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); _ren.sample_from_planes(None, pcl, coords, box_warp=1.0); a1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(a0.elapsed_time(a1))
        ts.sort()
        t_ms = ts[len(ts) // 2]
        hbm = {'kernel': 'triplane_sample_kernel (sample_from_planes contract, mean NOT fused)', 'ms': t_ms, 'algorithmic_bytes': SAMPLE_BYTES_PER_FRAME * B,
               'achieved': SAMPLE_BYTES_PER_FRAME * B / (t_ms * 1e-3) / 1e9, 'unit': 'GB/s'}
        del flush
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    hbm.update({'peak': pk['hbm_gbs'], 'frac': hbm['achieved'] / pk['hbm_gbs'], 'peak_source': pk['src'] + ' copy bandwidth', 'bound': 'hbm',
                'timing': '10 timed iterations, L2 flushed (256 MB write) before each, CUDA events, median'})
    # dominant kernel = the tensor-core conv (4 launches/step); its time is measured by CUDA-event pairs around every launch (library hook),
    # falling back to the whole SR-conv stage for the fp32 path
    if prof.get('conv_kernel_ms'):
        k_ms_per_step, k_launches = prof['conv_kernel_ms'] / args.steps, prof['conv_launches'] / args.steps
    else:
        k_ms_per_step, k_launches = prof['sr_conv_ms'] / args.steps, None
    sr_ms_per_step = prof['sr_conv_ms'] / args.steps
    sr_tflops = SR_GFLOP_PER_FRAME * B / k_ms_per_step                   # GFLOP / ms == TFLOP/s  (algorithmic: the reference's 197.63 GFLOP/frame)
    # dram__bytes_read+write of the 4 conv launches of one step from the committed ncu --set full capture (profiles/r1_ncu_full_summaries_3.md), N = 4
    NCU_TRAFFIC_PER_STEP = (13.2 + 82.0) + (140.0 + 97.9) + (161.9 + 226.8) + (272.8 + 12.2)         # MB
    line = {
        'metric': 'rendered frames/sec @512^2 (64^2 NeRF, 48 samples/ray)', 'value': fps, 'unit': 'frames/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32 render (decoder GEMMs: split-fp16 operands on tcgen05, f32 accumulate, f32-grade results); SR ' + ('f16 operands / f32 accumulate (tcgen05)' if sr_mode == 'tc' else 'f32'),
        'data': 'synthetic',
        'config': config_of(args, {'sr_mode': sr_mode, 'cuda_graph': not args.no_graph, 'l2_policy': f'inputs larger than L2: {P} distinct resident frames/GPU '
                                   f'({P * 25.2:.0f} MB) cycled', 'timing': 'CUDA events on the launch stream, barrier+sync both sides, max over ranks'}),
        'clocks': clocks, 'gpu_launches': int(launches),
        'e2e': {'value': e2e_fps, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': ksteps,
                'host_wall_ms': e2e_wall_ms, 'h2d_GBps': h2d * ksteps / (float(te.item()) * 1e6),
                'note': 'step_host(): pinned host in/out, copies of step i+1 overlap compute of step i; the H2D of the fp32 planes (PCIe) bounds it'},
        'roofline': {'bound': 'tensor', 'achieved': sr_tflops, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
                     'frac': sr_tflops / pk['tf_sustained'],
                     'traffic': (NCU_TRAFFIC_PER_STEP * 1e6 / 4 if (sr_mode == 'tc' and B == 4) else None),
                     'traffic_note': 'bytes per launch: dram read+write averaged over the 4 conv launches of a step (ncu --set full)',
                     'kernel': prof['sr_kernel'], 'launches_per_step': k_launches, 'kernel_ms_per_step': k_ms_per_step,
                     'peak_source': pk['src'] + ' bf16 sustained', 'algorithmic': f'{SR_GFLOP_PER_FRAME} GFLOP/frame x {B} frames/step',
                     'share_of_step': k_ms_per_step / (prof['total_ms'] / args.steps), 'sr_stage_ms_per_step': sr_ms_per_step},
        'stage_ms_per_step': {k: v / args.steps for k, v in prof['stages'].items()},
        'roofline_hbm': hbm,
    }
    if not args.no_cpu_baseline:
        cpu_fps, ts, nthr = time_cpu(1, 5, 1)
        torch.set_num_threads(1)                                   # also report the 1-thread figure (one frame, no warm-up repeat)
        fn1 = cpu_frame_fn(1)
        t1 = time.perf_counter(); fn1(); one_thread_fps = 1.0 / (time.perf_counter() - t1)
        torch.set_num_threads(nthr)
        line['cpu_baseline'] = {'value': cpu_fps, 'unit': 'frames/s', 'cores': nthr, 'kind': 'port', 'one_thread_value': one_thread_fps,
                                'sample': f'{len(ts)} x 1 frame of the same workload (oracle = port of the reference PyTorch CPU path), '
                                          f'{nthr} torch threads of {host_threads()} usable, median'}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
