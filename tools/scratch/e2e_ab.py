import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from real3dportrait_b200 import engine, synthetic as syn
dev = torch.device('cuda', 0)
B = 4
planes = syn.make_planes(B, seed=100).to(dev); cams = syn.make_cameras(B, seed=200).to(dev); u = syn.make_jitter(B, 4096, 48, 0, seed=300)[0].to(dev)
print('affinity', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], '...')
try:
    pr = torch.cuda.get_device_properties(0)
    addr = f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
    print('gpu', addr, 'local_cpulist', open(f'/sys/bus/pci/devices/{addr}/local_cpulist').read().strip(), 'numa_node', open(f'/sys/bus/pci/devices/{addr}/numa_node').read().strip())
except Exception as e:
    print('sysfs:', e)
def run(numa, inplace):
    eng = engine.FrameEngine(batch=B, sr_mode='tc', device=dev)
    eng.load_params(syn.make_decoder_params(seed=4), syn.make_sr_params(seed=5))
    if not inplace:
        eng.prepare = lambda inputs, max_graphs=32: 0
    ctx = engine.gpu_local_cpus(0) if numa else __import__('contextlib').nullcontext()
    with ctx:
        hp, hc, hu = planes.cpu().pin_memory(), cams.cpu().pin_memory(), u.cpu().pin_memory()
        ho = torch.empty(B, 3, 512, 512).pin_memory()
    for _ in range(4): eng.step_host(hp, hc, hu, ho)
    eng.sync_host()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40): eng.step_host(hp, hc, hu, ho)
    eng.sync_host(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 40
    # raw H2D rate of the same buffer
    d = torch.empty_like(planes)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): d.copy_(hp, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    print(f'numa={numa} inplace={inplace}: {ms:.3f} ms/step = {B / ms * 1e3:.0f} frames/s; raw H2D of the planes {hp.numel() * 4 / dt / 1e9:.1f} GB/s')
for numa in (0, 1):
    for inplace in (0, 1):
        run(numa, inplace)
