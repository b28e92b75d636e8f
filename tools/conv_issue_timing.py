"""Profiling tool: where do the MMA-issuing warp and the epilogue warps of conv_tc3_kernel spend their cycles?

    python -m real3dportrait_b200.build --debug-timing
    R3DP_LIB=real3dportrait_b200/lib/libr3dp_b200_dbg.so python tools/conv_issue_timing.py

The debug library carries clock64()/globaltimer probes (-DR3DP_TC_DEBUG_TIMING=1); with the production library the counters stay 0.
Prints, per SR conv launch: barrier-wait / issue clocks of the leader MMA warp, the sections of two epilogue warps, and the SM clock the
kernel actually ran at (clock64 / globaltimer)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import real3dportrait_b200 as r3
from real3dportrait_b200 import _capi, synthetic as syn

dev = 'cuda'
sr = r3.SuperresolutionHybrid8XDC(channels=32, img_resolution=512, sr_num_fp16_res=0, sr_antialias=True, sr_mode='tc')
sr.load_state_dict(syn.make_sr_params(seed=5)); sr = sr.to(dev).eval()
N = 4
x = (torch.rand(N, 32, 64, 64, device=dev) * 2 - 1)
ws = torch.ones(N, 14, 512, device=dev)
buf = torch.zeros(32, 24, dtype=torch.int64, device=dev)
L = _capi.lib()
with torch.no_grad():
    for _ in range(2):
        sr(x[:, :3].contiguous(), x, ws, noise_mode='none')
    torch.cuda.synchronize()
    L.r3dp_sr_tc_debug_buffer(_capi.ptr(buf, torch.int64))
    sr(x[:, :3].contiguous(), x, ws, noise_mode='none')
    torch.cuda.synchronize()
    L.r3dp_sr_tc_debug_buffer(None)
names = ['block0.conv0 (composed up)', 'block0.conv1 + ToRGB', 'block1.conv0 (4-phase up)', 'block1.conv1 + ToRGB']
for i, row in enumerate(buf.tolist()):
    acc, a, b, issue, tot, n, ns, tiles = row[:8]
    if n == 0:
        continue
    print(f'launch {i} {names[i] if i < 4 else ""}: {n} leader warps, {tiles / n:.1f} tiles each, loop {ns / n / 1e3:.1f} us, '
          f'{tot / n:.0f} clk -> {tot / max(ns, 1):.3f} GHz effective SM clock')
    for name, v in (('wait accumulator free', acc), ('wait A strips', a), ('wait B taps', b), ('issue MMAs + commit', issue)):
        print(f'    {name:24s} {v / n:10.0f} clk ({100.0 * v / tot:5.1f} %)   {v / max(tiles, 1):8.0f} clk/tile')
    print(f'    {"loop total":24s} {tot / n:10.0f} clk            {tot / max(tiles, 1):8.0f} clk/tile')
    for w, off in (('epilogue warp 2 (cols 0-63)', 8), ('epilogue warp 6 (cols 64-127)', 16)):
        full, ld, math_, xchg, fin, etot, etiles, pre = row[off:off + 8]
        print(f'    {w}: per tile  wait acc full {full / etiles:7.0f} | tcgen05.ld {ld / etiles:7.0f} | bias/act/store/rgb {math_ / etiles:7.0f} | '
              f'rgb exchange {xchg / etiles:7.0f} | image store {fin / etiles:7.0f} | skip prefetch {pre / etiles:7.0f} | total {etot / etiles:7.0f} clk')
