import os, sys, torch, time
sys.path.insert(0, '/root/repo' if os.path.exists('/root/repo/bench.py') else os.getcwd())
import real3dportrait_b200 as r3
from real3dportrait_b200 import synthetic as syn
dev='cuda'
N=4
planes=syn.make_planes(N).to(dev); cam=syn.make_cameras(N).to(dev)
c2w,K=syn.split_camera(cam)
dec=r3.OSGDecoder(32,{'decoder_lr_mul':1,'decoder_output_dim':32}); dec.load_state_dict(syn.make_decoder_params()); dec=dec.to(dev)
rs=r3.RaySampler(); ren=r3.ImportanceRenderer()
ro,rd=rs(c2w,K,64)
u_c,_=syn.make_jitter(N,4096,48)
opts=dict(syn.RENDERING_OPTIONS); opts['u_coarse']=u_c.to(dev)
pcl=r3.planes_to_channels_last(planes)
with torch.no_grad():
    out=ren(pcl,dec,ro,rd,opts)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    for _ in range(3): ren(pcl,dec,ro,rd,opts)
    e0.record()
    for _ in range(20): ren(pcl,dec,ro,rd,opts)
    e1.record(); torch.cuda.synchronize()
print(os.environ.get('R3DP_MLP','tc(default)'), 'render call ms', e0.elapsed_time(e1)/20, 'rgb mean', float(out[0].mean()), 'finite', bool(torch.isfinite(out[0]).all()))
torch.save([o.cpu() for o in out[:3]], f"/tmp/render_{os.environ.get('R3DP_MLP','tc')}.pt")
