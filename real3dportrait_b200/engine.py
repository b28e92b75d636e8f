"""FrameEngine — the per-GPU frame loop of a clip (the caller side of the hot path, inference/real3d_infer.py:480-492):
a batch of tri-planes + cameras + jitter in, 512^2 frames out, one process per GPU; with world > 1 every step ends with
an NCCL all-gather of that step's frames (the only collective of the path: frames are independent)."""
from __future__ import annotations

import contextlib
import os
from typing import Callable, Dict, Optional

import torch

from . import _capi as capi
from .synthesis import RenderHead


@contextlib.contextmanager
def gpu_local_cpus(device_index: int = 0):
    """Run the body on CPUs of the GPU's NUMA node (as far as the process may use them), then restore the affinity.

    Pinned host buffers are placed on the node of the allocating thread; staging buffers on the far socket cost a third of the
    PCIe rate on two-socket hosts.  Wrap only the ALLOCATION: `with gpu_local_cpus(i): buf = torch.empty(..., pin_memory=True)`.
    Anything unexpected (no sysfs entry, empty intersection, no permission) leaves the affinity untouched."""
    old = None
    try:
        pr = torch.cuda.get_device_properties(device_index)
        addr = f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
        txt = open(f'/sys/bus/pci/devices/{addr}/local_cpulist').read().strip()
        local = set()
        for part in txt.split(','):
            if part:
                lo, _, hi = part.partition('-')
                local.update(range(int(lo), int(hi or lo) + 1))
        cur = os.sched_getaffinity(0)
        both = cur & local
        if both and both != cur:
            os.sched_setaffinity(0, both)
            old = cur
    except Exception:
        old = None
    try:
        yield
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except Exception:
                pass


def default_sr_mode() -> str:
    try:
        from . import sr_tc  # noqa: F401
        return 'tc' if sr_tc.available() else 'fp32'
    except ImportError:
        return 'fp32'


class FrameEngine:
    """batch: frames per step; static_styles: the SR styles are constant (Real3D passes ws == 1, img2plane_baseline.py:142) so
    the folded fp16 weights are prepared once per parameter load; use_graph: replay the whole step (≈70 kernels) as ONE CUDA
    graph — inputs are copied into static buffers first, the only per-step host work is the replay."""

    def __init__(self, batch: int = 4, sr_mode: str = 'fp32', device=None, world: int = 1, rank: int = 0, dist=None, hp: Optional[dict] = None,
                 static_styles: bool = True, use_graph: bool = True):
        self.batch, self.world, self.rank, self.dist = batch, world, rank, dist
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.head = RenderHead(hp=hp, sr_mode=sr_mode).to(self.device).eval()
        self.static_styles, self.use_graph = static_styles, use_graph
        self.gathered = torch.empty(world * batch, 3, 512, 512, device=self.device) if world > 1 else None
        self.graph = None
        self.launches_per_step = 0
        self.s_planes = self.s_cams = self.s_u = self.s_out = None
        self.inplace = {}                 # (planes ptr, cameras ptr, jitter ptr) -> (graph, output, the input tensors kept alive): see prepare()
        self._pool = None

    def load_params(self, decoder_params: Dict[str, torch.Tensor], sr_params: Dict[str, torch.Tensor]) -> None:
        sd = {'decoder.' + k: v for k, v in decoder_params.items()}
        sd.update({'superresolution.' + k: v for k, v in sr_params.items()})
        self.head.load_state_dict(sd, strict=True)
        self.graph = None
        self.inplace = {}
        sr = self.head.superresolution
        sr.static_prepared = None
        if self.static_styles and sr.sr_mode == 'tc':
            from . import sr_tc
            ones = torch.ones(1, 3, self.head.hparams['w_dim'], device=self.device)
            with torch.no_grad():
                sr.static_prepared = sr_tc.Prepared(sr, ones)

    @torch.no_grad()
    def _body(self, planes, cameras, u_coarse) -> torch.Tensor:
        return self.head.synthesis(planes, cameras, u_coarse=u_coarse, lean=True)['image']

    def _capture_graph(self, planes, cameras, u_coarse):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):                                       # warm-up outside capture: lazy inits (func attributes, driver entry points)
                self._body(planes, cameras, u_coarse)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()               # all step graphs share one pool: they never run concurrently
        c0 = capi.lib().r3dp_launch_count()
        with torch.cuda.graph(g, pool=self._pool):
            out = self._body(planes, cameras, u_coarse)
        self.launches_per_step = int(capi.lib().r3dp_launch_count() - c0)      # libr3dp kernels inside one replay
        return g, out

    def _capture(self, planes, cameras, u_coarse) -> None:
        self.s_planes, self.s_cams, self.s_u = planes.clone(), cameras.clone(), u_coarse.clone()
        self.graph, self.s_out = self._capture_graph(self.s_planes, self.s_cams, self.s_u)

    @staticmethod
    def _key(planes, cameras, u_coarse):
        return (planes.data_ptr(), cameras.data_ptr(), u_coarse.data_ptr())

    @torch.no_grad()
    def prepare(self, inputs, max_graphs: int = 32) -> int:
        """Zero-copy steps for RESIDENT inputs: capture one step graph per (planes, cameras, u_coarse) triple that reads those very
        buffers, so `step()` on them replays without first copying 100 MB of planes into the static graph inputs.  The tensors are
        kept referenced (their addresses stay valid); refill them in place between steps.  Returns the number of graphs held."""
        if not self.use_graph:
            return 0
        for planes, cameras, u_coarse in inputs:
            k = self._key(planes, cameras, u_coarse)
            if k in self.inplace or len(self.inplace) >= max_graphs or planes.shape[0] != self.batch:
                continue
            assert planes.is_contiguous() and cameras.is_contiguous() and u_coarse.is_contiguous()
            g, out = self._capture_graph(planes, cameras, u_coarse)
            self.inplace[k] = (g, out, (planes, cameras, u_coarse))
        return len(self.inplace)

    @torch.no_grad()
    def step(self, planes: torch.Tensor, cameras: torch.Tensor, u_coarse: Optional[torch.Tensor] = None, u_fine=None) -> torch.Tensor:
        """planes [B,3,32,256,256], cameras [B,25], u_coarse [B,4096,S,1] (drawn here if None) -> frames [B,3,512,512]
        (world == 1) or the all-gathered [world*B,3,512,512] (rank-major).  The returned tensor is reused by the next step."""
        if u_coarse is None:
            S = self.head.rendering_kwargs['depth_resolution']
            u_coarse = torch.rand(planes.shape[0], self.head.neural_rendering_resolution ** 2, S, 1, device=planes.device)
        graphable = self.use_graph and capi.PROF is None and u_fine is None and planes.shape[0] == self.batch
        hit = self.inplace.get(self._key(planes, cameras, u_coarse)) if (graphable and self.inplace) else None
        if hit is not None:
            hit[0].replay()
            out = hit[1]
        elif graphable:
            if self.graph is None:
                self._capture(planes, cameras, u_coarse)
            if planes.data_ptr() != self.s_planes.data_ptr():
                self.s_planes.copy_(planes, non_blocking=True)
            if cameras.data_ptr() != self.s_cams.data_ptr():
                self.s_cams.copy_(cameras, non_blocking=True)
            if u_coarse.data_ptr() != self.s_u.data_ptr():
                self.s_u.copy_(u_coarse, non_blocking=True)
            self.graph.replay()
            out = self.s_out
        else:
            over = {'u_coarse': u_coarse}
            if u_fine is not None:
                over['u_fine'] = u_fine
            out = self.head.synthesis(planes, cameras, **over)['image']
        if self.world > 1:
            if capi.PROF is not None:                                   # profiling pass: serial, so the region time is the collective's own
                with capi.region('allgather'):
                    self.dist.all_gather_into_tensor(self.gathered, out.contiguous())
                return self.gathered
            return self._gather_async(out)
        return out

    def _gather_async(self, out: torch.Tensor) -> torch.Tensor:
        """Frame exchange of step i on a side stream so it runs under the compute of step i+1: the step's frames are copied to one of two
        staging slots (the graph's output buffer is rewritten by the next replay), NCCL all-gathers slot -> gathered[slot] on `comm`.
        The returned tensor is complete once `wait_gather()` (or a device sync) has run; a slot is reused two steps later."""
        if getattr(self, '_ga', None) is None:
            self._ga = {'comm': torch.cuda.Stream(device=self.device), 'k': 0,
                        'stage': [torch.empty_like(out) for _ in range(2)],
                        'dst': [torch.empty(self.world * self.batch, 3, 512, 512, device=self.device) for _ in range(2)],
                        'done': [torch.cuda.Event() for _ in range(2)], 'ready': [torch.cuda.Event() for _ in range(2)]}
            cur0 = torch.cuda.current_stream()
            for e in self._ga['done']:
                e.record(cur0)
        ga = self._ga
        k = ga['k'] & 1
        ga['k'] += 1
        cur = torch.cuda.current_stream()
        cur.wait_event(ga['done'][k])                                   # the collective that last used this slot has finished
        ga['stage'][k].copy_(out, non_blocking=True)
        ga['ready'][k].record(cur)
        with torch.cuda.stream(ga['comm']):
            ga['comm'].wait_event(ga['ready'][k])
            self.dist.all_gather_into_tensor(ga['dst'][k], ga['stage'][k])
            ga['done'][k].record(ga['comm'])
        return ga['dst'][k]

    def wait_gather(self) -> None:
        """Make the current stream wait for every frame exchange issued so far (call before consuming step()'s result when world > 1)."""
        if getattr(self, '_ga', None) is not None:
            torch.cuda.current_stream().wait_stream(self._ga['comm'])

    # ---- host-buffer entry point: H2D / compute / D2H of consecutive steps overlap on three streams -----------------------
    def _host_pipeline(self):
        if getattr(self, '_hp', None) is None:
            dev = self.device
            self._hp = {
                'copy_in': torch.cuda.Stream(device=dev), 'copy_out': torch.cuda.Stream(device=dev), 'k': 0,
                'in_ready': [torch.cuda.Event() for _ in range(2)], 'in_free': [torch.cuda.Event() for _ in range(2)],
                'out_ready': [torch.cuda.Event() for _ in range(2)], 'out_free': [torch.cuda.Event() for _ in range(2)],
                'stage': [None, None], 'out': [None, None],
            }
        return self._hp

    @torch.no_grad()
    def step_host(self, h_planes: torch.Tensor, h_cameras: torch.Tensor, h_u: torch.Tensor, h_out: torch.Tensor) -> None:
        """Same step, from PINNED HOST tensors to a pinned host output (frames of THIS rank), fully asynchronous: the call
        enqueues H2D (copy-in stream) -> step (current stream) -> D2H (copy-out stream) and returns; with two staging slots the
        copy of step i+1 runs under the compute of step i.  Call `sync_host()` before reading `h_out`."""
        hp = self._host_pipeline()
        k = hp['k'] & 1
        hp['k'] += 1
        cur = torch.cuda.current_stream()
        if hp['stage'][k] is None:
            hp['stage'][k] = (torch.empty_like(h_planes, device=self.device), torch.empty_like(h_cameras, device=self.device),
                              torch.empty_like(h_u, device=self.device))
            hp['out'][k] = torch.empty(self.batch, 3, 512, 512, device=self.device)
            self.prepare([hp['stage'][k]])                             # the step reads the staging slot in place (no device-side input copy)
            hp['in_free'][k].record(cur); hp['out_free'][k].record(cur)
        sp, sc, su = hp['stage'][k]
        with torch.cuda.stream(hp['copy_in']):
            hp['copy_in'].wait_event(hp['in_free'][k])                 # the step that last read this slot is done
            sp.copy_(h_planes, non_blocking=True); sc.copy_(h_cameras, non_blocking=True); su.copy_(h_u, non_blocking=True)
            hp['in_ready'][k].record(hp['copy_in'])
        cur.wait_event(hp['in_ready'][k])
        out = self.step(sp, sc, su)
        self.wait_gather()
        hp['in_free'][k].record(cur)
        cur.wait_event(hp['out_free'][k])
        mine = out[self.rank * self.batch:(self.rank + 1) * self.batch] if out.shape[0] > self.batch else out
        hp['out'][k].copy_(mine, non_blocking=True)
        hp['out_ready'][k].record(cur)
        with torch.cuda.stream(hp['copy_out']):
            hp['copy_out'].wait_event(hp['out_ready'][k])
            h_out.copy_(hp['out'][k], non_blocking=True)
            hp['out_free'][k].record(hp['copy_out'])

    def sync_host(self) -> None:
        hp = self._host_pipeline()
        hp['copy_in'].synchronize(); hp['copy_out'].synchronize(); torch.cuda.current_stream().synchronize()

    def static_inputs(self):
        """(planes, cameras, u_coarse) static buffers of the captured graph: a producer may write its outputs straight into them
        and call step() with these very tensors to skip the copy-in."""
        return self.s_planes, self.s_cams, self.s_u

    def profile_steps(self, inputs: Callable[[int], tuple], first: int, steps: int) -> Dict:
        """Re-run `steps` steps eagerly with CUDA events around every stage (on the launching stream)."""
        capi.PROF = capi.Profiler()
        self.step(*inputs(first))                                   # eager warm-up: one-time lazy initialisations stay out of the stage times
        torch.cuda.synchronize()
        capi.PROF = capi.Profiler()
        L = capi.lib()
        tc = self.head.superresolution.sr_mode == 'tc'
        if tc:
            L.r3dp_sr_tc_prof(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            self.step(*inputs(first + i))
        b.record()
        torch.cuda.synchronize()
        stages = capi.PROF.totals()
        capi.PROF = None
        conv_ms, conv_n = None, 0
        if tc:
            import ctypes as C
            ms, n = C.c_float(0), C.c_int(0)
            capi.check(L.r3dp_sr_tc_prof_read(C.byref(ms), C.byref(n)))
            L.r3dp_sr_tc_prof(0)
            conv_ms, conv_n = float(ms.value), int(n.value)
        kernel = 'conv_tc3_kernel<2> (tcgen05 cta_group::2 implicit-GEMM conv)' if self.head.superresolution.sr_mode == 'tc' else 'conv_taps_kernel (fp32 CUDA-core direct conv)'
        return {'stages': stages, 'sr_conv_ms': stages.get('sr_conv', float('nan')), 'total_ms': a.elapsed_time(b), 'sr_kernel': kernel,
                'conv_kernel_ms': conv_ms, 'conv_launches': conv_n}
