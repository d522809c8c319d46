"""Per-phase wall-clock breakdown of a multi-GPU stylize() call (debugging aid; run under torchrun)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import params as P, synthetic, stylizer
from brushstroke_engine_b200.engine import TriadPaintEngine, GanBrushOptions

rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
dev = torch.device('cuda', lr)
cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
eng = TriadPaintEngine(P.init_generator_params(cfg, 0, 0.1), P.init_encoder_params(ecfg, 1, 0.1), dev, mode='bf16')
guidance = synthetic.synthetic_guidance(4096, 4096, num_lines=256, seed=0)
d_guidance = torch.from_numpy(guidance).to(dev)
opts = GanBrushOptions(); opts.set_style(torch.from_numpy(np.random.RandomState(1).randn(1, 64)).to(dev))
marks = []
_orig = {}
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); marks.append((name, (time.perf_counter() - t0) * 1e3))
        return r
    setattr(obj, name, g)
for name in ('gather', 'owner_map', 'place', 'finish'):
    wrap(stylizer.CanvasJob, name)
wrap(stylizer, 'CanvasJob') if False else None
orig_init = stylizer.CanvasJob.__init__
def timed_init(self, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); orig_init(self, *a, **k); torch.cuda.synchronize()
    marks.append(('CanvasJob.__init__', (time.perf_counter() - t0) * 1e3))
stylizer.CanvasJob.__init__ = timed_init
orig_rt = eng.render_tiles
def timed_rt(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = orig_rt(*a, **k); torch.cuda.synchronize()
    marks.append(('render_tiles', (time.perf_counter() - t0) * 1e3)); return r
eng.render_tiles = timed_rt
if world > 1:
    orig_gather = dist.gather
    def timed_gather(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = orig_gather(*a, **k); torch.cuda.synchronize()
        marks.append(('dist.gather', (time.perf_counter() - t0) * 1e3)); return r
    dist.gather = timed_gather
with torch.no_grad():
    for rep in range(4):
        marks.clear()
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = stylizer.stylize(eng, d_guidance, opts, crop_margin=10, batch_size=256, to_host=False)
        torch.cuda.synchronize(); total = (time.perf_counter() - t0) * 1e3
if rank == 0:
    agg = {}
    for k, v in marks: agg[k] = agg.get(k, 0.0) + v
    print(f'world={world} total {total:.2f} ms (with per-phase syncs); phases: ' + ', '.join(f'{k} {v:.2f}' for k, v in agg.items()) + f'; unaccounted {total - sum(agg.values()):.2f}')
