"""Latency of one stroke patch through the CUDA-graph interactive session and of small batch-step graphs, for A/B runs with
NBE_NO_PDL=1 (programmatic dependent launch off)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import params as P, synthetic
from brushstroke_engine_b200.engine import TriadPaintEngine, GanBrushOptions, BatchSession
dev = 'cuda'
cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
eng = TriadPaintEngine(P.init_generator_params(cfg, 0, 0.1), P.init_encoder_params(ecfg, 1, 0.1), dev, mode='bf16')
opts = GanBrushOptions(); opts.set_style(P.style_z_from_seed(21).to(dev), '21')
tag = 'PDL off' if os.environ.get('NBE_NO_PDL') else 'PDL on '
with torch.no_grad():
    sess = eng.interactive_session(opts, crop_margin=0)
    patch = np.ascontiguousarray(((1.0 - synthetic.synthetic_patch(128, seed=4)[0, 0]) * 255).astype(np.uint8)[:, :, None])
    for _ in range(20): sess.render_stroke(patch, (5, 7))
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); sess.render_stroke(patch, (5, 7)); ts.append(time.perf_counter() - t0)
    print(f'{tag}: interactive stroke patch (host in, host out, one graph replay): median {np.median(ts) * 1e3:.4f} ms')
    for B in (1, 8, 24):
        bs = BatchSession(eng, B, 10)
        geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=i) for i in range(B)])).to(dev)
        z = torch.cat([P.style_z_from_seed(i) for i in range(B)]).to(dev); pos = torch.zeros((B, 2), dtype=torch.int64, device=dev)
        for _ in range(10): bs.run(geom, z, pos)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100): bs.run(geom, z, pos)
        b.record(); torch.cuda.synchronize()
        print(f'{tag}: batch step graph, B = {B}: {a.elapsed_time(b) / 100:.4f} ms per replay (device)')
