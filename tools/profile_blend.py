"""How much of a feature-blended canvas is GPU time and how much is launch overhead (debugging aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import params as P, synthetic, stylizer
from brushstroke_engine_b200.engine import TriadPaintEngine, GanBrushOptions

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device('cuda')
cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
eng = TriadPaintEngine(P.init_generator_params(cfg, 0, 0.1), P.init_encoder_params(ecfg, 1, 0.1), dev, mode='bf16')
guidance = torch.from_numpy(synthetic.synthetic_guidance(size, size, num_lines=size // 16, seed=0)).to(dev)
opts = GanBrushOptions(); opts.set_style(torch.from_numpy(np.random.RandomState(1).randn(1, 64)).to(dev), '1')
with torch.no_grad():
    for _ in range(2):
        stylizer.stylize(eng, guidance, opts, crop_margin=10, feature_blending_level=2, to_host=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stylizer.stylize(eng, guidance, opts, crop_margin=10, feature_blending_level=2, to_host=False)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        stylizer.stylize(eng, guidance, opts, crop_margin=10, feature_blending_level=2, to_host=False)
        torch.cuda.synchronize()
    gpu_us = sum(e.device_time_total for e in prof.key_averages())
    top = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]
print(f'canvas {size}^2 level 2: wall {wall:.1f} ms, sum of GPU kernel time {gpu_us / 1e3:.1f} ms')
for e in top:
    print(f'  {e.device_time_total / 1e3:8.2f} ms  x{e.count:5d}  {e.key[:90]}')
