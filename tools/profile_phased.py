"""Where does a phased feature-blended canvas spend its time (host set-up vs the three GPU phases)?  usage: profile_phased.py [size]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import params as P, synthetic, stylizer
from brushstroke_engine_b200.engine import TriadPaintEngine, GanBrushOptions

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda')
cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
eng = TriadPaintEngine(P.init_generator_params(cfg, 0, 0.1), P.init_encoder_params(ecfg, 1, 0.1), dev, mode='bf16')
guidance = torch.from_numpy(synthetic.synthetic_guidance(size, size, num_lines=size // 16, seed=0)).to(dev)
opts = GanBrushOptions(); opts.set_style(torch.from_numpy(np.random.RandomState(1).randn(1, 64)).to(dev), '1')
def T():
    torch.cuda.synchronize(); return time.perf_counter()
with torch.no_grad():
    for _ in range(2):
        stylizer.stylize(eng, guidance, opts, crop_margin=10, feature_blending_level=2, to_host=False)
    t0 = T()
    job = stylizer.CanvasJob(eng, guidance, 10, 'all')
    t1 = T()
    waves = stylizer.blending_wavefronts(job.crops_yx, eng.patch_width)
    t2 = T()
    n = len(job.crops_yx)
    zpp = torch.from_numpy(np.random.RandomState(2).randn(n, 64)).to(dev)
    for z in (None, zpp):
        t3 = T()
        c = stylizer._stylize_blended_phased(eng, job, opts, 2, z)
        t4 = T()
        out = job.finish(c, False, False)
        t5 = T()
        print(f'{size}^2, {n} patches, z_per_patch={z is not None}: CanvasJob {1e3*(t1-t0):.2f} ms, wavefronts {1e3*(t2-t1):.2f} ms, phased {1e3*(t4-t3):.2f} ms, finish {1e3*(t5-t4):.2f} ms')
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    c = stylizer._stylize_blended_phased(eng, job, opts, 2, zpp); torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
