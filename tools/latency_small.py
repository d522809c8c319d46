"""Wall-clock latency of one engine step at small batch sizes (interactive use: one stroke patch at a time)."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import params as P
from brushstroke_engine_b200.generator import Generator
from brushstroke_engine_b200.geo_encoder import GeometryEncoder

cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
gp = P.init_generator_params(cfg, 0, 0.1); ep = P.init_encoder_params(ecfg, 1, 0.1)
G = Generator(gp, cfg, 'cuda', mode='bf16'); enc = GeometryEncoder(ep, ecfg, 'cuda', mode='bf16')
for B in (1, 4, 16, 64):
    z = torch.randn(B, 64, device='cuda', dtype=torch.float64)
    geom = (torch.rand(B, 1, 128, 128, device='cuda') > 0.2).float()
    pos = torch.randint(0, 4000, (B, 2), device='cuda')
    def step():
        ws = G.mapping(z, None).contiguous()
        gf, dests, scales = G.alloc_injection(ws)
        enc.encode_into(geom, dests, scales)
        return G.forward_pre_mapped(ws, gf, positions=pos, noise_mode='const')
    for _ in range(5): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); n = 30
    for _ in range(n): step()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); 
    for _ in range(n): step()
    b.record(); torch.cuda.synchronize()
    print(f'B={B}: wall {wall:.3f} ms/step, device {a.elapsed_time(b) / n:.3f} ms/step')
