"""Per-kernel timing of one bf16 generator forward at batch B (CUDA events around every C-ABI call)."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import _lib, params as P
from brushstroke_engine_b200.generator import Generator
from brushstroke_engine_b200.geo_encoder import GeometryEncoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mode = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
gp = P.init_generator_params(cfg, 0, 0.1); ep = P.init_encoder_params(ecfg, 1, 0.1)
G = Generator(gp, cfg, 'cuda', mode=mode)
enc = GeometryEncoder(ep, ecfg, 'cuda', mode=mode)
z = torch.randn(B, 64, device='cuda', dtype=torch.float64)
geom = (torch.rand(B, 1, 128, 128, device='cuda') > 0.2).float()
pos = torch.randint(0, 4000, (B, 2), device='cuda')

records = []
orig_call = _lib.call
def timed_call(name, *args):
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); orig_call(name, *args); e.record()
    records.append((name, args, s, e))
def run(timed):
    _lib.call = timed_call if timed else orig_call
    import brushstroke_engine_b200.generator as gm, brushstroke_engine_b200.conv2d_resample as cr, brushstroke_engine_b200.modconv as mc, brushstroke_engine_b200.upfirdn2d as up, brushstroke_engine_b200.bias_act as ba
    import brushstroke_engine_b200.geo_encoder as ge
    for m in (gm, cr, mc, up, ba, ge):
        m._lib.call = _lib.call
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
    t0.record()
    if mode == 'bf16':
        ws = G.mapping(z, None).contiguous()
        gf, dests, scales = G.alloc_injection(ws)
        enc.encode_into(geom, dests, scales)
        t1.record()
        img, dbg = G.forward_pre_mapped(ws, gf, positions=pos, return_debug_data=True, noise_mode='const')
    else:
        gf = enc.encode(geom)
        t1.record()
        img, dbg = G(z, None, gf, positions=pos, return_debug_data=True, noise_mode='const')
    t2.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1), t1.elapsed_time(t2)

for _ in range(3): run(False)
te, tg = run(False)
print(f'B={B} mode={mode}: encoder {te:.3f} ms, generator {tg:.3f} ms -> {B/(te+tg)*1e3:.0f} patches/s (gen only {B/tg*1e3:.0f})')
records.clear(); run(True)
tot = 0
for name, args, s, e in records:
    ms = s.elapsed_time(e); tot += ms
    shape = ''
    if name == 'nbe_conv_tc_bf16_ex': shape = f'N={args[3]} R={args[4]} Cin={args[6]} Cout={args[8]} stride={args[12]}'
    elif name == 'nbe_conv_tc_bf16': shape = f'N={args[3]} R={args[4]} Cin={args[6]} valid={args[11]}'
    elif name in ('nbe_conv3x3_flat_bf16',): shape = f'N={args[3]} R={args[4]} Cin={args[6]}'
    elif name == 'nbe_convT3x3s2_flat_bf16': shape = f'N={args[3]} H={args[4]} Cin={args[6]}'
    elif name == 'nbe_fir_act_nhwc_bf16': shape = f'N={args[3]} OH={args[4]} C={args[6]}'
    elif name == 'nbe_conv2d_f32': shape = f'N={args[3]} Cin={args[4]} H={args[5]} Cout={args[7]} K={args[8]} s={args[10]}'
    elif name == 'nbe_upsample2x_nhwc_bf16': shape = f'N={args[4]} H={args[5]} C={args[7]}'
    if ms > 0.02: print(f'  {name:28s} {ms:8.3f} ms  {shape}')
print(f'  sum of timed kernels {tot:.3f} ms over {len(records)} launches')
