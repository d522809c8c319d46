"""A/B of the packed-FP32 (FFMA2) arithmetic against the scalar code it replaces: same inputs through both builds of the
FIR pass (NBE_FIR_SCALAR) and of the staged upfirdn2d kernel (NBE_UPF_SCALAR; NBE_UPF_LDS16 = halfword shared-memory loads) in child processes (the switches are
read once per process), outputs compared bit for bit, times side by side.

    python tools/ab_packed.py            # parent: runs both children, prints the comparison
"""
import os, subprocess, sys, tempfile, math
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def med_ms(fn, iters=12, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def child(out_path):
    from brushstroke_engine_b200 import _lib, params as P, upfirdn2d as U, synthetic
    from brushstroke_engine_b200.engine import GanBrushOptions, TriadPaintEngine
    dev = 'cuda'
    res, times = {}, {}
    f4 = U.setup_filter([1, 3, 3, 1], device=dev)
    gen = torch.Generator().manual_seed(5)
    torch.manual_seed(11)
    with torch.no_grad():
        for dt in (torch.float32, torch.float16, torch.bfloat16):
            for (n, c, h) in ((3, 5, 17), (2, 8, 65), (1, 4, 257)):
                x = torch.randn(n, c, h, h, generator=gen).to(dt).to(dev)
                res[f'upf_{dt}_{h}'] = U.upfirdn2d(x, f4, padding=[1, 1, 1, 1], gain=4).cpu()
                res[f'flt_{dt}_{h}'] = U.upfirdn2d(x[..., :h - 1], f4, padding=[2, 1, 2, 1]).cpu()
        # generator forward, bf16 mode (FIR pass at every up-sampling layer)
        cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
        gp = P.init_generator_params(cfg, 0, 0.1); ep = P.init_encoder_params(ecfg, 1, 0.1)
        import numpy as np
        B = 6
        geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=3 + i, radius=1 + i) for i in range(B)]))
        z = torch.cat([P.style_z_from_seed(100 + i) for i in range(B)])
        pos = torch.tensor([[88 * i, 176 + 40 * i] for i in range(B)])
        engine = TriadPaintEngine(gp, ep, torch.device(dev), mode='bf16')
        opts = GanBrushOptions(); opts.set_style(z.to(dev)); opts.position = pos.to(dev)
        rgba, raw, _ = engine._render_stroke_torch(geom.to(dev), None, opts)
        res['rgba'] = rgba.float().cpu()
        # timings: operator surface at the micro-benchmark's sizes, and the 128^2 FIR pass on its own
        for dt in (torch.float32, torch.bfloat16, torch.float16):
            for R in (32, 64, 128):
                Bx = 128 if R == 128 else 256                                   # the ops keep the reference's INT_MAX element limit
                xs = [torch.randn(Bx, 128, 2 * R + 1, 2 * R + 1, device=dev, dtype=dt) for _ in range(2 if R == 128 else 3)]
                i = [0]
                def run():
                    U.upfirdn2d(xs[i[0] % len(xs)], f4, padding=[1, 1, 1, 1], gain=4); i[0] += 1
                ms = med_ms(run)
                nb = Bx * 128 * ((2 * R + 1) ** 2 + (2 * R) ** 2) * (torch.finfo(dt).bits // 8)
                times[f'upfirdn2d_gen {str(dt)[6:]} R={R}'] = (ms, nb / ms / 1e6)
                del xs
                torch.cuda.empty_cache()
        Bn = 256
        tt = [torch.randn(Bn, 130, 130, 128, device=dev).to(torch.bfloat16) for _ in range(2)]
        y = torch.empty(Bn, 128, 128, 128, device=dev, dtype=torch.bfloat16)
        scale = torch.rand(Bn, 128, device=dev) + 0.5; nscale = torch.rand(Bn, 128, device=dev) + 0.5
        bias = torch.randn(128, device=dev); noise = torch.randn(Bn, 128, 128, device=dev)
        f4f = (f4 * 1.0).contiguous()
        j = [0]
        def fir():
            t = tt[j[0] % 2]; j[0] += 1
            _lib.call('nbe_fir_act_nhwc_bf16', _lib.ptr(t), _lib.ptr(f4f), _lib.ptr(y), Bn, 128, 128, 128, 129, 129, 1,
                      128, 130, 130 * 130, 128, 128, 128 * 128, 4.0,
                      _lib.ptr(scale), _lib.ptr(noise), 128 * 128, 0.3, _lib.ptr(bias), 0.2, math.sqrt(2), 256.0, _lib.ptr(nscale), _lib.stream())
        ms = med_ms(fir)
        times['fir_act_nhwc 128^2 B=256'] = (ms, Bn * (129 * 129 + 128 * 128) * 128 * 2 / ms / 1e6)
        res['fir'] = y[:2].float().cpu()
    torch.save({'res': res, 'times': times}, out_path)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == 'child':
        return child(sys.argv[2])
    outs = {}
    for tag, env in (('packed', {}), ('scalar', {'NBE_FIR_SCALAR': '1', 'NBE_UPF_SCALAR': '1'}), ('lds16', {'NBE_UPF_LDS16': '1'})):
        path = os.path.join(tempfile.gettempdir(), f'ab_{tag}.pt')
        e = dict(os.environ); e.update(env)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), 'child', path], env=e)
        outs[tag] = torch.load(path)
    a = outs['packed']
    bad = []
    for tag in ('scalar', 'lds16'):
        b = outs[tag]
        for k in a['res']:
            if not torch.equal(a['res'][k], b['res'][k]):
                d = (a['res'][k].double() - b['res'][k].double()).abs().max()
                rel = float(d) / max(1e-30, float(b['res'][k].double().abs().max()))
                print(f'DIFF default vs {tag} {k}: max abs {float(d):.3e} (rel {rel:.1e})')
                if rel > 1e-6:
                    bad.append(k)
    print(f"outputs compared: {len(a['res'])} x 2; beyond 1e-6 relative: {len(bad)}")
    print('| case | default ms | GB/s | scalar FP32 ms | GB/s | halfword LDS ms | GB/s |')
    print('|---|---:|---:|---:|---:|---:|---:|')
    for k in a['times']:
        (m1, g1), (m0, g0), (m2, g2) = a['times'][k], outs['scalar']['times'][k], outs['lds16']['times'][k]
        print(f'| {k} | {m1:.4f} | {g1:.0f} | {m0:.4f} | {g0:.0f} | {m2:.4f} | {g2:.0f} |')
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
