"""BASELINE configs[2]: upfirdn2d / bias_act / modulated_conv2d micro-benchmark sweep against the HBM and tensor rooflines.

    python tools/microbench.py [--quick] [--out profiles/microbench.json]

Every case runs through the public operator shims (C ABI underneath); times are CUDA-event medians over `iters`
launches after warm-up, with the two tensors of each case rotated over enough copies to exceed the 126 MB L2.
"""
import argparse, json, os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import _lib
from brushstroke_engine_b200.bias_act import bias_act
from brushstroke_engine_b200 import upfirdn2d as U
from brushstroke_engine_b200.modconv import modulated_conv2d

PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}
L2 = 126e6


def timeit(fn, n_variants, iters=12, warmup=3):
    for i in range(warmup):
        fn(i % n_variants)
    torch.cuda.synchronize()
    evs = []
    for i in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(i % n_variants); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def variants(nbytes):
    return max(2, int(math.ceil(2 * L2 / max(nbytes, 1))) if nbytes < 2 * L2 else 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    dev = 'cuda'
    f4 = U.setup_filter([1, 3, 3, 1], device=dev)
    rows = []
    Rs = (8, 16, 32, 64, 128)
    Cs = (128, 144, 256, 384, 512) if not args.quick else (128,)
    Bs = (1, 16, 256) if not args.quick else (256,)
    dts = (torch.float32, torch.bfloat16, torch.float16)
    with torch.no_grad():
        for dt in dts:
            es = torch.finfo(dt).bits // 8
            for B in Bs:
                for C in Cs:
                    for R in Rs:
                        if B * C * (2 * R + 1) ** 2 > 2 ** 31 - 1:       # the ops keep the reference's INT_MAX element limit
                            continue
                        # ---- bias_act (lrelu, gain sqrt2, clamp 256) on [B,C,R,R]
                        nb = 2 * B * C * R * R * es
                        nv = min(variants(nb), 16)
                        xs = [torch.randn(B, C, R, R, device=dev, dtype=dt) for _ in range(nv)]
                        b = torch.randn(C, device=dev, dtype=dt)
                        ms = timeit(lambda i: bias_act(xs[i], b, act='lrelu', gain=math.sqrt(2), clamp=256), nv)
                        rows.append(dict(op='bias_act', dtype=str(dt).split('.')[-1], B=B, C=C, R=R, ms=ms, gbs=nb / ms / 1e6))
                        del xs
                        # ---- upfirdn2d generator case (2R+1)^2 -> (2R)^2, pad 1, gain 4
                        nb = B * C * ((2 * R + 1) ** 2 + (2 * R) ** 2) * es
                        nv = min(variants(nb), 8)
                        xs = [torch.randn(B, C, 2 * R + 1, 2 * R + 1, device=dev, dtype=dt) for _ in range(nv)]
                        ms = timeit(lambda i: U.upfirdn2d(xs[i], f4, padding=[1, 1, 1, 1], gain=4), nv)
                        rows.append(dict(op='upfirdn2d_gen', dtype=str(dt).split('.')[-1], B=B, C=C, R=R, ms=ms, gbs=nb / ms / 1e6))
                        del xs
                        # ---- upsample2d x2: R^2 -> (2R)^2
                        nb = B * C * (R * R + (2 * R) ** 2) * es
                        nv = min(variants(nb), 8)
                        xs = [torch.randn(B, C, R, R, device=dev, dtype=dt) for _ in range(nv)]
                        ms = timeit(lambda i: U.upsample2d(xs[i], f4), nv)
                        rows.append(dict(op='upsample2d', dtype=str(dt).split('.')[-1], B=B, C=C, R=R, ms=ms, gbs=nb / ms / 1e6))
                        del xs
                        torch.cuda.empty_cache()
    # ---- modulated_conv2d through the operator surface (NCHW in / out): bf16 / fp16 on the tcgen05 kernels behind
    #      nbe_modulated_conv2d (pack -> flat conv | transposed conv + FIR -> unpack), against the tensor roofline; beside it
    #      the same convolution on cuDNN (bf16, channels_last; conv2d / conv_transpose2d only, without modulation, FIR, noise
    #      or layout changes) -- the incumbent kernel bar of SURVEY 2.3 K5/K6
    import torch.nn.functional as F
    mrows = []
    tf_peak = PEAKS.get('bf16_tflops', 1633.0)
    Cm = (128, 144, 256, 384, 512) if not args.quick else (128, 384)
    Bm = (1, 16, 256) if not args.quick else (256,)
    with torch.no_grad():
        for dt in (torch.bfloat16, torch.float16, torch.float32):
            for B in Bm:
                for C in Cm:
                    cin, cout = (C, 128) if C in (144, 384) else (C, C)      # 144 / 384: the generator's concatenated inputs
                    for R in Rs:
                        for up in (1, 2):
                            flops = 2.0 * cout * cin * 9 * R * R * B          # algorithmic: per INPUT pixel for up = 2
                            if dt == torch.float32 and flops > 2e12:
                                continue                                     # the FP32 parity kernel is not a throughput path
                            OR = R * up
                            if B * max(cin, cout) * OR * OR > 2 ** 31 - 1 or B * cout * OR * OR * 2 * 4 > 60e9:
                                continue
                            try:
                                x = torch.randn(B, cin, R, R, device=dev, dtype=dt)
                                w = torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(cin * 9)
                                st = torch.randn(B, cin, device=dev) * 0.3 + 1
                                nz = torch.randn(B, 1, OR, OR, device=dev) * 0.1
                                ms = timeit(lambda i: modulated_conv2d(x, w, st, noise=nz, up=up, padding=1, resample_filter=f4,
                                                                       flip_weight=(up == 1)), 1, iters=10)
                                row = dict(op='modulated_conv2d', dtype=str(dt).split('.')[-1], B=B, Cin=cin, Cout=cout, R=R, up=up, ms=ms,
                                           tflops=flops / ms / 1e9, frac_tensor=flops / ms / 1e9 / tf_peak)
                                if dt == torch.bfloat16:
                                    xc = x.to(memory_format=torch.channels_last)
                                    wc = w.to(dt).to(memory_format=torch.channels_last)
                                    if up == 1:
                                        row['cudnn_ms'] = timeit(lambda i: F.conv2d(xc, wc, padding=1), 1, iters=10)
                                    else:
                                        wt = w.to(dt).transpose(0, 1).contiguous().to(memory_format=torch.channels_last)
                                        row['cudnn_ms'] = timeit(lambda i: F.conv_transpose2d(xc, wt, stride=2), 1, iters=10)
                                mrows.append(row)
                                del x, w, st, nz
                            except RuntimeError as ex:
                                mrows.append(dict(op='modulated_conv2d', dtype=str(dt).split('.')[-1], B=B, Cin=cin, Cout=cout, R=R, up=up,
                                                  error=str(ex)[:120]))
                            torch.cuda.empty_cache()
    hb = PEAKS['hbm_gbs']
    for r in rows:
        r['frac_hbm'] = r['gbs'] / hb
    print(f'# microbench vs measured HBM peak {hb:.0f} GB/s')
    print('| op | dtype | B | C | R | ms | GB/s | frac of HBM |')
    print('|---|---|---:|---:|---:|---:|---:|---:|')
    for r in rows:
        print(f"| {r['op']} | {r['dtype']} | {r['B']} | {r['C']} | {r['R']} | {r['ms']:.4f} | {r['gbs']:.0f} | {r['frac_hbm']:.2f} |")
    print(f'\n# modulated_conv2d (operator surface, NCHW in/out incl. pack / unpack) vs measured BF16 burst peak {tf_peak:.0f} TFLOP/s; '
          f'cuDNN column = the bare bf16 channels_last conv2d / conv_transpose2d of the same shape')
    print('| dtype | B | Cin | Cout | R | up | ms | TFLOP/s (algorithmic) | frac of tensor peak | cuDNN conv ms |')
    print('|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|')
    for r in mrows:
        if 'error' in r:
            print(f"| {r['dtype']} | {r['B']} | {r['Cin']} | {r['Cout']} | {r['R']} | {r['up']} | error: {r['error']} | | | |")
        else:
            cud = f"{r['cudnn_ms']:.4f}" if 'cudnn_ms' in r else ''
            print(f"| {r['dtype']} | {r['B']} | {r['Cin']} | {r['Cout']} | {r['R']} | {r['up']} | {r['ms']:.4f} | {r['tflops']:.1f} | {r['frac_tensor']:.3f} | {cud} |")
    if args.out:
        json.dump(rows + mrows, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
