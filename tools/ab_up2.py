"""upsample2d (up = 2) through the persistent staged kernel with 6 / 8 CTAs per SM and with one item per CTA: parity of
every variant against the oracle and against each other, and times at the micro-benchmark's size."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brushstroke_engine_b200 import upfirdn2d as U
from oracle import neube_oracle as O

VARIANTS = {'persistent6': {'NBE_UPF_ONESHOT': '0'}, 'persistent8': {'NBE_UPF_ONESHOT': '0', 'NBE_UPF_PER_SM': '8'},
            'persistent12': {'NBE_UPF_ONESHOT': '0', 'NBE_UPF_PER_SM': '12'}, 'oneshot': {'NBE_UPF_ONESHOT': '1'}}


def use(env):
    for k in ('NBE_UPF_PER_SM', 'NBE_UPF_ONESHOT'):
        os.environ.pop(k, None)
    os.environ.update(env)


def main():
    dev = 'cuda'
    f4 = U.setup_filter([1, 3, 3, 1], device=dev)
    gen = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for dt, tol in ((torch.float32, 3e-6), (torch.bfloat16, 2e-2), (torch.float16, 2e-3)):
            for shape in ((2, 3, 33, 33), (1, 2, 128, 128), (3, 5, 7, 9)):
                x = torch.randn(*shape, generator=gen).to(dt)
                ref = O.upsample2d(x.float(), f4.cpu())
                outs = {}
                for name, env in VARIANTS.items():
                    use(env)
                    outs[name] = U.upsample2d(x.to(dev), f4).float().cpu()
                    err = float((outs[name] - ref).abs().max())
                    assert err < tol * 4, (name, dt, shape, err)
                    assert torch.equal(outs[name], outs['persistent6']), (name, dt, shape)
        print('parity: every variant matches the oracle and the default bit for bit')
        for dt in (torch.float32, torch.bfloat16):
            R = 64
            xs = [torch.randn(256, 128, R, R, device=dev, dtype=dt) for _ in range(3)]
            nb = 256 * 128 * (R * R + 4 * R * R) * (torch.finfo(dt).bits // 8)
            for name, env in VARIANTS.items():
                use(env)
                for i in range(3):
                    U.upsample2d(xs[i % 3], f4)
                torch.cuda.synchronize()
                ts = []
                for i in range(9):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); U.upsample2d(xs[i % 3], f4); b.record(); torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
                ms = sorted(ts)[4]
                print(f'upsample2d {str(dt)[6:]} R={R} {name}: {ms:.4f} ms  {nb / ms / 1e6:.0f} GB/s')
            del xs


if __name__ == '__main__':
    main()
