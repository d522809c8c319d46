"""Randomised shape sweep of the flat tensor-core kernels against float64 torch references (debugging aid; the fixed
parametrisations live in tests/test_conv_flat_gpu.py).  usage: fuzz_flat.py [n_cases] [seed]"""
import sys, os
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import _lib

DEV = 'cuda'
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


def prep_w(w):
    cout, cin = w.shape[:2]
    wq = torch.empty((9, cout, (cin + 63) // 64 * 64), dtype=torch.bfloat16, device=DEV)
    _lib.call('nbe_prepare_weights_bf16', _lib.ptr(w.to(DEV).contiguous()), _lib.ptr(wq), cout, cin, 3, 0, _lib.stream())
    return wq


def pitched(x, pitch, cs):
    N, C, H, W = x.shape
    buf = torch.zeros((N, H, pitch, cs), dtype=torch.bfloat16, device=DEV)
    buf[:, :, :W, :C] = x.to(DEV).permute(0, 2, 3, 1).to(torch.bfloat16)
    return buf


def md(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


worst = 0.0
for case in range(n_cases):
    kind = ['conv', 'convT', 's2', 'fir'][case % 4]
    g = torch.Generator().manual_seed(int(rng.randint(1 << 30)))
    B = int(rng.randint(1, 8))
    if kind == 'conv':
        R, cin, valid, gap = int(rng.randint(3, 140)), int(rng.randint(1, 50)) * 8, int(rng.randint(2)), int(rng.randint(1, 4))
        IH = R + 2 if valid else R
        x = torch.randn(B, cin, IH, IH, generator=g); w = torch.randn(128, cin, 3, 3, generator=g) / np.sqrt(cin * 9)
        pitch = IH + gap
        xq, wq = pitched(x, pitch, cin), prep_w(w)
        y = torch.zeros((B, R, R + 1, 128), dtype=torch.bfloat16, device=DEV)
        _lib.call('nbe_conv3x3_flat_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(y), B, R, R, cin, cin, pitch, valid, 128, 128, R + 1, R * (R + 1),
                  None, None, 0, 0.0, None, 1.0, 1.0, -1.0, None, _lib.stream())
        ref = F.conv2d(x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double(), padding=0 if valid else 1)
        got = y[:, :, :R, :].permute(0, 3, 1, 2).float()
        desc = f'conv R={R} cin={cin} B={B} valid={valid} gap={gap}'
    elif kind == 'convT':
        H, cin, gap = int(rng.randint(2, 70)), int(rng.randint(1, 50)) * 8, int(rng.randint(1, 4))
        x = torch.randn(B, cin, H, H, generator=g); w = torch.randn(128, cin, 3, 3, generator=g) / np.sqrt(cin * 9)
        xq, wq = pitched(x, H + gap, cin), prep_w(w)
        TP = 2 * H + 2
        t = torch.zeros((B, TP, TP, 128), dtype=torch.bfloat16, device=DEV)
        _lib.call('nbe_convT3x3s2_flat_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(t), B, H, H, cin, cin, H + gap, 128, 128, TP, TP * TP, None, _lib.stream())
        ref = F.conv_transpose2d(x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double().transpose(0, 1), stride=2)
        got = t[:, :2 * H + 1, :2 * H + 1, :].permute(0, 3, 1, 2).float()
        desc = f'convT H={H} cin={cin} B={B} gap={gap}'
    elif kind == 's2':
        H, cin, cout = int(rng.randint(2, 70)) * 2, int(rng.choice([64, 128, 192, 256])), int(rng.choice([128, 256]))
        xp = torch.randn(B, cin, H + 2, H + 2, generator=g); w = torch.randn(cout, cin, 3, 3, generator=g) / np.sqrt(cin * 9)
        xq = xp.to(DEV).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16); wq = prep_w(w)
        OH = H // 2
        y = torch.zeros((B, OH, OH, cout), dtype=torch.bfloat16, device=DEV)
        _lib.call('nbe_conv3x3s2_flat_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(y), B, H, H, cin, cout, cout, OH, OH * OH, None, 1.0, 1.0, -1.0, None, _lib.stream())
        ref = F.conv2d(xp.to(torch.bfloat16).double(), w.to(torch.bfloat16).double(), stride=2)
        got = y.permute(0, 3, 1, 2).float()
        desc = f's2 H={H} cin={cin} cout={cout} B={B}'
    else:
        OH = int(rng.randint(2, 140)); C = 128
        TH = OH + 1
        tt = torch.randn(B, C, TH, TH, generator=g)
        f = torch.tensor([1., 3., 3., 1.]); f2 = torch.outer(f, f); f2 = (f2 / f2.sum()).to(DEV)
        tq = torch.zeros((B, TH + 1, TH + 1, C), dtype=torch.bfloat16, device=DEV)
        tq[:, :TH, :TH] = tt.to(DEV).permute(0, 2, 3, 1).to(torch.bfloat16)
        y = torch.zeros((B, OH, OH + 1, C), dtype=torch.bfloat16, device=DEV)
        _lib.call('nbe_fir_act_nhwc_bf16', _lib.ptr(tq), _lib.ptr(f2), _lib.ptr(y), B, OH, OH, C, TH, TH, 1, C, TH + 1, (TH + 1) * (TH + 1),
                  C, OH + 1, OH * (OH + 1), 4.0, None, None, 0, 0.0, None, 1.0, 1.0, -1.0, None, _lib.stream())
        xpad = F.pad(tt.to(torch.bfloat16).double(), (1, 1, 1, 1))
        ref = F.conv2d(xpad, (f2.cpu().double() * 4).flip(0, 1)[None, None].expand(C, 1, 4, 4), groups=C)
        got = y[:, :, :OH, :].permute(0, 3, 1, 2).float()
        desc = f'fir OH={OH} B={B}'
    torch.cuda.synchronize()
    err = md(got, ref) / max(float(ref.abs().max()), 1.0)
    worst = max(worst, err)
    flag = 'OK ' if err < 1e-2 else 'BAD'
    if flag == 'BAD' or case < 8:
        print(f'{flag} {desc}: rel err {err:.2e}')
print(f'{n_cases} cases, worst relative error {worst:.2e}')
sys.exit(0 if worst < 1e-2 else 1)
