"""Time the encoder's first layer alone (block-Toeplitz kernel vs the im2col kernel).  usage: one_enc7.py [N] [reps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
H = W = 128
dev = 'cuda'
lib = _lib.load()
x = (torch.rand(N, H, W, device=dev) > 0.2).float()
w = torch.randn(64, 49, device=dev) / 7
bias = torch.randn(64, device=dev) * 0.1
wt = torch.empty((7, 512, 16), dtype=torch.bfloat16, device=dev)
st = _lib.stream()
_lib.call('nbe_enc_conv7x7_toeplitz_weights', _lib.ptr(w), _lib.ptr(wt), 64, st)
wq = torch.zeros((64, 64), dtype=torch.bfloat16, device=dev); wq[:, :49] = w.to(torch.bfloat16)
nb = lib.nbe_enc_conv7x7_toeplitz_scratch_bytes(N, H, W)
scratch = torch.empty(nb // 2, dtype=torch.bfloat16, device=dev)
y = torch.zeros((N, H + 2, W + 2, 64), dtype=torch.bfloat16, device=dev)
runs = {
    'toeplitz (pad + GEMM + border)': lambda: _lib.call('nbe_enc_conv7x7_toeplitz_bf16', _lib.ptr(x), _lib.ptr(wt), _lib.ptr(bias), _lib.ptr(y), _lib.ptr(scratch), nb, N, H, W, 64, 64, 0.01, 0, st),
    'im2col kernel (no border)': lambda: _lib.call('nbe_enc_conv7x7_tc_bf16', _lib.ptr(x), _lib.ptr(wq), _lib.ptr(bias), _lib.ptr(y), N, H, W, 64, 64, 0.01, 0, st),
}
for name, run in runs.items():
    for _ in range(3): run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    gb = N * (H + 2) * (W + 2) * 64 * 2 / 1e9
    print(f'{name}: {min(ts):.3f} ms  ({gb / min(ts) * 1e3:.0f} GB/s of output)')
