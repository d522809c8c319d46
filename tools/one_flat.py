"""Time one flat transposed conv / conv3x3 launch (debugging aid).  usage: one_flat.py convT|conv N H Cin [reps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import _lib

kind, N, H, Cin = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dev = 'cuda'
P = H + 1
x = torch.zeros(N, H, P, Cin, dtype=torch.bfloat16, device=dev)
x[:, :, :H] = torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16)
w = torch.randn(128, Cin, 3, 3, device=dev)
Cin_pad = (Cin + 63) // 64 * 64
wq = torch.zeros(9 * 128 * Cin_pad, dtype=torch.bfloat16, device=dev)
st = _lib.stream()
_lib.call('nbe_prepare_weights_bf16', _lib.ptr(w), _lib.ptr(wq), 128, Cin, 3, 0, st)
if kind == 'convT':
    TP = 2 * H + 2
    out = torch.zeros(N, TP, TP, 128, dtype=torch.bfloat16, device=dev)
    run = lambda: _lib.call('nbe_convT3x3s2_flat_bf16', _lib.ptr(x), _lib.ptr(wq), _lib.ptr(out), N, H, H, Cin, Cin, P, 128, 128, TP, TP * TP, None, st)
else:
    out = torch.zeros(N, H, P, 128, dtype=torch.bfloat16, device=dev)
    run = lambda: _lib.call('nbe_conv3x3_flat_bf16', _lib.ptr(x), _lib.ptr(wq), _lib.ptr(out), N, H, H, Cin, Cin, P, 0, 128, 128, P, H * P,
                            None, None, 0, 0.0, None, 1.0, 1.0, -1.0, None, st)
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
flops = 2.0 * N * H * H * 9 * 128 * Cin
print(f'{kind} N={N} H={H} Cin={Cin}: {min(ts):.3f} ms  ({flops / min(ts) / 1e9:.0f} TFLOP/s algorithmic)')
