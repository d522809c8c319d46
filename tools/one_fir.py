"""Time one NHWC FIR pass (the second kernel of an up-sampling layer).  usage: one_fir.py N OH [reps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import _lib
from brushstroke_engine_b200.upfirdn2d import setup_filter

N, OH = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = 'cuda'
TP = OH + 2
T = torch.randn(N, TP, TP, 128, device=dev).to(torch.bfloat16)
y = torch.empty(N, OH, OH, 128, dtype=torch.bfloat16, device=dev)
f = setup_filter([1, 3, 3, 1], device=dev).contiguous()
scale = torch.rand(N, 128, device=dev) + 0.5
bias = torch.randn(128, device=dev)
noise = torch.randn(N, OH, OH, device=dev)
st = _lib.stream()
run = lambda: _lib.call('nbe_fir_act_nhwc_bf16', _lib.ptr(T), _lib.ptr(f), _lib.ptr(y), N, OH, OH, 128, OH + 1, OH + 1, 1, 128, TP, TP * TP,
                        128, OH, OH * OH, 4.0, _lib.ptr(scale), _lib.ptr(noise), OH * OH, 0.1, _lib.ptr(bias), 0.2, 2 ** 0.5, 256.0,
                        _lib.ptr(scale), st)
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
gb = (T.numel() + y.numel()) * 2 / 1e9
print(f'fir N={N} OH={OH}: {min(ts):.3f} ms  ({gb / min(ts):.2f} TB/s of T read + y written)')
