"""One operator-surface modulated_conv2d call (pack -> tcgen05 conv -> unpack) for profiling.  usage: one_modconv.py B C R up [reps]"""
import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200.modconv import modulated_conv2d
from brushstroke_engine_b200 import upfirdn2d as U
B, C, R, up = (int(a) for a in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 7
dev = 'cuda'
f4 = U.setup_filter([1, 3, 3, 1], device=dev)
x = torch.randn(B, C, R, R, device=dev, dtype=torch.bfloat16)
w = torch.randn(C, C, 3, 3, device=dev) / math.sqrt(C * 9)
st = torch.randn(B, C, device=dev) * 0.3 + 1
nz = torch.randn(B, 1, R * up, R * up, device=dev) * 0.1
fn = lambda: modulated_conv2d(x, w, st, noise=nz, up=up, padding=1, resample_filter=f4, flip_weight=(up == 1))
for _ in range(3): fn()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(f'modulated_conv2d bf16 B={B} C={C} R={R} up={up}: {sorted(ts)[len(ts)//2]:.4f} ms')
