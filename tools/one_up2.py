"""One upsample2d (up = 2) launch for profiling / timing.  usage: one_up2.py dtype N C R [reps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import upfirdn2d as U
dt = {'bf16': torch.bfloat16, 'fp16': torch.float16, 'fp32': torch.float32}[sys.argv[1]]
N, C, R = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 9
up = int(os.environ.get('UP', '2'))
f4 = U.setup_filter([1, 3, 3, 1], device='cuda')
xs = [torch.randn(N, C, R, R, device='cuda', dtype=dt) for _ in range(3)]
fn = (lambda x: U.upsample2d(x, f4)) if up == 2 else (lambda x: U.upfirdn2d(x, f4, padding=[1, 1, 1, 1], gain=4.0))
for i in range(3): fn(xs[i])
torch.cuda.synchronize()
ts = []
for i in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); y = fn(xs[i % 3]); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
nb = (xs[0].numel() + y.numel()) * xs[0].element_size()
print(f'up={up} {sys.argv[1]} N={N} C={C} R={R}: {ms:.4f} ms  {nb / ms / 1e6:.0f} GB/s  ({nb / ms / 1e6 / 6534.5:.2f} of the measured HBM peak)')
