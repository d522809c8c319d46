import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import _lib as _L
if os.environ.get("NBE_OLD_LIB"): _L.LIB_PATH = _L.LIB_PATH.replace("libnbe_b200.so", "libnbe_b200_old.so")
from brushstroke_engine_b200.modconv import modulated_conv2d
from brushstroke_engine_b200 import upfirdn2d as U
dev='cuda'
f4 = U.setup_filter([1, 3, 3, 1], device=dev)
for dt in (torch.bfloat16, torch.float16):
  for (B,cin,cout,R,up) in [(256,384,128,16,1),(256,384,128,16,2),(256,384,128,32,1),(256,144,128,8,1),(256,128,128,32,1)]:
    x = torch.randn(B, cin, R, R, device=dev, dtype=dt)
    w = torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(cin * 9)
    st = torch.randn(B, cin, device=dev) * 0.3 + 1
    nz = torch.randn(B, 1, R*up, R*up, device=dev) * 0.1
    fn = lambda: modulated_conv2d(x, w, st, noise=nz, up=up, padding=1, resample_filter=f4, flip_weight=(up == 1))
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(15):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    print(str(dt)[6:], B,cin,cout,R,up, f'median {ts[7]:.4f} min {ts[0]:.4f} max {ts[-1]:.4f}')
