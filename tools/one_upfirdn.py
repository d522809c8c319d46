import sys, os, torch, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import upfirdn2d as U
from brushstroke_engine_b200.bias_act import bias_act
f4 = U.setup_filter([1, 3, 3, 1], device='cuda')
for dt in (torch.float32, torch.bfloat16):
    x = torch.randn(256, 128, 129, 129, device='cuda', dtype=dt)
    for _ in range(3):
        y = U.upfirdn2d(x, f4, padding=[1, 1, 1, 1], gain=4)
    x2 = torch.randn(256, 128, 64, 64, device='cuda', dtype=dt)
    for _ in range(3):
        y2 = U.upsample2d(x2, f4)
    b = torch.randn(128, device='cuda', dtype=dt)
    for _ in range(3):
        y3 = bias_act(x2, b, act='lrelu', gain=math.sqrt(2), clamp=256)
torch.cuda.synchronize()
