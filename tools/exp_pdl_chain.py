"""Does programmatic dependent launch shorten a chain of small persistent launches?  Eager and CUDA-graph, run with and without NBE_NO_PDL."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import _lib
N, H, Cin, K = int(sys.argv[1]), int(sys.argv[2]), 128, 40
dev = 'cuda'
P = H + 1
x = torch.zeros(N, H, P, Cin, dtype=torch.bfloat16, device=dev)
x[:, :, :H] = torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16)
w = torch.randn(128, Cin, 3, 3, device=dev)
wq = torch.zeros(9 * 128 * Cin, dtype=torch.bfloat16, device=dev)
_lib.call('nbe_prepare_weights_bf16', _lib.ptr(w), _lib.ptr(wq), 128, Cin, 3, 0, _lib.stream())
y = torch.zeros(N, H, P, 128, dtype=torch.bfloat16, device=dev)
def chain():
    st = _lib.stream()
    for i in range(K):
        a, b = (x, y) if i % 2 == 0 else (y, x)
        _lib.call('nbe_conv3x3_flat_bf16', _lib.ptr(a), _lib.ptr(wq), _lib.ptr(b), N, H, H, Cin, Cin, P, 0, 128, 128, P, H * P,
                  None, None, 0, 0.0, None, 1.0, 1.0, -1.0, None, st)
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
t_eager = timeit(chain)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    chain()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    chain()
t_graph = timeit(g.replay)
print(f'PDL={"off" if os.environ.get("NBE_NO_PDL") else "on"} N={N} H={H}: chain of {K} conv launches: eager {t_eager*1e3/K:.2f} us/launch, graph {t_graph*1e3/K:.2f} us/launch')
