import sys, os, time, torch, numpy as np
sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT', '/root/repo'))
from brushstroke_engine_b200 import params as P, synthetic
from brushstroke_engine_b200.engine import TriadPaintEngine
cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
eng = TriadPaintEngine(P.init_generator_params(cfg, 0, 0.1), P.init_encoder_params(ecfg, 1, 0.1), 'cuda', mode='bf16')
B = 256
hs = []
for k in range(4):
    p = torch.from_numpy((np.random.RandomState(k).rand(B, 128, 128) > 0.2).astype(np.uint8) * 255).pin_memory()
    z = torch.from_numpy(np.random.RandomState(k).randn(B, 64)).pin_memory()
    pos = torch.randint(0, 4000, (B, 2), dtype=torch.int64).pin_memory()
    hs.append((p, z, pos))
outs = [torch.empty((B, 108, 108, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
sess = eng.batch_session(B, 10)
def loop(mode, n=60):
    torch.cuda.synchronize(); t0 = time.perf_counter(); pending = None
    for i in range(n):
        p, z, pos = hs[i % 4]
        if mode == 'full':
            ev = eng.render_patches_host(p, z, pos, crop_margin=10, out=outs[i & 1], wait=False)[1]
        elif mode == 'graph_only':
            sess._graph.replay(); ev = torch.cuda.Event(); ev.record()
        elif mode == 'run_host':
            sess.run_host(p, z, pos); ev = torch.cuda.Event(); ev.record()
        elif mode == 'run_host_clone':
            t = sess.run_host(p, z, pos).clone(); ev = torch.cuda.Event(); ev.record()
        if pending is not None: pending.synchronize()
        pending = ev
    pending.synchronize(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
with torch.no_grad():
    for m in ('full', 'graph_only', 'run_host', 'run_host_clone'):
        loop(m, 10)
    for rep in range(2):
        print({m: round(loop(m), 4) for m in ('graph_only', 'run_host', 'run_host_clone', 'full')})
