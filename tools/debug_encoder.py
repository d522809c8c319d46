import sys, os, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brushstroke_engine_b200 import params as P, synthetic
from brushstroke_engine_b200.geo_encoder import GeometryEncoder
import numpy as np
cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
ep = P.init_encoder_params(ecfg, 1, 0.1)
enc = GeometryEncoder(ep, ecfg, 'cuda', mode='bf16')
geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=3), synthetic.synthetic_patch(128, seed=4, radius=3)])).cuda()
B = 2
dests = []
for r, R in ((0, 16), (1, 32)):
    dests.append((torch.zeros((B, R, R, ecfg.feature_channels(r)), dtype=torch.bfloat16, device='cuda'), 0))
enc.encode_into(geom, dests)
torch.cuda.synchronize()
ws = enc._ws[(B, 128)]
# reference per layer (fp32, folded weights)
x = geom.float()
refs = []
for i, (w, b, stride, pad, up) in enumerate(enc._layers):
    if up:
        x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)
        refs.append(('up', x))
    x = F.leaky_relu(F.conv2d(F.pad(x, (pad,)*4, mode='reflect'), w, b, stride=stride), 0.01)
    refs.append((f'L{i}', x))
def cmp(name, got_nhwc_padded, ref, padded=True):
    C = ref.shape[1]
    g = got_nhwc_padded[..., :C].permute(0, 3, 1, 2).float()
    if padded:
        r = F.pad(ref, (1, 1, 1, 1), mode='reflect')
    else:
        r = ref
    d = (g - r).abs()
    inner = d[..., 1:-1, 1:-1].max().item() if padded else d.max().item()
    print(f'{name}: ref max {r.abs().max().item():.4f}  max err all {d.max().item():.4f}  interior {inner:.4f}  padchan max {got_nhwc_padded[..., C:].abs().max().item() if got_nhwc_padded.shape[3] > C else 0}')
wi = 0
ri = 0
for i, (w, b, stride, pad, up) in enumerate(enc._layers):
    if up:
        name, r = refs[ri]; ri += 1
        cmp(f'layer{i} bilinear+pad', ws[wi], r); wi += 1
    name, r = refs[ri]; ri += 1
    if i == 5:
        cmp('layer5 -> g0 dest', dests[0][0], r, padded=False)
    elif i == 6:
        cmp('layer6 -> g1 dest', dests[1][0], r, padded=False)
    else:
        cmp(f'layer{i} out', ws[wi], r)
    wi += 1
