"""First-contact diagnostic for a GPU box: runs every kernel family once and prints max errors instead of
asserting, so that one gpurun call tells us as much as possible.  Each section runs in a subprocess with a
timeout (a trapped / hung kernel must not take the rest down)."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SECTIONS = {
'bias_act': '''
from brushstroke_engine_b200.bias_act import bias_act
x=torch.randn(4,128,32,32); b=torch.randn(128)
for dt in (torch.float32, torch.float16, torch.bfloat16):
    y=bias_act(x.to(dt).cuda(), b.to(dt).cuda(), act='lrelu', gain=2**0.5, clamp=256)
    print('bias_act', dt, md(y, O.bias_act(x.to(dt).float(), b.to(dt).float(), act='lrelu', gain=2**0.5, clamp=256)))
''',
'upfirdn2d': '''
from brushstroke_engine_b200 import upfirdn2d as U
f4=O.setup_filter([1,3,3,1]); x=torch.randn(2,8,33,33)
print('upfirdn gen', md(U.upfirdn2d(x.cuda(), f4.cuda(), padding=[1,1,1,1], gain=4), O.upfirdn2d(x,f4,padding=[1,1,1,1],gain=4.0)))
print('upsample2d', md(U.upsample2d(x.cuda(), f4.cuda()), O.upsample2d(x,f4)))
print('generic   ', md(U.upfirdn2d(x.cuda(), f4.cuda(), up=[3,2], down=[2,1], padding=[2,-1,0,3], gain=1.5), O.upfirdn2d(x,f4,up=[3,2],down=[2,1],padding=[2,-1,0,3],gain=1.5)))
''',
'conv_f32': '''
from brushstroke_engine_b200.conv2d_resample import conv2d_f32
import torch.nn.functional as F
for (cin,cout,H,K,s) in ((128,128,16,3,1),(64,128,16,3,2),(1,64,20,7,1),(128,3,16,1,1)):
    x=torch.randn(2,cin,H,H); w=torch.randn(cout,cin,K,K)/ (cin*K*K)**0.5
    print('conv_f32',cin,cout,H,K,s, md(conv2d_f32(x.cuda(), w.cuda(), padding=K//2, stride=s), F.conv2d(x.double(), w.double(), padding=K//2, stride=s)))
''',
'conv_tc': '''
import numpy as np, torch.nn.functional as F
from brushstroke_engine_b200 import _lib
def pack(x, cs=None):
    N,C,H,W=x.shape; cs=cs or C
    d=torch.zeros((N,H,W,cs),dtype=torch.bfloat16,device='cuda')
    _lib.call('nbe_pack_nhwc_bf16', _lib.ptr(x), _lib.ptr(d), N,C,H,W,cs,0,None,_lib.stream()); return d
for (R,cin,B,valid) in ((16,64,1,0),(16,128,2,0),(128,128,1,0),(4,128,3,0),(8,128,5,1),(32,144,2,1),(64,384,1,0)):
    cout=128; IH=R+2 if valid else R
    x=torch.randn(B,cin,IH,IH); w=torch.randn(cout,cin,3,3)/np.sqrt(cin*9)
    xq=pack(x.cuda()); wq=torch.empty((9,cout,(cin+63)//64*64),dtype=torch.bfloat16,device='cuda')
    _lib.call('nbe_prepare_weights_bf16', _lib.ptr(w.cuda()), _lib.ptr(wq), cout,cin,3,0,_lib.stream())
    y=torch.zeros((B,R,R,cout),dtype=torch.bfloat16,device='cuda')
    _lib.call('nbe_conv_tc_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(y), B,R,R,cin,cin,cout,cout,3,valid, None,None,0,0.0,None,1.0,1.0,-1.0,None,_lib.stream())
    torch.cuda.synchronize()
    ref=F.conv2d(x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double(), padding=0 if valid else 1)
    got=y.permute(0,3,1,2).float()
    print('conv_tc R',R,'cin',cin,'B',B,'valid',valid,'maxdiff',md(got,ref),'refmax',float(ref.abs().max()), flush=True)
''',
'generator': '''
from brushstroke_engine_b200 import params as P
from brushstroke_engine_b200.generator import Generator
from brushstroke_engine_b200.geo_encoder import GeometryEncoder
from conftest import load_golden, t
cfg,ecfg=P.GeneratorConfig(),P.EncoderConfig()
gp=P.init_generator_params(cfg,0,0.1); ep=P.init_encoder_params(ecfg,1,0.1)
g=load_golden('generator')
enc=GeometryEncoder(ep,ecfg,'cuda'); gf=enc.encode(t(g['geom']).cuda())
print('encoder g0', md(gf[0], g['g0']), 'g1', md(gf[1][:, ::8], g['g1_sub']))
for mode in ('fp32','bf16'):
    G=Generator(gp,cfg,'cuda',mode=mode)
    for tag,pos in (('nopos',None),('pos',t(g['positions']).cuda())):
        img,dbg=G(t(g['z']).cuda(),None,gf,positions=pos,return_debug_data=True,return_features=[64],noise_mode='const')
        torch.cuda.synchronize()
        print('generator',mode,tag,'img',md(img,g['img32_'+tag]),'uvs',md(dbg['uvs'],g['uvs32_'+tag]),'feat64',md(dbg['features64'][:, ::16, ::2, ::2], g['feat64_sub_'+tag]), flush=True)
''',
}

PRELUDE = '''
import sys, os, torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))
from oracle import neube_oracle as O
def md(a,b): return float((torch.as_tensor(a).detach().cpu().double()-torch.as_tensor(b).detach().cpu().double()).abs().max())
torch.manual_seed(0)
print(torch.cuda.get_device_name(0), flush=True)
''' % (REPO, REPO)

if __name__ == '__main__':
    names = sys.argv[1:] or list(SECTIONS)
    for name in names:
        print(f'===== {name} =====', flush=True)
        try:
            r = subprocess.run([sys.executable, '-c', PRELUDE + SECTIONS[name]], timeout=300, capture_output=True, text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print(f'[{name}] exit code {r.returncode}\n{r.stderr[-4000:]}')
        except subprocess.TimeoutExpired as e:
            print(f'[{name}] TIMEOUT\n{(e.stdout or b"")[-2000:]}')
