"""The incumbent GPU execution of the NeuBE hot path, restated on plain ATen / cuDNN calls -- MEASUREMENT INFRASTRUCTURE
(bench.py's `gpu_incumbent` leg and tools/microbench.py); never imported by the product, and not the parity oracle.

What it stands for: on torch >= 2 the reference's two CUDA plugins fail
their post-build import (SG2/torch_utils/custom_ops.py:107-111), `_init()` swallows the error and every `bias_act` /
`upfirdn2d` call runs its `impl='ref'` torch implementation on the GPU (bias_act.py:93-123, upfirdn2d.py:168-208); all
convolutions go to cuDNN through `conv2d_gradfix` (conv2d_gradfix.py:38,43).  So "the reference as shipped on this box" is:
per layer ~8 ATen glue kernels + cuDNN conv / conv_transpose + a depthwise cuDNN conv for the FIR.  /root/reference does not
travel to the GPU box, hence this restatement; every function cites what it follows.  Precision policy as the pickled
generators: blocks with resolution >= 16 in fp16 + channels_last (networks_modified.py:71,108; num_fp16_res = 4), un-fused
modulation when fp16 and batch > 1 (networks.py:636-638), fused grouped conv otherwise; `force_fp32` runs everything in
float32 (cuDNN TF32 as torch defaults allow).  `lowp` may be torch.bfloat16 to give cuDNN the same operand type as ours.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


def bias_act_ref(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """_bias_act_ref, bias_act.py:93-123 (one ATen kernel per step, as the reference executes it)."""
    def_gain = {'linear': 1.0, 'lrelu': SQRT2, 'tanh': 1.0}[act]
    gain = def_gain if gain is None else gain
    if b is not None:
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    if act == 'lrelu':
        x = F.leaky_relu(x, 0.2 if alpha is None else alpha)
    elif act == 'tanh':
        x = torch.tanh(x)
    if gain != 1:
        x = x * gain
    if clamp is not None and clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def upfirdn2d_ref(x, f, up=1, padding=(0, 0, 0, 0), gain=1.0):
    """_upfirdn2d_ref, upfirdn2d.py:168-208 (down = 1, 2-D filter, flip_filter = False)."""
    B, C, H, W = x.shape
    px0, px1, py0, py1 = padding
    if up > 1:
        x = x.reshape([B, C, H, 1, W, 1])
        x = F.pad(x, [0, up - 1, 0, 0, 0, up - 1])
        x = x.reshape([B, C, H * up, W * up])
    x = F.pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0): x.shape[2] - max(-py1, 0), max(-px0, 0): x.shape[3] - max(-px1, 0)]
    f = (f * gain).to(x.dtype).flip([0, 1])
    f = f[None, None].repeat([C, 1, 1, 1])
    return F.conv2d(x, f, groups=C)


def conv2d_resample(x, w, f=None, up=1, padding=1, groups=1, flip_weight=True):
    """conv2d_resample.py:59-154 for the generator's two cases: up = 1 (plain conv, :144-147) and up = 2
    (conv_transpose2d stride 2 + upfirdn2d, :124-142)."""
    kh = w.shape[2]
    if up == 1:
        if not flip_weight:
            w = w.flip([2, 3])
        return F.conv2d(x, w, padding=padding, groups=groups)
    fw = f.shape[-1]
    px0 = padding + (fw + up - 1) // 2 - (kh - 1)
    px1 = padding + (fw - up) // 2 - (kh - up)
    oc, icg = w.shape[0], w.shape[1]
    if groups == 1:
        wt = w.transpose(0, 1)
    else:
        wt = w.reshape(groups, oc // groups, icg, kh, kh).transpose(1, 2).reshape(groups * icg, oc // groups, kh, kh)
    if flip_weight:                                                 # wrapper gets `not flip_weight` and flips when that is False
        wt = wt.flip([2, 3])
    x = F.conv_transpose2d(x, wt, stride=up, padding=0, groups=groups)
    return upfirdn2d_ref(x, f, padding=(px0, px1, px0, px1), gain=up ** 2)


def modulated_conv2d(x, weight, styles, noise=None, up=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    """networks.py:31-88, both branches."""
    B = x.shape[0]
    O, I, kh, kw = weight.shape
    if x.dtype == torch.float16 and demodulate:
        weight = weight * (1 / np.sqrt(I * kh * kw) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
    w = dcoefs = None
    if demodulate or fused_modconv:
        w = weight.unsqueeze(0) * styles.reshape(B, 1, -1, 1, 1)
    if demodulate:
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    if demodulate and fused_modconv:
        w = w * dcoefs.reshape(B, -1, 1, 1, 1)
    if not fused_modconv:
        x = x * styles.to(x.dtype).reshape(B, -1, 1, 1)
        x = conv2d_resample(x, weight.to(x.dtype), f=resample_filter, up=up, padding=padding, flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = torch.addcmul(noise.to(x.dtype), x, dcoefs.to(x.dtype).reshape(B, -1, 1, 1))      # fma.py:15
        elif demodulate:
            x = x * dcoefs.to(x.dtype).reshape(B, -1, 1, 1)
        elif noise is not None:
            x = x.add_(noise.to(x.dtype))
        return x
    x = x.reshape(1, -1, *x.shape[2:])
    w = w.reshape(-1, I, kh, kw)
    x = conv2d_resample(x, w.to(x.dtype), f=resample_filter, up=up, padding=padding, groups=B, flip_weight=flip_weight)
    x = x.reshape(B, -1, *x.shape[2:])
    if noise is not None:
        x = x.add_(noise)
    return x


def fc(x, weight, bias, lr=1.0, act='linear'):
    """FullyConnectedLayer.forward, networks.py:109-122."""
    w = weight.to(x.dtype) * (lr / math.sqrt(weight.shape[1]))
    b = bias.to(x.dtype) * lr if lr != 1 else bias.to(x.dtype)
    if act == 'linear':
        return torch.addmm(b.unsqueeze(0), x, w.t())
    return bias_act_ref(x.matmul(w.t()), b, act=act)


class Incumbent:
    """Weights on the device once (as the reference's nn.Module holds them); forward = encoder + generator + composite."""

    def __init__(self, gp, ep, cfg, ecfg, device, lowp=torch.float16, force_fp32=False):
        self.cfg, self.ecfg, self.dev = cfg, ecfg, device
        self.p = {k: v.to(device) for k, v in gp.items()}
        self.e = {k: v.to(device) for k, v in ep.items()}
        self.lowp, self.force_fp32 = lowp, force_fp32
        f = torch.tensor([1., 3., 3., 1.], device=device)
        f = f.ger(f)
        self.filter = f / f.sum()                                     # setup_filter([1,3,3,1]), upfirdn2d.py:72-116
        self.grids = {}
        for res in cfg.block_resolutions:                            # create_sampling_grid, networks.py:295-299
            lin = torch.linspace(0, 1, res)
            xv, yv = torch.meshgrid(lin, lin, indexing='ij')
            self.grids[res] = torch.stack([xv, yv], dim=2).unsqueeze(0).to(device)

    # ---- generator ---------------------------------------------------------------------------------------------
    def mapping(self, z):
        p, cfg = self.p, self.cfg
        x = z.to(torch.float32)
        x = x * (x.square().mean(dim=1, keepdim=True) + 1e-8).rsqrt()                      # networks.py:24-26
        for i in range(cfg.mapping_layers):
            x = fc(x, p[f'mapping.fc{i}.weight'], p[f'mapping.fc{i}.bias'], lr=cfg.mapping_lr_multiplier, act='lrelu')
        return x.unsqueeze(1).repeat([1, cfg.num_ws, 1])

    def layer(self, prefix, x, w, up, res, norm_pos, fused):
        """SynthesisLayer.forward, networks.py:362-391 (noise_mode='const')."""
        p = self.p
        styles = fc(w, p[f'{prefix}.affine.weight'], p[f'{prefix}.affine.bias'])
        noise = p[f'{prefix}.noise_const']
        if norm_pos is not None:
            grid = (self.grids[res] + norm_pos.unsqueeze(1).unsqueeze(1)) % 1                   # networks.py:377-381
            noise = F.grid_sample(noise.unsqueeze(0).unsqueeze(0).expand(x.shape[0], -1, -1, -1), grid * 2 - 1,
                                  padding_mode='reflection', align_corners=True)
        noise = noise * p[f'{prefix}.noise_strength']
        x = modulated_conv2d(x, p[f'{prefix}.weight'], styles, noise=noise, up=up, padding=1, resample_filter=self.filter,
                             flip_weight=(up == 1), fused_modconv=fused)
        clamp = self.cfg.conv_clamp
        return bias_act_ref(x, p[f'{prefix}.bias'].to(x.dtype), act='lrelu', gain=SQRT2, clamp=clamp)

    def synthesis(self, ws, geom_feature, positions):
        p, cfg = self.p, self.cfg
        B = ws.shape[0]
        norm_pos = None
        if positions is not None:
            norm_pos = (positions % cfg.img_resolution) / (cfg.img_resolution - 1)          # networks_modified.py:351-353
        x, w_idx, geo_idx = None, 0, 0
        for res in cfg.block_resolutions:
            use_fp16 = (not self.force_fp32) and res >= 16
            dtype = self.lowp if use_fp16 else torch.float32
            mf = torch.channels_last if use_fp16 else torch.contiguous_format
            fused = dtype == torch.float32 or B == 1                                           # networks.py:636-638
            name = f'synthesis.b{res}'
            if res == 4:
                x = p['synthesis.b4.const'].to(dtype=dtype, memory_format=mf).unsqueeze(0).repeat([B, 1, 1, 1])
                x = self.layer(f'{name}.conv1', x, ws[:, w_idx], 1, res, norm_pos, fused)
                nconv = 1
            else:
                x = x.to(dtype=dtype, memory_format=mf)
                x = self.layer(f'{name}.conv0', x, ws[:, w_idx], 2, res, norm_pos, fused)
                x = self.layer(f'{name}.conv1', x, ws[:, w_idx + 1], 1, res, norm_pos, fused)
                nconv = 2
            if res == cfg.block_resolutions[-1]:
                # ToRGBColorTriadLayer.forward, networks.py:451-485
                k = f'{name}.torgb'
                scaled = fc(ws[:, w_idx + nconv], p[f'{k}.affine.weight'], p[f'{k}.affine.bias'])
                colors = bias_act_ref(scaled[:, :9], p[f'{k}.color_bias'], act='tanh').reshape(-1, 3, 3)
                styles = scaled[:, 9:] * (1 / math.sqrt(x.shape[1]))
                t = modulated_conv2d(x, p[f'{k}.weight'], styles, demodulate=False, fused_modconv=fused)
                t = bias_act_ref(t, p[f'{k}.bias'].to(t.dtype), clamp=cfg.conv_clamp)
                uvs = torch.softmax(t[:, :3], dim=1)
                img = torch.sum(uvs.unsqueeze(1) * colors.unsqueeze(-1).unsqueeze(-1), dim=2).to(torch.float32)
            if res in cfg.geom_feature_resolutions:
                x = torch.cat([x, geom_feature[geo_idx].to(x.dtype)], dim=1)                 # networks_modified.py:219
                geo_idx += 1
            w_idx += nconv
        return img, uvs, colors

    # ---- encoder (simple_autoencoder.py:95-126,155-199,251-297) ----------------------------------------------------
    def _single(self, prefix, x, stride, pad):
        e = self.e
        x = F.conv2d(F.pad(x, (pad,) * 4, mode='reflect'), e[f'{prefix}.0.weight'], e[f'{prefix}.0.bias'], stride=stride)
        x = F.batch_norm(x, e[f'{prefix}.1.running_mean'], e[f'{prefix}.1.running_var'], e[f'{prefix}.1.weight'],
                         e[f'{prefix}.1.bias'], False, 0.0, self.ecfg.bn_eps)
        return F.leaky_relu(x, self.ecfg.neg_slope)

    def encode(self, geom):
        ecfg = self.ecfg
        x = self._single('encoder.model.0.conv', geom, 1, 3)
        idx = 1
        for _ in ecfg.down_filters:
            x = self._single(f'encoder.model.{idx}.conv', x, 2, 1)
            idx += 1
        for _ in ecfg.post_filters:
            x = self._single(f'encoder.model.{idx}.conv', x, 1, 1)
            idx += 1
        out = [x]
        for i in range(max(ecfg.encode_resolutions)):
            x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)
            x = self._single(f'decoder.model.{i}.conv.conv', x, 1, 1)
            out.append(x)
        return [out[r] for r in ecfg.encode_resolutions]

    # ---- engine tail (brush.py:763-792, 369-377) ---------------------------------------------------------------------
    def render_tiles(self, geom, z, positions, crop_margin=10):
        gf = self.encode(geom)
        ws = self.mapping(z)
        img, uvs, colors = self.synthesis(ws, gf, positions)
        uvs = uvs.to(torch.float32)
        c01 = (colors + 1) / 2
        rgb = torch.sum(uvs.unsqueeze(1) * c01.unsqueeze(-1).unsqueeze(-1), dim=2)
        alpha = uvs[:, :1] + uvs[:, 1:2]
        rgba = torch.cat([rgb, alpha], dim=1)
        m = crop_margin
        tile = rgba[:, :, m:rgba.shape[2] - m, m:rgba.shape[3] - m].permute(0, 2, 3, 1)
        return (tile * 255).clip(0, 255).to(torch.uint8), rgba
