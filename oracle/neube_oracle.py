"""CPU oracle for the NeuBE generator-forward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (torch-CPU / numpy, float32 or
float64) of the algorithm on the path SURVEY.md section 8 scopes.  It is the checker
for the CUDA product in ``brushstroke_engine_b200``; nothing in the product
imports it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.

Parity pin: the reference ships no tests and no golden vectors, so
the oracle is pinned against *outputs of the reference itself*, produced by
importing the unmodified reference from ``/root/reference`` inside the build
container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``); ``tests/
test_oracle_golden.py`` re-checks the oracle against those fixtures on every
run, with no access to ``/root/reference``.

Third-party arithmetic: convolution, grid_sample, bilinear interpolation and
softmax live in PyTorch/ATen (reference pins torch 1.8.1, container has 2.11);
the oracle calls the same ATen CPU ops for those primitives
(reference call sites: torch_utils/ops/conv2d_gradfix.py:38,43;
training/networks.py:118,377,470; forger/experimental/autoenc/simple_autoencoder.py:98,117).

All functions take a flat parameter bundle (see brushstroke_engine_b200/params.py)
instead of ``nn.Module`` objects.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Bundle = Dict[str, torch.Tensor]

SQRT2 = math.sqrt(2.0)

# --------------------------------------------------------------------------------------
# bias_act -- thirdparty/stylegan2_ada_pytorch/torch_utils/ops/bias_act.py:23-33,93-123
# --------------------------------------------------------------------------------------

#            name        def_alpha def_gain  cuda_idx
ACTIVATIONS = {
    'linear':   (0.0, 1.0,   1),
    'relu':     (0.0, SQRT2, 2),
    'lrelu':    (0.2, SQRT2, 3),
    'tanh':     (0.0, 1.0,   4),
    'sigmoid':  (0.0, 1.0,   5),
    'elu':      (0.0, 1.0,   6),
    'selu':     (0.0, 1.0,   7),
    'softplus': (0.0, 1.0,   8),
    'swish':    (0.0, SQRT2, 9),
}


def _act(x: torch.Tensor, act: str, alpha: float) -> torch.Tensor:
    if act == 'linear':
        return x
    if act == 'relu':
        return torch.clamp_min(x, 0)
    if act == 'lrelu':
        return torch.where(x > 0, x, x * alpha)
    if act == 'tanh':
        return torch.tanh(x)
    if act == 'sigmoid':
        return torch.sigmoid(x)
    if act == 'elu':
        return torch.where(x >= 0, x, torch.expm1(x))
    if act == 'selu':
        scale, a = 1.0507009873554804934193349852946, 1.6732632423543772848170429916717
        return scale * torch.where(x >= 0, x, a * torch.expm1(x))
    if act == 'softplus':
        return F.softplus(x)
    if act == 'swish':
        return torch.sigmoid(x) * x
    raise ValueError(act)


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """``y = clamp(act(x + b[dim]) * gain, +-clamp)`` (bias_act.py:93-123)."""
    def_alpha, def_gain, _ = ACTIVATIONS[act]
    alpha = float(def_alpha if alpha is None else alpha)
    gain = float(def_gain if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    if b is not None:
        assert b.ndim == 1 and b.shape[0] == x.shape[dim]
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    x = _act(x, act, alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


# --------------------------------------------------------------------------------------
# upfirdn2d -- torch_utils/ops/upfirdn2d.py:72-116,168-208 ; upfirdn2d.cpp:32-33
# --------------------------------------------------------------------------------------

def setup_filter(f=(1, 3, 3, 1), normalize=True, flip_filter=False, gain=1.0, separable=None) -> torch.Tensor:
    """upfirdn2d.py:72-116."""
    f = torch.as_tensor(1 if f is None else f, dtype=torch.float32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    return f * (gain ** (f.ndim / 2))


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _pad4(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    padding = list(padding)
    if len(padding) == 2:
        padding = [padding[0], padding[0], padding[1], padding[1]]
    return tuple(padding)


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1.0):
    """The reference's formulation (upfirdn2d.py:168-208): zero-stuff, pad / crop, depthwise correlation with the
    flipped filter (ATen conv2d, the same primitive the reference calls), decimate.  Cross-checked against the
    tap-by-tap closed form :func:`upfirdn2d_closed_form` in tests/test_oracle_golden.py."""
    assert x.ndim == 4
    upx, upy = _pair(up)
    downx, downy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    N, C, H, W = x.shape
    u = x.new_zeros(N, C, H * upy, W * upx)
    u[:, :, ::upy, ::upx] = x
    u = F.pad(u, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    u = u[:, :, max(-py0, 0): u.shape[2] - max(-py1, 0), max(-px0, 0): u.shape[3] - max(-px1, 0)]
    f = (f * (gain ** (f.ndim / 2))).to(x.dtype)
    if not flip_filter:
        f = f.flip(list(range(f.ndim)))
    k = f[None, None].repeat([C, 1] + [1] * f.ndim)
    if f.ndim == 2:
        u = F.conv2d(u, k, groups=C)
    else:
        u = F.conv2d(u, k.unsqueeze(2), groups=C)
        u = F.conv2d(u, k.unsqueeze(3), groups=C)
    return u[:, :, ::downy, ::downx]


def upfirdn2d_closed_form(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1.0):
    """Closed form of SURVEY.md appendix C.2 evaluated tap by tap (what the CUDA kernels implement):

    ``y[oy,ox] = gain * sum_{a,b} ft[a,b] * xhat[oy*downy + a - pady0, ox*downx + b - padx0]``
    with ``xhat`` the zero-stuffed input and ``ft`` = f flipped unless flip_filter.
    1-D filters are applied as two passes with sqrt(gain) each (upfirdn2d.py:239-240).
    """
    assert x.ndim == 4
    upx, upy = _pair(up)
    downx, downy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    if f.ndim == 1:
        g = math.sqrt(gain)
        x = upfirdn2d_closed_form(x, f[None, :], up=(upx, 1), down=(downx, 1), padding=(px0, px1, 0, 0), flip_filter=flip_filter, gain=g)
        return upfirdn2d_closed_form(x, f[:, None], up=(1, upy), down=(1, downy), padding=(0, 0, py0, py1), flip_filter=flip_filter, gain=g)
    N, C, H, W = x.shape
    fh, fw = f.shape
    ft = (f if flip_filter else f.flip([0, 1])).to(x.dtype) * gain
    # zero-stuffed, padded / cropped canvas
    UH, UW = H * upy, W * upx
    PH, PW = UH + py0 + py1, UW + px0 + px1
    canvas = x.new_zeros(N, C, max(PH, 0), max(PW, 0))
    # xhat index (u, v) lives at canvas (u + py0, v + px0)
    us = torch.arange(H) * upy + py0
    vs = torch.arange(W) * upx + px0
    uk = (us >= 0) & (us < PH)
    vk = (vs >= 0) & (vs < PW)
    if uk.any() and vk.any():
        canvas[:, :, us[uk][:, None], vs[vk][None, :]] = x[:, :, uk][:, :, :, vk]
    OH = (PH - fh + downy) // downy
    OW = (PW - fw + downx) // downx
    assert OH >= 1 and OW >= 1
    y = x.new_zeros(N, C, OH, OW)
    for a in range(fh):
        for b in range(fw):
            y += ft[a, b] * canvas[:, :, a: a + (OH - 1) * downy + 1: downy, b: b + (OW - 1) * downx + 1: downx]
    return y


def filter2d(x, f, padding=0, flip_filter=False, gain=1.0):
    px0, px1, py0, py1 = _pad4(padding)
    fh, fw = (f.shape[0], f.shape[-1]) if f is not None else (1, 1)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1.0):
    """upfirdn2d.py:308-343."""
    upx, upy = _pair(up)
    px0, px1, py0, py1 = _pad4(padding)
    fh, fw = (f.shape[0], f.shape[-1]) if f is not None else (1, 1)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1.0):
    """upfirdn2d.py:347-382."""
    dx, dy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    fh, fw = (f.shape[0], f.shape[-1]) if f is not None else (1, 1)
    p = [px0 + (fw - dx + 1) // 2, px1 + (fw - dx) // 2, py0 + (fh - dy + 1) // 2, py1 + (fh - dy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


# --------------------------------------------------------------------------------------
# conv2d_resample -- torch_utils/ops/conv2d_resample.py:59-154
# --------------------------------------------------------------------------------------

def _resample_padding(f, up, down, padding):
    """conv2d_resample.py:94-104."""
    fw = fh = 1
    if f is not None:
        fh, fw = f.shape[0], f.shape[-1]
    px0, px1, py0, py1 = _pad4(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2
    return px0, px1, py0, py1


def conv2d_resample_definition(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """Mathematical definition = the reference's generic fallback (conv2d_resample.py:149-154): FIR-upsample, convolve,
    FIR-downsample.  This FIR-first order is what the CUDA product executes for up-sampling layers."""
    px0, px1, py0, py1 = _resample_padding(f, up, down, padding)
    x = upfirdn2d(x, f if up > 1 else None, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if not flip_weight:
        w = w.flip([2, 3])
    x = F.conv2d(x, w.to(x.dtype), groups=groups)
    if down > 1:
        x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
    return x


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """conv2d_resample.py:59-154 with the reference's own choice of path for the cases the generator hits:
    up > 1 -> transposed strided convolution followed by the FIR (:124-142); up == down == 1 with symmetric
    padding -> plain convolution (:144-147); anything else -> the definition above.  (The two up-sampling orders are
    algebraically identical -- full convolutions commute -- and tests assert they agree.)"""
    kh, kw = w.shape[2], w.shape[3]
    px0, px1, py0, py1 = _resample_padding(f, up, down, padding)
    if up > 1 and groups == 1 and down == 1:
        wt = w.transpose(0, 1)
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        # _conv2d_wrapper(transpose=True, flip_weight=(not flip_weight)): flips when flip_weight is True
        wt = wt.flip([2, 3]) if flip_weight else wt
        x = F.conv_transpose2d(x, wt.to(x.dtype), stride=up, padding=[pyt, pxt])
        return upfirdn2d(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
    if up == 1 and down == 1 and px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return F.conv2d(x, (w if flip_weight else w.flip([2, 3])).to(x.dtype), padding=[py0, px0], groups=groups)
    return conv2d_resample_definition(x, w, f, up, down, padding, groups, flip_weight, flip_filter)


# --------------------------------------------------------------------------------------
# modulated_conv2d -- training/networks.py:31-88
# --------------------------------------------------------------------------------------

def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None,
                     demodulate=True, flip_weight=True):
    """``y[b,o] = d[b,o] * sum_i (W[o,i] (*) (s[b,i] x[b,i])) + noise[b]`` with
    ``d = rsqrt(sum_{i,k} (W s)^2 + 1e-8)`` (networks.py:55-75; un-fused order, which
    the reference states is equal to the fused grouped conv of :78-88)."""
    B = x.shape[0]
    s = styles.to(x.dtype)
    y = conv2d_resample(x * s.reshape(B, -1, 1, 1), weight.to(x.dtype), f=resample_filter, up=up, down=down,
                        padding=padding, flip_weight=flip_weight)
    if demodulate:
        wsq = weight.to(x.dtype).square().sum(dim=[2, 3])                     # [O, I]
        d = (s.square() @ wsq.t() + 1e-8).rsqrt()                              # [B, O]
        y = y * d.reshape(B, -1, 1, 1)
    if noise is not None:
        y = y + noise.to(x.dtype)
    return y


# --------------------------------------------------------------------------------------
# Layers -- training/networks.py:24-26,109-122,255-290,295-299,362-391,451-485
# --------------------------------------------------------------------------------------

def fully_connected(x, weight, bias, activation='linear', lr_multiplier=1.0):
    """networks.py:109-122."""
    w = weight.to(x.dtype) * (lr_multiplier / math.sqrt(weight.shape[1]))
    b = bias.to(x.dtype) * lr_multiplier if bias is not None else None
    y = x @ w.t()
    return bias_act(y, b, act=activation)


def mapping_network(p: Bundle, z, num_ws: int, num_layers: int = 4, lr_multiplier: float = 0.01,
                    truncation_psi: float = 1.0, truncation_cutoff=None, dtype=torch.float32):
    """networks.py:255-290 with c_dim == 0."""
    x = z.to(dtype)
    x = x * (x.square().mean(dim=1, keepdim=True) + 1e-8).rsqrt()
    for i in range(num_layers):
        x = fully_connected(x, p[f'mapping.fc{i}.weight'], p[f'mapping.fc{i}.bias'], 'lrelu', lr_multiplier)
    ws = x.unsqueeze(1).repeat(1, num_ws, 1)
    if truncation_psi != 1:
        w_avg = p['mapping.w_avg'].to(dtype)
        if truncation_cutoff is None:
            ws = w_avg.lerp(ws, truncation_psi)
        else:
            ws[:, :truncation_cutoff] = w_avg.lerp(ws[:, :truncation_cutoff], truncation_psi)
    return ws


def noise_grid(resolution: int) -> torch.Tensor:
    """``create_sampling_grid`` (networks.py:295-299): meshgrid 'ij', stack [xv, yv]."""
    lin = torch.linspace(0, 1, resolution)
    xv, yv = torch.meshgrid(lin, lin, indexing='ij')
    return torch.stack([xv, yv], dim=-1).unsqueeze(0)


def shifted_noise(noise_const: torch.Tensor, norm_positions: torch.Tensor) -> torch.Tensor:
    """networks.py:371-382 -- note channel 0 of the grid (grid_sample's x) carries
    the *row* coordinate, so the sample is the transpose of ``noise_const``
    shifted by ``norm_positions`` with wrap-around."""
    R = noise_const.shape[-1]
    B = norm_positions.shape[0]
    grid = ((noise_grid(R) + norm_positions.to(torch.float32).unsqueeze(1).unsqueeze(1)) % 1) * 2 - 1
    return F.grid_sample(noise_const.to(torch.float32)[None, None].expand(B, -1, -1, -1), grid,
                         padding_mode='reflection', align_corners=True)


def shifted_noise_closed_form(noise_const: torch.Tensor, norm_positions: torch.Tensor) -> torch.Tensor:
    """Index-level restatement of :func:`shifted_noise` (what the CUDA kernel does):
    ``out[b,i,j] = bilinear(noise, row=((lin[j]+p1)%1)(R-1), col=((lin[i]+p0)%1)(R-1))``."""
    R = noise_const.shape[-1]
    lin = torch.linspace(0, 1, R)
    out = torch.empty(norm_positions.shape[0], 1, R, R)
    nc = noise_const.to(torch.float32)
    for b in range(norm_positions.shape[0]):
        p0, p1 = norm_positions[b, 0].to(torch.float32), norm_positions[b, 1].to(torch.float32)
        gx = (((lin + p0) % 1) * 2 - 1)            # indexed by output row i -> input column
        gy = (((lin + p1) % 1) * 2 - 1)            # indexed by output col j -> input row
        cx = ((gx + 1) / 2) * (R - 1)
        cy = ((gy + 1) / 2) * (R - 1)
        x0 = cx.floor(); y0 = cy.floor()
        tx = cx - x0; ty = cy - y0
        x0 = x0.long().clamp(0, R - 1); y0 = y0.long().clamp(0, R - 1)
        x1 = (x0 + 1).clamp(0, R - 1); y1 = (y0 + 1).clamp(0, R - 1)
        # out[i, j]: column index from i, row index from j
        X0, X1, TX = x0[:, None], x1[:, None], tx[:, None]
        Y0, Y1, TY = y0[None, :], y1[None, :], ty[None, :]
        out[b, 0] = (nc[Y0, X0] * (1 - TX) * (1 - TY) + nc[Y0, X1] * TX * (1 - TY) +
                     nc[Y1, X0] * (1 - TX) * TY + nc[Y1, X1] * TX * TY)
    return out


def synthesis_layer(p: Bundle, prefix: str, x, w, up: int, noise_mode='const', norm_positions=None,
                    input_noise=None, gain: float = 1.0, conv_clamp: Optional[float] = 256.0,
                    resample_filter=None):
    """``SynthesisLayer.forward`` (networks.py:362-391)."""
    dtype = x.dtype
    styles = fully_connected(w, p[f'{prefix}.affine.weight'], p[f'{prefix}.affine.bias'])
    noise = None
    strength = p[f'{prefix}.noise_strength'].to(dtype)
    if noise_mode == 'const':
        nc = input_noise if input_noise is not None else p[f'{prefix}.noise_const']
        if norm_positions is not None:
            nc = shifted_noise(nc, norm_positions)
        noise = nc.to(dtype) * strength
    elif noise_mode == 'random':
        raise NotImplementedError('random noise is drawn by the caller; pass input_noise with noise_mode="const"')
    x = modulated_conv2d(x, p[f'{prefix}.weight'], styles, noise=noise, up=up, padding=1,
                         resample_filter=resample_filter, flip_weight=(up == 1))
    act_clamp = conv_clamp * gain if conv_clamp is not None else None
    return bias_act(x, p[f'{prefix}.bias'].to(dtype), act='lrelu', gain=SQRT2 * gain, clamp=act_clamp)


def torgb_triad(p: Bundle, prefix: str, x, w, conv_clamp: Optional[float] = 256.0, color_format: str = 'triad', extra: Optional[dict] = None):
    """``ToRGBColorTriadLayer.forward`` with ``color_w_channels == 0``
    (networks.py:451-485).  Returns (img, uvs, colors); for ``color_format == 'canvas'`` the weight has 3 + 5 output
    channels and ``extra`` (if given) receives 'canvas', 'alpha_fg', 'alpha' (networks.py:476-481)."""
    dtype = x.dtype
    cin = x.shape[1]
    scaled = fully_connected(w, p[f'{prefix}.affine.weight'], p[f'{prefix}.affine.bias'])
    colors = bias_act(scaled[:, :9], p[f'{prefix}.color_bias'].to(dtype), dim=1, act='tanh').reshape(-1, 3, 3)
    styles = scaled[:, 9:] * (1.0 / math.sqrt(cin))
    t = modulated_conv2d(x, p[f'{prefix}.weight'], styles, demodulate=False)
    t = bias_act(t, p[f'{prefix}.bias'].to(dtype), clamp=conv_clamp)
    uvs = torch.softmax(t[:, :3], dim=1)
    img = torch.sum(uvs.unsqueeze(1) * colors.unsqueeze(-1).unsqueeze(-1), dim=2)
    if color_format == 'canvas':
        canvas = t[:, 3:6]
        alpha = torch.softmax(t[:, 6:8], dim=1)
        img = alpha[:, :1] * img + alpha[:, 1:] * canvas
        if extra is not None:
            extra['canvas'], extra['alpha_fg'], extra['alpha'] = canvas, alpha[:, :1], alpha
    elif color_format != 'triad':
        raise RuntimeError(f'Unknown format {color_format}')
    return img, uvs, colors


# --------------------------------------------------------------------------------------
# SynthesisNetwork / Generator -- training/networks_modified.py:123-223,346-400 ;
# SynthesisBlock.forward networks.py:630-680 (architecture 'orig')
# --------------------------------------------------------------------------------------

class BlendedFeatures:
    """forger/train/stitching.py:18-25."""
    def __init__(self, features, alpha):
        self.features = features
        self.alpha = alpha

    def blend(self, other):
        return self.alpha * self.features + (1 - self.alpha) * other


def synthesis_network(p: Bundle, cfg, ws, geom_feature: Sequence[torch.Tensor], norm_positions=None,
                      return_features: Sequence[int] = (), blended_features: Optional[dict] = None,
                      noise_buffers: Optional[dict] = None, noise_mode='const', dtype=torch.float32):
    """Returns ``(img, debug)`` with debug keys 'uvs','colors','features{res}[_preblend]'.
    All-fp32 (= the reference's ``force_fp32=True``) unless ``dtype`` says float64."""
    blended_features = blended_features or {}
    rf = setup_filter([1, 3, 3, 1])
    ws = ws.to(dtype)
    debug = {}
    x = None
    img = None
    w_idx = 0
    geo_idx = 0
    B = ws.shape[0]
    last = cfg.block_resolutions[-1]
    for res in cfg.block_resolutions:
        name = f'synthesis.b{res}'
        nb = noise_buffers or {}
        n0 = nb.get(f'b{res}.conv0.noise_const')
        n1 = nb.get(f'b{res}.conv1.noise_const')
        if res == 4:
            x = p['synthesis.b4.const'].to(dtype).unsqueeze(0).repeat(B, 1, 1, 1)
            x = synthesis_layer(p, f'{name}.conv1', x, ws[:, w_idx], up=1, noise_mode=noise_mode,
                                norm_positions=norm_positions, input_noise=n1, conv_clamp=cfg.conv_clamp)
            nconv = 1
        else:
            x = synthesis_layer(p, f'{name}.conv0', x, ws[:, w_idx], up=2, noise_mode=noise_mode,
                                norm_positions=norm_positions, input_noise=n0, conv_clamp=cfg.conv_clamp,
                                resample_filter=rf)
            x = synthesis_layer(p, f'{name}.conv1', x, ws[:, w_idx + 1], up=1, noise_mode=noise_mode,
                                norm_positions=norm_positions, input_noise=n1, conv_clamp=cfg.conv_clamp)
            nconv = 2
        if res == last:
            img, uvs, colors = torgb_triad(p, f'{name}.torgb', x, ws[:, w_idx + nconv], conv_clamp=cfg.conv_clamp,
                                           color_format=getattr(cfg, 'color_format', 'triad'), extra=debug)
            debug['uvs'], debug['colors'] = uvs, colors
        if res in return_features:
            debug[f'features{res}_preblend'] = x
        if res in blended_features:
            x = blended_features[res].blend(x).to(x.dtype)
            if res == last:
                img, uvs, colors = torgb_triad(p, f'{name}.torgb', x, ws[:, w_idx + nconv], conv_clamp=cfg.conv_clamp,
                                               color_format=getattr(cfg, 'color_format', 'triad'), extra=debug)
                debug['uvs'], debug['colors'] = uvs, colors
        if res in return_features:
            debug[f'features{res}'] = x
        if res in cfg.geom_feature_resolutions:
            x = torch.cat([x, geom_feature[geo_idx].to(dtype)], dim=1)
            geo_idx += 1
        w_idx += nconv
    return img.to(torch.float32) if dtype == torch.float32 else img, debug


def generator_forward(p: Bundle, cfg, z, geom_feature, positions=None, truncation_psi=1.0, ws=None,
                      dtype=torch.float32, **synthesis_kwargs):
    """``Generator.forward`` / ``forward_pre_mapped`` (networks_modified.py:346-400)."""
    if ws is None:
        ws = mapping_network(p, z, cfg.num_ws, cfg.mapping_layers, cfg.mapping_lr_multiplier,
                             truncation_psi=truncation_psi, dtype=dtype)
    norm_positions = None
    if positions is not None:
        norm_positions = (positions % cfg.img_resolution) / (cfg.img_resolution - 1)
    img, debug = synthesis_network(p, cfg, ws, geom_feature, norm_positions=norm_positions, dtype=dtype,
                                   **synthesis_kwargs)
    debug['ws'] = ws
    return img, debug


# --------------------------------------------------------------------------------------
# Geometry encoder -- forger/experimental/autoenc/simple_autoencoder.py:95-126,155-199,251-297,
#                     base.py:32-58,123-134
# --------------------------------------------------------------------------------------

def _single_convolution(p: Bundle, prefix: str, x, stride: int, pad: int, neg_slope=0.01, eps=1e-5):
    """conv(reflect) -> eval BatchNorm -> LeakyReLU (simple_autoencoder.py:95-109)."""
    x = F.pad(x, (pad, pad, pad, pad), mode='reflect')
    x = F.conv2d(x, p[f'{prefix}.0.weight'].to(x.dtype), p[f'{prefix}.0.bias'].to(x.dtype), stride=stride)
    g, b = p[f'{prefix}.1.weight'].to(x.dtype), p[f'{prefix}.1.bias'].to(x.dtype)
    m, v = p[f'{prefix}.1.running_mean'].to(x.dtype), p[f'{prefix}.1.running_var'].to(x.dtype)
    x = (x - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + eps) * g[None, :, None, None] + b[None, :, None, None]
    return F.leaky_relu(x, neg_slope)


def _bn_eval(p: Bundle, prefix: str, x, eps):
    g, b = p[f'{prefix}.weight'].to(x.dtype), p[f'{prefix}.bias'].to(x.dtype)
    m, v = p[f'{prefix}.running_mean'].to(x.dtype), p[f'{prefix}.running_var'].to(x.dtype)
    return (x - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + eps) * g[None, :, None, None] + b[None, :, None, None]


def _single_convolution_bn_last(p: Bundle, prefix: str, x, stride: int, pad: int, neg_slope, eps=1e-5):
    """``batchnorm_after_activation=True``: conv(reflect) -> LeakyReLU -> eval BatchNorm (simple_autoencoder.py:100-103)."""
    x = F.pad(x, (pad, pad, pad, pad), mode='reflect')
    x = F.conv2d(x, p[f'{prefix}.0.weight'].to(x.dtype), p[f'{prefix}.0.bias'].to(x.dtype), stride=stride)
    return _bn_eval(p, f'{prefix}.2', F.leaky_relu(x, neg_slope), eps)


def _scale_up_v2(p: Bundle, prefix: str, x, neg_slope, eps=1e-5):
    """``ScaleUpV2``: ConvTranspose2d(3, stride 2, padding 1, output_padding 1) -> LeakyReLU -> eval BatchNorm
    (simple_autoencoder.py:128-148)."""
    x = F.conv_transpose2d(x, p[f'{prefix}.0.weight'].to(x.dtype), p[f'{prefix}.0.bias'].to(x.dtype), stride=2, padding=1, output_padding=1)
    return _bn_eval(p, f'{prefix}.2', F.leaky_relu(x, neg_slope), eps)


def encoder_preprocess(geom, preproc_type=None):
    """base.py:32-58."""
    if preproc_type in (None, 'none'):
        return geom
    if preproc_type == '-11inverse':
        return (1 - geom) * 2 - 1
    if preproc_type == 'inverse':
        return 1 - geom
    raise ValueError(preproc_type)


def geometry_encode(p: Bundle, ecfg, geom, dtype=torch.float32) -> List[torch.Tensor]:
    """``BaseGeoEncoder.encode`` -> ``AutoEncoder._encode`` for res list [0, 1]:
    returns ``[g0 [B,16,16,16], g1 [B,256,32,32]]`` for 128x128 input."""
    x = encoder_preprocess(geom.to(dtype), ecfg.preproc_type)
    if getattr(ecfg, 'bn_after_activation', False):
        # the --neg_slope variant (simple_autoencoder.py:48-53): BatchNorm after the activation, ScaleUpV2 decoder stages; the
        # post layers are built without a slope (:180-185) and keep LeakyReLU's default
        x = _single_convolution_bn_last(p, 'encoder.model.0.conv', x, 1, 3, ecfg.neg_slope, ecfg.bn_eps)
        idx = 1
        for _ in ecfg.down_filters:
            x = _single_convolution_bn_last(p, f'encoder.model.{idx}.conv', x, 2, 1, ecfg.neg_slope, ecfg.bn_eps)
            idx += 1
        for _ in ecfg.post_filters:
            x = _single_convolution_bn_last(p, f'encoder.model.{idx}.conv', x, 1, 1, ecfg.post_neg_slope, ecfg.bn_eps)
            idx += 1
        results = [x]
        for i in range(max(ecfg.encode_resolutions)):
            x = _scale_up_v2(p, f'decoder.model.{i}.conv', x, ecfg.neg_slope, ecfg.bn_eps)
            results.append(x)
        return [results[r] for r in ecfg.encode_resolutions]
    x = _single_convolution(p, 'encoder.model.0.conv', x, 1, 3, ecfg.neg_slope, ecfg.bn_eps)
    idx = 1
    for _ in ecfg.down_filters:
        x = _single_convolution(p, f'encoder.model.{idx}.conv', x, 2, 1, ecfg.neg_slope, ecfg.bn_eps)
        idx += 1
    for _ in ecfg.post_filters:
        x = _single_convolution(p, f'encoder.model.{idx}.conv', x, 1, 1, ecfg.neg_slope, ecfg.bn_eps)
        idx += 1
    results = [x]
    for i in range(max(ecfg.encode_resolutions)):
        x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)
        x = _single_convolution(p, f'decoder.model.{i}.conv.conv', x, 1, 1, ecfg.neg_slope, ecfg.bn_eps)
        results.append(x)
    return [results[r] for r in ecfg.encode_resolutions]


# --------------------------------------------------------------------------------------
# Engine composite -- forger/ui/brush.py:731-805, 514-527 ; UVS mapper forger/ui/mapper.py:53-72,117-135
# --------------------------------------------------------------------------------------

def map_style_s(sfactor, uvs):
    """``StyleUVSMapper._map_style_s`` (mapper.py:53-72)."""
    U, V, S = uvs[:, 0:1], uvs[:, 1:2], uvs[:, 2:3]
    Sp = torch.clamp_max(sfactor * S, 1.0)
    delta = 1 - Sp
    zero = delta <= 0.000001
    uvfactor = torch.where(zero, torch.zeros_like(delta), delta / (U + V))
    return torch.cat([uvfactor * U, uvfactor * V, Sp], dim=1)


def uvs_sfactor(S, bmask):
    """``get_sfactor`` (mapper.py:117-135): 1 / min_i min(topk15(S_i[bg_i]))."""
    vals = [torch.topk(S[i][bmask[i]], k=15)[0].min() for i in range(S.shape[0])]
    return 1 / torch.stack(vals).min()


def triad_composite(uvs, colors, render_mode='clear', color0=None, color1=None, canvas_color=None, sfactor=None):
    """``TriadGanPaintEngine._render_stroke_torch`` tail (brush.py:763-792):
    returns ``[B,4,H,W]`` straight RGBA in [0,1]."""
    C = (colors + 1) / 2.0
    if sfactor is not None:
        uvs = map_style_s(sfactor, uvs)
    C = C.clone()
    for idx, col in enumerate((color0, color1, canvas_color)):
        if col is not None:
            C[:, :, idx] = col
    stroke = torch.sum(uvs.unsqueeze(1) * C.unsqueeze(-1).unsqueeze(-1), dim=2)
    if render_mode == 'clear':
        alpha = torch.sum(uvs[:, 0:2], dim=1, keepdim=True)
    elif render_mode == 'full':
        alpha = torch.ones_like(stroke[:, :1])
    else:
        raise ValueError(render_mode)
    return torch.cat([stroke, alpha], dim=1)


def canvas_composite(uvs, colors, alpha_fg, gen_canvas, render_mode='clear', color0=None, color1=None, canvas_color=None):
    """``CanvasPaintEngine._render_stroke_torch`` tail (brush.py:893-935): ``[B,4,H,W]`` straight RGBA in [0,1];
    render modes 'clear' (UVS stroke colour + generated foreground alpha), 'stroke' (opaque UVS stroke), 'canvas'
    (the generated canvas only), 'full' (canvas under the stroke)."""
    C = ((colors + 1) / 2.0).clone()
    for idx, col in enumerate((color0, color1, canvas_color)):
        if col is not None:
            C[:, :, idx] = col
    stroke = torch.sum(uvs.unsqueeze(1) * C.unsqueeze(-1).unsqueeze(-1), dim=2)
    ones = torch.ones_like(stroke[:, :1])
    if render_mode == 'clear':
        return torch.cat([stroke, alpha_fg], dim=1)
    if render_mode == 'stroke':
        return torch.cat([stroke, ones], dim=1)
    if render_mode == 'canvas':
        return torch.cat([(gen_canvas + 1.0) / 2.0, ones], dim=1)
    if render_mode == 'full':
        return torch.cat([(1 - alpha_fg) * (gen_canvas + 1.0) / 2.0 + alpha_fg * stroke, ones], dim=1)
    raise ValueError(render_mode)


def to_uint8_tile(rgba, crop_margin: int):
    """``PaintingHelper.render_stroke`` tail (brush.py:369-377): crop the margin,
    x255, clip, *truncate* to uint8, HWC."""
    if crop_margin > 0:
        rgba = rgba[..., crop_margin: rgba.shape[-2] - crop_margin, crop_margin: rgba.shape[-1] - crop_margin]
    return (rgba.permute(0, 2, 3, 1) * 255).clip(0, 255).to(torch.uint8).numpy()


def prepare_geom_input(stroke_patch: np.ndarray) -> torch.Tensor:
    """``GanPaintEngine.prepare_geom_input`` (brush.py:672-681): last channel is
    the stroke alpha, 255 = stroke; output 0 = stroke, 1 = background."""
    g = 1 - torch.from_numpy(np.ascontiguousarray(stroke_patch[:, :, -1:])).to(torch.float32).permute(2, 0, 1) / 255.0
    return g.unsqueeze(0)


# --------------------------------------------------------------------------------------
# Patch scheduling / tile placement -- forger/viz/style_transfer.py:15-48,
# forger/viz/paint_image_main.py:59-62,145-186
# --------------------------------------------------------------------------------------

def pad_geo(geo: np.ndarray, crop_margin: int) -> np.ndarray:
    """paint_image_main.py:59-62."""
    out = np.ones((geo.shape[0] + crop_margin, geo.shape[1] + crop_margin, geo.shape[2]), dtype=np.uint8) * 255
    out[crop_margin:, crop_margin:, :] = geo
    return out


def generate_stitching_crops(stroke_image: np.ndarray, patch_width: int, mode='all', overlap_margin=15):
    """style_transfer.py:15-48."""
    rwidth = patch_width - overlap_margin * 2
    H, W, C = stroke_image.shape
    nrows = H // rwidth + 1
    ncols = W // rwidth + 1
    padded = np.full((nrows * rwidth + patch_width, ncols * rwidth + patch_width, C), 255, dtype=np.uint8)
    padded[:H, :W] = stroke_image
    crops = []
    for r in range(nrows):
        for c in range(ncols):
            y, x = r * rwidth, c * rwidth
            patch = padded[y:y + patch_width, x:x + patch_width]
            if mode == 'all' or np.sum(patch < 0.001) > 10:
                crops.append((y, x, patch_width, patch_width))
    return crops, padded


def place_tiles(canvas_hw: Tuple[int, int], tiles: Sequence[np.ndarray], metas: Sequence[Tuple[int, int]]):
    """The placement loop of paint_image_main.py:155-177: later tiles overwrite."""
    result = np.zeros((canvas_hw[0], canvas_hw[1], 4), dtype=np.uint8)
    for tile, (y, x) in zip(tiles, metas):
        result[y:y + tile.shape[0], x:x + tile.shape[1], :] = tile
    return result


def composite_on_white(result: np.ndarray) -> np.ndarray:
    """paint_image_main.py:179-183 (the ``result[..., 3:] = 255`` line is a no-op
    on the 3-channel array)."""
    alpha = result[..., 3:].astype(np.float32) / 255
    out = result[..., :3].astype(np.float32) * alpha + 255 * (1 - alpha)
    return out.clip(0, 255).astype(np.uint8)


def tile_owner_closed_form(Y: np.ndarray, X: np.ndarray, crop_margin: int, rwidth: int, nrows: int, ncols: int):
    """Closed form of last-writer-wins for ``stitching_mode='all'``:
    pixel (Y, X) of the padded canvas is owned by tile
    ``(clip((Y-m)//rwidth, 0, nrows-1), clip((X-m)//rwidth, 0, ncols-1))``; pixels with
    Y < m or X < m are never written."""
    r = np.clip((Y - crop_margin) // rwidth, 0, nrows - 1)
    c = np.clip((X - crop_margin) // rwidth, 0, ncols - 1)
    return r, c


# --------------------------------------------------------------------------------------
# Feature blending -- forger/ui/brush.py:159-242 ; forger/train/stitching.py:34-37,86-89,110-120
# --------------------------------------------------------------------------------------

def dirty_area_alpha(width: int, margin: int, crop_margin: int = 0) -> torch.Tensor:
    """``generate_dirty_area_alpha`` for the full-patch dirty area (brush.py:159-187):
    1 inside ``[margin+crop, width-margin-crop)``, linear fall-off of width
    ``margin`` outside (Euclidean distance in the corners)."""
    lo = margin + crop_margin
    hi = lo + width - 2 * margin - 2 * crop_margin            # exclusive end
    x = torch.linspace(0, width - 1, steps=width)
    gy, gx = torch.meshgrid(x, x, indexing='ij')
    dx = torch.min((gx - lo) ** 2, (gx - hi + 1) ** 2)
    dy = torch.min((gy - lo) ** 2, (gy - hi + 1) ** 2)
    d = dx + dy
    d[0:lo, lo:hi] = dy[0:lo, lo:hi]
    d[hi:, lo:hi] = dy[hi:, lo:hi]
    d[lo:hi, 0:lo] = dx[lo:hi, 0:lo]
    d[lo:hi, hi:] = dx[lo:hi, hi:]
    res = 1 - torch.sqrt(d) / margin
    res[res < 0] = 0
    res[lo:hi, lo:hi] = 1
    return res


class FeatureCanvasOracle:
    """Sequential (raster-order) feature blending exactly as
    ``PaintingHelper.render_stroke`` does it with ``feature_blending_level=L``
    (brush.py:190-242, 33-92).  ``level`` 2 -> features after b64."""
    def __init__(self, canvas_h: int, canvas_w: int, level: int, patch_width: int = 128, blending_margin: int = 16):
        self.down = 2 ** (level - 1)
        self.res = patch_width // self.down
        self.h = int(math.ceil(canvas_h / self.down))
        self.w = int(math.ceil(canvas_w / self.down))
        self.margin = blending_margin // self.down
        self.features = None
        self.mask = None

    def snap(self, v: int) -> int:
        return (v // self.down) * self.down

    def inputs_for(self, y: int, x: int, crop_margin: int):
        """-> (BlendedFeatures or None, update_mask) for the patch at canvas (y, x)."""
        cm = crop_margin // self.down
        alpha = dirty_area_alpha(self.res, self.margin, cm)
        update = alpha > 0.99
        blended = None
        ys, xs = y // self.down, x // self.down
        if self.mask is not None:
            m = self.mask[ys:ys + self.res, xs:xs + self.res]
            feat = self.features[..., ys:ys + self.res, xs:xs + self.res]
            update = update | (m & (alpha > 0))
            alpha = alpha.clone()
            alpha[~m] = 1
            blended = BlendedFeatures(feat.clone(), (1 - alpha)[None, None])
        if cm > 0:
            update = update.clone()
            update[:cm, :] = False
            update[-cm:, :] = False
            update[:, :cm] = False
            update[:, -cm:] = False
        return blended, update

    def update(self, y: int, x: int, feature_patch: torch.Tensor, update_mask: torch.Tensor):
        if self.features is None:
            C = feature_patch.shape[1]
            self.features = torch.zeros(1, C, self.h, self.w, dtype=feature_patch.dtype)
            self.mask = torch.zeros(self.h, self.w, dtype=torch.bool)
        ys, xs = y // self.down, x // self.down
        self.mask[ys:ys + self.res, xs:xs + self.res][update_mask] = True
        um = update_mask[None, None].expand(-1, self.features.shape[1], -1, -1)
        self.features[..., ys:ys + self.res, xs:xs + self.res][um] = feature_patch[um]


# --------------------------------------------------------------------------------------
# End-to-end stylizer -- forger/viz/paint_image_main.py:145-186 (the loop the patch scheduler replaces)
# --------------------------------------------------------------------------------------

def stylize(gp: Bundle, ep: Bundle, cfg, ecfg, guidance: np.ndarray, z, crop_margin: int = 10,
            stitching_mode: str = 'all', render_mode: str = 'clear', feature_blending_level: int = 0,
            sfactor=None, color0=None, color1=None, z_per_patch=None, on_white: bool = False,
            max_patches: Optional[int] = None):
    """guidance: [H, W, 1] uint8, 0 = stroke.  Returns (canvas uint8 cropped to the
    input size, crops, metas)."""
    patch = cfg.img_resolution
    H0, W0 = guidance.shape[:2]
    geom = pad_geo(guidance, crop_margin)
    crops, geom = generate_stitching_crops(geom, patch, mode=stitching_mode, overlap_margin=crop_margin * 2)
    fc = FeatureCanvasOracle(geom.shape[0], geom.shape[1], feature_blending_level, patch) if feature_blending_level > 0 else None
    tiles, metas = [], []
    for i, (y, x, _, _) in enumerate(crops[:max_patches]):
        gpatch = 255 - geom[y:y + patch, x:x + patch, :]
        g = prepare_geom_input(gpatch)
        gf = geometry_encode(ep, ecfg, g)
        zi = z if z_per_patch is None else z_per_patch[i:i + 1]
        kwargs = {}
        yy, xx = y, x
        upd = None
        if fc is not None:
            yy, xx = fc.snap(y), fc.snap(x)
            blended, upd = fc.inputs_for(yy, xx, crop_margin)
            kwargs['return_features'] = [fc.res]
            kwargs['blended_features'] = {fc.res: blended} if blended is not None else {}
        _, dbg = generator_forward(gp, cfg, zi, gf, positions=torch.tensor([[y, x]]), **kwargs)
        if fc is not None:
            fc.update(yy, xx, dbg[f'features{fc.res}'], upd)
        rgba = triad_composite(dbg['uvs'], dbg['colors'], render_mode, color0, color1, None, sfactor)
        tiles.append(to_uint8_tile(rgba, crop_margin)[0])
        metas.append((yy + crop_margin, xx + crop_margin))
    canvas = place_tiles(geom.shape[:2], tiles, metas)
    if on_white:
        canvas = composite_on_white(canvas)
    return canvas[crop_margin:crop_margin + H0, crop_margin:crop_margin + W0], crops, metas
