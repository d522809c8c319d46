// Small / memory-bound operators of the generator forward: fully-connected layers (mapping + affines),
// shifted noise, NCHW<->NHWC packing, the FIR-first x2 upsample feeding the tensor-core conv, the
// fused ToRGB + triad composite, and feature blending.
#include "common.cuh"

namespace nbe {

// ---------------------------------------------------------------------------------------------
// Fully connected: one warp per output element (n, o); x row optionally 2nd-moment normalised.
// ---------------------------------------------------------------------------------------------
template <class XT>
__global__ void __launch_bounds__(256)
fc_kernel(const XT* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ y,
          int N, int In, int Out, int64_t xs_n, int64_t ys_n, float wgain, float bgain, int act, float alpha,
          float act_gain, int normalize) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= N * Out) return;
    const int n = warp / Out, o = warp - n * Out;
    const XT* xr = x + n * xs_n;
    const float* wr = w + (int64_t)o * In;
    float dot = 0.f, sq = 0.f;
    for (int i = lane; i < In; i += 32) {
        float xv = (float)xr[i];
        dot = fmaf(xv, wr[i] * wgain, dot);
        sq = fmaf(xv, xv, sq);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        dot += __shfl_xor_sync(0xffffffffu, dot, off);
        sq += __shfl_xor_sync(0xffffffffu, sq, off);
    }
    if (lane == 0) {
        if (normalize) dot *= rsqrtf(sq / (float)In + 1e-8f);      // x * rsqrt(mean(x^2) + eps), networks.py:24-26
        if (b) dot += b[o] * bgain;
        y[n * ys_n + o] = apply_act(dot, act, alpha) * act_gain;
    }
}

// ---------------------------------------------------------------------------------------------
// Shifted noise (grid_sample restated; op order mirrors ATen's fp32 arithmetic so that the wrap
// discontinuity at frac == 0 falls on the same element).
// ---------------------------------------------------------------------------------------------
__global__ void shifted_noise_kernel(const float* __restrict__ nc, const float* __restrict__ lin,
                                     const int64_t* __restrict__ positions, float* __restrict__ out, int N, int R, int mod) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * R * R) return;
    const int j = idx % R, i = (idx / R) % R, n = idx / (R * R);
    int64_t py = positions[2 * n] % mod, px = positions[2 * n + 1] % mod;
    if (py < 0) py += mod;                                         // python % semantics
    if (px < 0) px += mod;
    const float p0 = __fdiv_rn((float)py, (float)(mod - 1));
    const float p1 = __fdiv_rn((float)px, (float)(mod - 1));
    // grid x (-> input column) comes from the output ROW i, grid y (-> input row) from the output COLUMN j
    float sx = __fadd_rn(lin[i], p0); sx = __fsub_rn(sx, floorf(sx));
    float sy = __fadd_rn(lin[j], p1); sy = __fsub_rn(sy, floorf(sy));
    const float gx = __fsub_rn(__fmul_rn(sx, 2.f), 1.f), gy = __fsub_rn(__fmul_rn(sy, 2.f), 1.f);
    const float cx = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(R - 1));
    const float cy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(R - 1));
    const float fx0 = floorf(cx), fy0 = floorf(cy);
    const float tx = cx - fx0, ty = cy - fy0;
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    auto at = [&](int yy, int xx) -> float { return (yy >= 0 && yy < R && xx >= 0 && xx < R) ? nc[yy * R + xx] : 0.f; };
    out[idx] = at(y0, x0) * (1.f - tx) * (1.f - ty) + at(y0, x1) * tx * (1.f - ty) +
               at(y1, x0) * (1.f - tx) * ty + at(y1, x1) * tx * ty;
}

// ---------------------------------------------------------------------------------------------
// NCHW f32 -> NHWC bf16 through a 32(c) x 32(pixel) shared-memory transpose; and back.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int HW, int dst_cs, int c_off,
                 const float* __restrict__ scale) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        int c = c0 + r, p = p0 + tx;
        float v = 0.f;
        if (c < C && p < HW) {
            v = src[((int64_t)n * C + c) * HW + p];
            if (scale) v *= scale[(int64_t)n * C + c];
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int p = p0 + r, c = c0 + tx;
        if (c < C && p < HW) dst[((int64_t)n * HW + p) * dst_cs + c_off + c] = __float2bfloat16_rn(tile[tx][r]);
    }
}

__global__ void __launch_bounds__(256)
unpack_nchw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int C, int HW, int src_cs) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        int p = p0 + r, c = c0 + tx;
        tile[r][tx] = (c < C && p < HW) ? __bfloat162float(src[((int64_t)n * HW + p) * src_cs + c]) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int c = c0 + r, p = p0 + tx;
        if (c < C && p < HW) dst[((int64_t)n * C + c) * HW + p] = tile[tx][r];
    }
}

// ---------------------------------------------------------------------------------------------
// FIR-first x2 upsample on NHWC bf16:  U[n, uy, ux, c] = sum over the (<=2)x(<=2) taps with even parity of
//   g[a,b] * x[n, (uy+a-3)/2, (ux+b-3)/2, c] * scale[n,c],   g = f flipped * 4,  U is (2H+2) x (2W+2).
// The four outputs of a 2x2 output quad (uy = 2q, 2q+1; ux = 2p, 2p+1) read the SAME 2x2 input pixels
// (rows q-1, q; cols p-1, p), so one thread = one quad x 8 channels: 4 x 16-byte loads, 16 tap products per channel,
// 4 x 16-byte stores.  Consecutive threads walk the channel vectors of a quad, then the quads of a row: loads and
// stores are whole 32-byte sectors.  HBM-bound: (H*W + (2H+2)(2W+2)) * C * 2 bytes per image.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
upsample2x_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ f, const float* __restrict__ scale,
                       __nv_bfloat16* __restrict__ u, int N, int H, int W, int C, int xs_c, int x_pitch) {
    __shared__ float s_g[16];
    if (threadIdx.x < 16) {
        int a = threadIdx.x >> 2, b = threadIdx.x & 3;
        s_g[threadIdx.x] = f[(3 - a) * 4 + (3 - b)] * 4.f;          // flip_filter = False, gain = up^2
    }
    __syncthreads();
    const int CV = C >> 3;
    const int QW = W + 1;                                           // quads per output row
    const int UW = 2 * W + 2, UH = 2 * H + 2;
    const int q = blockIdx.x % (H + 1);                             // quad row
    const int n = blockIdx.x / (H + 1);
    const int idx = blockIdx.y * blockDim.x + threadIdx.x;
    if (idx >= QW * CV) return;
    const int cv = idx % CV, pq = idx / CV;
    // input pixels (q-1, q) x (pq-1, pq); out-of-range -> 0
    float in[2][2][8];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int iy = q - 1 + dy, ix = pq - 1 + dx;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
                const int4 raw = ld_stream16(x + (((long long)n * H + iy) * x_pitch + ix) * xs_c + cv * 8);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 v = __bfloat1622float2(h2[k]);
                    in[dy][dx][2 * k] = v.x; in[dy][dx][2 * k + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) in[dy][dx][k] = 0.f;
            }
        }
    float sc[8];
    if (scale) {
        const float4 s0 = *reinterpret_cast<const float4*>(scale + (long long)n * C + cv * 8);
        const float4 s1 = *reinterpret_cast<const float4*>(scale + (long long)n * C + cv * 8 + 4);
        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) sc[k] = 1.f;
    }
    // output (uy = 2q + py, ux = 2pq + px): taps a = ((uy+1)&1) + 2*da -> input row (uy + a - 3) >> 1 = q - 1 + da
    // (py = 0: a = 1 + 2da ; py = 1: a = 2da), likewise for columns.
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
        for (int px = 0; px < 2; ++px) {
            float acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
            for (int da = 0; da < 2; ++da)
#pragma unroll
                for (int db = 0; db < 2; ++db) {
                    const float g = s_g[((1 - py) + 2 * da) * 4 + (1 - px) + 2 * db];
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] = fmaf(g, in[da][db][k], acc[k]);
                }
            int4 outv;
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&outv);
#pragma unroll
            for (int k = 0; k < 4; ++k) o2[k] = __floats2bfloat162_rn(acc[2 * k] * sc[2 * k], acc[2 * k + 1] * sc[2 * k + 1]);
            const int uy = 2 * q + py, ux = 2 * pq + px;
            st_stream16(u + (((long long)n * UH + uy) * UW + ux) * C + cv * 8, outv);
        }
}

// ---------------------------------------------------------------------------------------------
// ToRGB (1x1 modconv, no demod) + bias + clamp + softmax(3) + triad colour mix.
// NHWC bf16 input: one warp per 32 pixels?  No -- one thread per pixel would stride 256 B between lanes.
// Instead a warp cooperates on 8 pixels at a time: 4 lanes per pixel, each lane reads 32 channels
// (4 x 16-byte vectors, the 4 lanes of a pixel cover its 256 contiguous bytes), partial dot products are
// combined with two shuffles.  Reads are fully coalesced; writes are per-plane NCHW rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
torgb_triad_nhwc_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, const float* __restrict__ w,
                        const float* __restrict__ styles, const float* __restrict__ bias,
                        const float* __restrict__ colors, float clamp, float* __restrict__ img, float* __restrict__ uvs,
                        int N, int C, int HW) {
    extern __shared__ float s_w[];                                 // [3][C] modulated weights of this image
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_w[i] = w[i] * styles[(int64_t)n * C + (i % C)];
    __syncthreads();
    const int sub = threadIdx.x & 3;                               // quarter of the channel range
    const int cpl = C / 4;                                         // channels per lane
    float col[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) col[i] = colors[n * 9 + i];
    const float b0 = bias[0], b1 = bias[1], b2 = bias[2];
    for (int pix = blockIdx.x * (blockDim.x / 4) + (threadIdx.x >> 2); pix < HW; pix += gridDim.x * (blockDim.x / 4)) {
        const __nv_bfloat16* xp = x + ((int64_t)n * HW + pix) * x_cs + sub * cpl;
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        for (int c = 0; c < cpl; c += 8) {
            const int4 raw = *reinterpret_cast<const int4*>(xp + c);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 v = __bfloat1622float2(h2[k]);
                const int cc = sub * cpl + c + 2 * k;
                t0 = fmaf(v.x, s_w[cc], t0);          t0 = fmaf(v.y, s_w[cc + 1], t0);
                t1 = fmaf(v.x, s_w[C + cc], t1);      t1 = fmaf(v.y, s_w[C + cc + 1], t1);
                t2 = fmaf(v.x, s_w[2 * C + cc], t2);  t2 = fmaf(v.y, s_w[2 * C + cc + 1], t2);
            }
        }
        t0 += __shfl_xor_sync(0xffffffffu, t0, 1); t1 += __shfl_xor_sync(0xffffffffu, t1, 1); t2 += __shfl_xor_sync(0xffffffffu, t2, 1);
        t0 += __shfl_xor_sync(0xffffffffu, t0, 2); t1 += __shfl_xor_sync(0xffffffffu, t1, 2); t2 += __shfl_xor_sync(0xffffffffu, t2, 2);
        t0 += b0; t1 += b1; t2 += b2;
        if (clamp >= 0.f) {
            t0 = fminf(fmaxf(t0, -clamp), clamp); t1 = fminf(fmaxf(t1, -clamp), clamp); t2 = fminf(fmaxf(t2, -clamp), clamp);
        }
        const float m = fmaxf(t0, fmaxf(t1, t2));
        const float e0 = expf(t0 - m), e1 = expf(t1 - m), e2 = expf(t2 - m);
        const float inv = 1.f / (e0 + e1 + e2);
        const float u[3] = {e0 * inv, e1 * inv, e2 * inv};
        if (sub < 3) {                                              // lane `sub` writes channel `sub`
            const int64_t o = ((int64_t)n * 3 + sub) * HW + pix;
            if (uvs) uvs[o] = u[sub];
            if (img) img[o] = u[0] * col[sub * 3 + 0] + u[1] * col[sub * 3 + 1] + u[2] * col[sub * 3 + 2];
        }
    }
}

// NCHW float32 variant (FP32 mode): one thread per pixel, channel loop with coalesced plane reads.
__global__ void __launch_bounds__(256)
torgb_triad_nchw_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ styles,
                        const float* __restrict__ bias, const float* __restrict__ colors, float clamp,
                        float* __restrict__ img, float* __restrict__ uvs, int N, int C, int HW) {
    extern __shared__ float s_w[];
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_w[i] = w[i] * styles[(int64_t)n * C + (i % C)];
    __syncthreads();
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    const float* xp = x + (int64_t)n * C * HW + pix;
    for (int c = 0; c < C; ++c) {
        const float v = xp[(int64_t)c * HW];
        t0 = fmaf(v, s_w[c], t0); t1 = fmaf(v, s_w[C + c], t1); t2 = fmaf(v, s_w[2 * C + c], t2);
    }
    t0 += bias[0]; t1 += bias[1]; t2 += bias[2];
    if (clamp >= 0.f) {
        t0 = fminf(fmaxf(t0, -clamp), clamp); t1 = fminf(fmaxf(t1, -clamp), clamp); t2 = fminf(fmaxf(t2, -clamp), clamp);
    }
    const float m = fmaxf(t0, fmaxf(t1, t2));
    const float e0 = expf(t0 - m), e1 = expf(t1 - m), e2 = expf(t2 - m);
    const float inv = 1.f / (e0 + e1 + e2);
    const float u[3] = {e0 * inv, e1 * inv, e2 * inv};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int64_t o = ((int64_t)n * 3 + c) * HW + pix;
        if (uvs) uvs[o] = u[c];
        if (img) img[o] = u[0] * colors[n * 9 + c * 3 + 0] + u[1] * colors[n * 9 + c * 3 + 1] + u[2] * colors[n * 9 + c * 3 + 2];
    }
}

// ---------------------------------------------------------------------------------------------
// ToRGBColorTriadLayer in the 'canvas' colour format (networks.py:433-481): 3 UVS logits + a 3-channel canvas + 2 alpha
// logits from one 1x1 modulated conv; uvs = softmax(t[0:3]), alpha = softmax(t[6:8]),
// img = alpha_fg * sum_k uvs_k colors_k + alpha_bg * canvas.  One thread per pixel, 8 dot products over the channels.
// ---------------------------------------------------------------------------------------------
template <bool NHWC>
__global__ void __launch_bounds__(128)
torgb_canvas_kernel(const void* __restrict__ xv, int x_cs, const float* __restrict__ w, const float* __restrict__ styles,
                    const float* __restrict__ bias, const float* __restrict__ colors, float clamp,
                    float* __restrict__ img, float* __restrict__ uvs, float* __restrict__ canvas, float* __restrict__ alpha,
                    int N, int C, int HW) {
    extern __shared__ float s_w[];                                 // [8][C] modulated weights of this image
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < 8 * C; i += blockDim.x) s_w[i] = w[i] * styles[(int64_t)n * C + (i % C)];
    __syncthreads();
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    float t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = 0.f;
    if (NHWC) {
        const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(xv) + ((int64_t)n * HW + pix) * x_cs;
        for (int c = 0; c < C; c += 8) {
            const int4 raw = *reinterpret_cast<const int4*>(xp + c);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 v = __bfloat1622float2(h2[q]);
#pragma unroll
                for (int k = 0; k < 8; ++k) { t[k] = fmaf(v.x, s_w[k * C + c + 2 * q], t[k]); t[k] = fmaf(v.y, s_w[k * C + c + 2 * q + 1], t[k]); }
            }
        }
    } else {
        const float* xp = reinterpret_cast<const float*>(xv) + (int64_t)n * C * HW + pix;
        for (int c = 0; c < C; ++c) {
            const float v = xp[(int64_t)c * HW];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = fmaf(v, s_w[k * C + c], t[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        t[k] += bias[k];
        if (clamp >= 0.f) t[k] = fminf(fmaxf(t[k], -clamp), clamp);
    }
    const float m = fmaxf(t[0], fmaxf(t[1], t[2]));
    const float e0 = expf(t[0] - m), e1 = expf(t[1] - m), e2 = expf(t[2] - m);
    const float inv = 1.f / (e0 + e1 + e2);
    const float u[3] = {e0 * inv, e1 * inv, e2 * inv};
    const float ma = fmaxf(t[6], t[7]);
    const float a0 = expf(t[6] - ma), a1 = expf(t[7] - ma);
    const float ainv = 1.f / (a0 + a1);
    const float afg = a0 * ainv, abg = a1 * ainv;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int64_t o = ((int64_t)n * 3 + c) * HW + pix;
        const float stroke = u[0] * colors[n * 9 + c * 3 + 0] + u[1] * colors[n * 9 + c * 3 + 1] + u[2] * colors[n * 9 + c * 3 + 2];
        if (uvs) uvs[o] = u[c];
        if (canvas) canvas[o] = t[3 + c];
        if (img) img[o] = afg * stroke + abg * t[3 + c];
    }
    if (alpha) { alpha[((int64_t)n * 2 + 0) * HW + pix] = afg; alpha[((int64_t)n * 2 + 1) * HW + pix] = abg; }
}

// ---------------------------------------------------------------------------------------------
// BlendedFeatures.blend: x = alpha * saved + (1 - alpha) * x
// ---------------------------------------------------------------------------------------------
__global__ void blend_nchw_kernel(float* __restrict__ x, const float* __restrict__ saved, const float* __restrict__ alpha,
                                  int64_t alpha_sn, int C, int HW, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int pix = (int)(i % HW);
        const int n = (int)(i / ((int64_t)C * HW));
        const float a = alpha[n * alpha_sn + pix];
        x[i] = a * saved[i] + (1.f - a) * x[i];
    }
}
__global__ void blend_nhwc_kernel(__nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ saved,
                                  const float* __restrict__ alpha, int64_t alpha_sn, int C, int HW, int cs, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t pixn = i / C;
        const int pix = (int)(pixn % HW);
        const int n = (int)(pixn / HW);
        const float a = alpha[n * alpha_sn + pix];
        const int64_t o = pixn * cs + c;
        x[o] = __float2bfloat16_rn(a * __bfloat162float(saved[o]) + (1.f - a) * __bfloat162float(x[o]));
    }
}

}  // namespace nbe

using namespace nbe;

extern "C" int nbe_fc_f32(const void* x, int x_is_f64, const float* w, const float* b, float* y, int N, int In, int Out,
                          int64_t xs_n, int64_t ys_n, float wgain, float bgain, int act, float alpha, float act_gain,
                          int normalize, nbe_stream_t stream) {
    NBE_REQUIRE(x && w && y && N >= 0 && In >= 1 && Out >= 1, "fc: bad arguments");
    NBE_REQUIRE(act >= NBE_ACT_LINEAR && act <= NBE_ACT_SWISH, "fc: bad activation");
    if (N == 0) return NBE_OK;
    const int64_t threads = (int64_t)N * Out * 32;
    const int blocks = (int)((threads + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (x_is_f64) fc_kernel<double><<<blocks, 256, 0, s>>>((const double*)x, w, b, y, N, In, Out, xs_n, ys_n, wgain, bgain, act, alpha, act_gain, normalize);
    else          fc_kernel<float><<<blocks, 256, 0, s>>>((const float*)x, w, b, y, N, In, Out, xs_n, ys_n, wgain, bgain, act, alpha, act_gain, normalize);
    return launched("fc_kernel");
}

extern "C" int nbe_shifted_noise_f32(const float* noise_const, const float* lin, const int64_t* positions, float* out,
                                     int N, int R, int mod, nbe_stream_t stream) {
    NBE_REQUIRE(noise_const && lin && positions && out && N >= 0 && R >= 2 && mod >= 2, "shifted_noise: bad arguments");
    if (N == 0) return NBE_OK;
    const int total = N * R * R;
    shifted_noise_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(noise_const, lin, positions, out, N, R, mod);
    return launched("shifted_noise_kernel");
}

extern "C" int nbe_pack_nhwc_bf16(const float* src, void* dst, int N, int C, int H, int W, int dst_cs, int c_off,
                                  const float* scale, nbe_stream_t stream) {
    NBE_REQUIRE(src && dst && N >= 0 && C >= 1 && H >= 1 && W >= 1 && c_off >= 0 && c_off + C <= dst_cs, "pack_nhwc: bad arguments");
    if (N == 0) return NBE_OK;
    NBE_REQUIRE(N <= 65535, "pack_nhwc: batch too large for one launch");
    dim3 grid((H * W + 31) / 32, (C + 31) / 32, N);
    pack_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, C, H * W, dst_cs, c_off, scale);
    return launched("pack_nhwc_kernel");
}

extern "C" int nbe_unpack_nchw_f32(const void* src, float* dst, int N, int C, int H, int W, int src_cs, nbe_stream_t stream) {
    NBE_REQUIRE(src && dst && N >= 0 && C >= 1 && H >= 1 && W >= 1 && C <= src_cs, "unpack_nchw: bad arguments");
    if (N == 0) return NBE_OK;
    NBE_REQUIRE(N <= 65535, "unpack_nchw: batch too large for one launch");
    dim3 grid((H * W + 31) / 32, (C + 31) / 32, N);
    unpack_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, dst, C, H * W, src_cs);
    return launched("unpack_nchw_kernel");
}

extern "C" int nbe_upsample2x_nhwc_bf16(const void* x, const float* f, const float* scale, void* u,
                                        int N, int H, int W, int C, int xs_c, nbe_stream_t stream) {
    return nbe_upsample2x_nhwc_bf16_ex(x, f, scale, u, N, H, W, C, xs_c, W, stream);
}

extern "C" int nbe_upsample2x_nhwc_bf16_ex(const void* x, const float* f, const float* scale, void* u,
                                           int N, int H, int W, int C, int xs_c, int x_pitch, nbe_stream_t stream) {
    NBE_REQUIRE(x_pitch >= W, "upsample2x: input row pitch smaller than the width");
    NBE_REQUIRE(x && f && u && N >= 0 && H >= 1 && W >= 1 && C >= 8, "upsample2x: bad arguments");
    NBE_REQUIRE(C % 8 == 0 && xs_c % 8 == 0 && xs_c >= C, "upsample2x: channels must be a multiple of 8");
    NBE_REQUIRE((((uintptr_t)x | (uintptr_t)u) & 15) == 0, "upsample2x: tensors must be 16-byte aligned");
    if (N == 0) return NBE_OK;
    NBE_REQUIRE((int64_t)N * (H + 1) <= INT32_MAX, "upsample2x: too many rows");
    const int per_row = (W + 1) * (C / 8);
    dim3 grid(N * (H + 1), (per_row + 255) / 256);
    NBE_REQUIRE(grid.y <= 65535u, "upsample2x: rows too wide");
    upsample2x_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, f, scale, (__nv_bfloat16*)u, N, H, W, C, xs_c, x_pitch);
    return launched("upsample2x_nhwc_kernel");
}

extern "C" int nbe_torgb_triad(const void* x, int x_is_bf16, int x_cs, const float* w, const float* styles, const float* bias,
                               const float* colors, float clamp, float* img, float* uvs, int N, int C, int H, int W,
                               nbe_stream_t stream) {
    NBE_REQUIRE(x && w && styles && bias && colors && (img || uvs), "torgb_triad: null tensor");
    NBE_REQUIRE(N >= 0 && C >= 1 && H >= 1 && W >= 1, "torgb_triad: bad shape");
    if (N == 0) return NBE_OK;
    NBE_REQUIRE(N <= 65535, "torgb_triad: batch too large for one launch");
    const int HW = H * W;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)3 * C * sizeof(float);
    NBE_REQUIRE(smem <= 48 * 1024, "torgb_triad: too many channels");
    if (x_is_bf16) {
        NBE_REQUIRE(C % 32 == 0 && x_cs % 8 == 0 && x_cs >= C && ((uintptr_t)x & 15) == 0, "torgb_triad: NHWC input needs C %% 32 == 0");
        int bx = (HW + 63) / 64;
        if (bx > 64) bx = 64;
        dim3 grid(bx, N);
        torgb_triad_nhwc_kernel<<<grid, 256, smem, s>>>((const __nv_bfloat16*)x, x_cs, w, styles, bias, colors, clamp, img, uvs, N, C, HW);
        return launched("torgb_triad_nhwc_kernel");
    }
    dim3 grid((HW + 255) / 256, N);
    torgb_triad_nchw_kernel<<<grid, 256, smem, s>>>((const float*)x, w, styles, bias, colors, clamp, img, uvs, N, C, HW);
    return launched("torgb_triad_nchw_kernel");
}

extern "C" int nbe_torgb_canvas(const void* x, int x_is_bf16, int x_cs, const float* w, const float* styles, const float* bias,
                                const float* colors, float clamp, float* img, float* uvs, float* canvas, float* alpha,
                                int N, int C, int H, int W, nbe_stream_t stream) {
    NBE_REQUIRE(x && w && styles && bias && colors && N >= 0 && C >= 1 && H >= 1 && W >= 1, "torgb_canvas: bad arguments");
    if (N == 0) return NBE_OK;
    NBE_REQUIRE(N <= 65535, "torgb_canvas: batch too large");
    const int HW = H * W;
    const size_t smem = (size_t)8 * C * sizeof(float);
    NBE_REQUIRE(smem <= 48 * 1024, "torgb_canvas: too many channels");
    dim3 grid((HW + 127) / 128, N);
    cudaStream_t s = (cudaStream_t)stream;
    if (x_is_bf16) {
        NBE_REQUIRE(C % 8 == 0 && x_cs % 8 == 0 && x_cs >= C && (((uintptr_t)x) & 15) == 0, "torgb_canvas: NHWC bf16 input needs C, x_cs multiples of 8");
        torgb_canvas_kernel<true><<<grid, 128, smem, s>>>(x, x_cs, w, styles, bias, colors, clamp, img, uvs, canvas, alpha, N, C, HW);
    } else {
        torgb_canvas_kernel<false><<<grid, 128, smem, s>>>(x, 0, w, styles, bias, colors, clamp, img, uvs, canvas, alpha, N, C, HW);
    }
    return launched("torgb_canvas_kernel");
}

extern "C" int nbe_blend_features(void* x, const void* saved, const float* alpha, int64_t alpha_sn, int N, int C, int H, int W,
                                  int is_nhwc_bf16, int cs, nbe_stream_t stream) {
    NBE_REQUIRE(x && saved && alpha && N >= 0 && C >= 1 && H >= 1 && W >= 1, "blend_features: bad arguments");
    if (N == 0) return NBE_OK;
    const int64_t total = (int64_t)N * C * H * W;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
    cudaStream_t s = (cudaStream_t)stream;
    if (is_nhwc_bf16) {
        NBE_REQUIRE(cs >= C, "blend_features: bad channel stride");
        blend_nhwc_kernel<<<(int)blocks, 256, 0, s>>>((__nv_bfloat16*)x, (const __nv_bfloat16*)saved, alpha, alpha_sn, C, H * W, cs, total);
    } else {
        blend_nchw_kernel<<<(int)blocks, 256, 0, s>>>((float*)x, (const float*)saved, alpha, alpha_sn, C, H * W, total);
    }
    return launched("blend_features_kernel");
}

// ---------------------------------------------------------------------------------------------
// All per-layer affines + demodulation coefficients of one forward pass in ONE launch, and all shifted-noise maps in
// another: at batch 256 the ~40 tiny per-layer launches they replace cost ~0.5 ms of pure launch/drain latency.
// ---------------------------------------------------------------------------------------------
namespace nbe {

constexpr int SD_MAX_LAYERS = 24;

struct StylesTable {
    int n_layers;
    const float* affine_w[SD_MAX_LAYERS];     // [cin, w_dim]
    const float* affine_b[SD_MAX_LAYERS];     // [cin]
    const float* wsq[SD_MAX_LAYERS];          // [cout, cin] or NULL (no demodulation)
    float* styles[SD_MAX_LAYERS];             // [N, cin]
    float* dcoef[SD_MAX_LAYERS];              // [N, cout] or NULL
    int cin[SD_MAX_LAYERS], cout[SD_MAX_LAYERS], w_index[SD_MAX_LAYERS];
    float post_scale[SD_MAX_LAYERS];          // styles *= post_scale for channels >= post_from (ToRGB: 1/sqrt(C) after the 9 colour outputs)
    int post_from[SD_MAX_LAYERS];
    // optional: the blocks of layer in_layer also write the first block's modulated constant input,
    // in_out[n, y, x, c] = bf16(in_const[y, x, c] * styles[n, c]) (networks.py:642-643 `const` + networks.py:68 `x * styles`)
    int in_layer;                             // -1: none
    const float* in_const;                    // [in_h, in_w, cin]
    __nv_bfloat16* in_out;                    // [N, in_h, in_pitch, cin], columns >= in_w are left alone (the zero gap)
    int in_h, in_w, in_pitch;
};

// grid = (ceil(N / SD_NB), n_layers); block = SD_THREADS threads.  styles = affine(w) (FullyConnectedLayer, lr 1, bias_init 1:
// networks.py:109-122), d[o] = rsqrt(sum_i styles[i]^2 wsq[o,i] + 1e-8) (networks.py:59-64).  A block handles SD_NB samples
// so that every affine / wsq weight it loads feeds SD_NB FMAs (the weights are re-read by every block: L2-bound otherwise).
constexpr int SD_NB = 8;
constexpr int SD_THREADS = 256;
static_assert(SD_NB == 8, "the affine inner loop is written for 8 samples");

__global__ void __launch_bounds__(SD_THREADS)
styles_demod_kernel(const float* __restrict__ ws, int N, int num_ws, int w_dim, const StylesTable tab) {
    extern __shared__ __align__(16) float s_sd[];                              // [SD_NB][w_dim] latents, then [SD_NB][cin] styles^2
    const int n0 = blockIdx.x * SD_NB, l = blockIdx.y;
    const int cin = tab.cin[l], cout = tab.cout[l];
    float* s_w = s_sd;
    float* s_s2 = s_sd + SD_NB * w_dim;
    // latents transposed to [w_dim][SD_NB]: the SD_NB values that meet one weight are two 16-byte broadcast loads
    for (int i = threadIdx.x; i < SD_NB * w_dim; i += SD_THREADS) {
        const int nb = i / w_dim, k = i - nb * w_dim;
        const int n = min(n0 + nb, N - 1);                          // tail block: duplicates, never stored
        s_w[k * SD_NB + nb] = ws[((long long)n * num_ws + tab.w_index[l]) * w_dim + k];
    }
    __syncthreads();
    const float wgain = rsqrtf((float)w_dim);
    for (int c = threadIdx.x; c < cin; c += SD_THREADS) {
        const float* wr = tab.affine_w[l] + (long long)c * w_dim;
        float acc[SD_NB];
#pragma unroll
        for (int nb = 0; nb < SD_NB; ++nb) acc[nb] = 0.f;
        auto step = [&](int i, float wv) {
            const float w = wv * wgain;
            const float4 a = *reinterpret_cast<const float4*>(s_w + i * SD_NB);
            const float4 b = *reinterpret_cast<const float4*>(s_w + i * SD_NB + 4);
            acc[0] = fmaf(a.x, w, acc[0]); acc[1] = fmaf(a.y, w, acc[1]); acc[2] = fmaf(a.z, w, acc[2]); acc[3] = fmaf(a.w, w, acc[3]);
            acc[4] = fmaf(b.x, w, acc[4]); acc[5] = fmaf(b.y, w, acc[5]); acc[6] = fmaf(b.z, w, acc[6]); acc[7] = fmaf(b.w, w, acc[7]);
        };
        if ((w_dim & 3) == 0 && (((uintptr_t)wr) & 15) == 0) {
            for (int i = 0; i < w_dim; i += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(wr + i);
                step(i, w4.x); step(i + 1, w4.y); step(i + 2, w4.z); step(i + 3, w4.w);
            }
        } else {
            for (int i = 0; i < w_dim; ++i) step(i, wr[i]);
        }
        const float bias = tab.affine_b[l][c];
        const float post = (c >= tab.post_from[l]) ? tab.post_scale[l] : 1.f;
#pragma unroll
        for (int nb = 0; nb < SD_NB; ++nb) {
            const float a = acc[nb] + bias;
            s_s2[nb * cin + c] = a * a;
            if (n0 + nb < N && blockIdx.z == 0) tab.styles[l][(long long)(n0 + nb) * cin + c] = (c >= tab.post_from[l]) ? a * post : a;
        }
        if (l == tab.in_layer && blockIdx.z == 0) {
            for (int y = 0; y < tab.in_h; ++y)
                for (int x = 0; x < tab.in_w; ++x) {
                    const float cv = tab.in_const[(y * tab.in_w + x) * cin + c];
#pragma unroll
                    for (int nb = 0; nb < SD_NB; ++nb)
                        if (n0 + nb < N)
                            tab.in_out[(((long long)(n0 + nb) * tab.in_h + y) * tab.in_pitch + x) * cin + c] = __float2bfloat16_rn(cv * (acc[nb] + bias));
                }
        }
    }
    __syncthreads();
    if (tab.wsq[l] == nullptr || tab.dcoef[l] == nullptr) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // small batches have few (sample group, layer) blocks: the output channels are then split over gridDim.z blocks, each of
    // which recomputes the (cheap) styles and reads only its slice of wsq
    const int per_z = (cout + (int)gridDim.z - 1) / (int)gridDim.z;
    const int o_end = min(cout, ((int)blockIdx.z + 1) * per_z);
    for (int o = (int)blockIdx.z * per_z + warp; o < o_end; o += SD_THREADS / 32) {
        const float* q = tab.wsq[l] + (long long)o * cin;
        float sm[SD_NB];
#pragma unroll
        for (int nb = 0; nb < SD_NB; ++nb) sm[nb] = 0.f;
        for (int i = lane; i < cin; i += 32) {
            const float qv = q[i];
#pragma unroll
            for (int nb = 0; nb < SD_NB; ++nb) sm[nb] = fmaf(s_s2[nb * cin + i], qv, sm[nb]);
        }
#pragma unroll
        for (int nb = 0; nb < SD_NB; ++nb) {
            float v = sm[nb];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0 && n0 + nb < N) tab.dcoef[l][(long long)(n0 + nb) * cout + o] = rsqrtf(v + 1e-8f);
        }
    }
}

constexpr int SN_MAX_LAYERS = 16;
struct NoiseTable {
    int n_layers;
    const float* noise_const[SN_MAX_LAYERS];
    const float* lin[SN_MAX_LAYERS];
    float* out[SN_MAX_LAYERS];                // [N, R, R]
    int res[SN_MAX_LAYERS];
    long long start[SN_MAX_LAYERS + 1];       // prefix sums of N * tiles(layer): first block of each layer
};

// One block = one 32 x 32 output tile of ONE (layer, sample).  The reference's sampling grid is TRANSPOSED (output row i
// selects the source COLUMN, output column j the source ROW), so a thread-per-output-column mapping reads the noise constant
// down its columns (32 cache lines per load).  Here lanes run along the output ROW index i -- the four taps are then
// (nearly) contiguous reads -- and the tile is transposed through shared memory so that the stores are contiguous too.
// The separable part of the grid (source index + weight per output row / column; two 64-bit modulos and two IEEE divisions
// per sample) is computed once per block by 64 threads.  The float expression order is the one of nbe_shifted_noise_f32 /
// the oracle (grid_sample restated), so results are bit-identical to the per-layer kernel.
constexpr int SN_TILE = 32;

__global__ void __launch_bounds__(256)
shifted_noise_all_kernel(const int64_t* __restrict__ positions, int N, int mod, const NoiseTable tab) {
    __shared__ int s_i0[2][SN_TILE];                                // [0]: x0 of output row i0 + k, [1]: y0 of output column j0 + k
    __shared__ float s_t[2][SN_TILE];                               // the matching fractional weights
    __shared__ float s_tile[SN_TILE][SN_TILE + 1];                  // [i][j]
    // block -> (layer, sample, tile): tab.start[] holds prefix sums of N * tiles(layer)
    int l = 0;
    while ((long long)blockIdx.x >= tab.start[l + 1]) ++l;
    const int R = tab.res[l];
    const int tpr = (R + SN_TILE - 1) / SN_TILE;
    const int b = (int)(blockIdx.x - tab.start[l]);
    const int n = b / (tpr * tpr), t2 = b - n * tpr * tpr;
    const int i0 = (t2 / tpr) * SN_TILE, j0 = (t2 % tpr) * SN_TILE;
    const float* __restrict__ nc = tab.noise_const[l];
    if (threadIdx.x < 2 * SN_TILE) {
        const int which = threadIdx.x >= SN_TILE, kk = threadIdx.x & (SN_TILE - 1);
        const int k = (which ? j0 : i0) + kk;
        if (k < R) {
            int64_t pp = positions[2 * n + which] % mod;
            if (pp < 0) pp += mod;
            const float pf = __fdiv_rn((float)pp, (float)(mod - 1));
            float sv = __fadd_rn(tab.lin[l][k], pf); sv = __fsub_rn(sv, floorf(sv));
            const float g = __fsub_rn(__fmul_rn(sv, 2.f), 1.f);
            const float c = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(R - 1));   // x / 2 == x * 0.5 exactly
            const float f0 = floorf(c);
            s_i0[which][kk] = (int)f0;
            s_t[which][kk] = c - f0;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (i0 + lane < R) {
        const int x0 = s_i0[0][lane], x1 = x0 + 1;
        const float tx = s_t[0][lane];
#pragma unroll
        for (int q = 0; q < SN_TILE / 8; ++q) {
            const int jj = warp + q * 8;
            if (j0 + jj < R) {
                const int y0 = s_i0[1][jj], y1 = y0 + 1;
                const float ty = s_t[1][jj];
                auto at = [&](int yy, int xx) -> float { return (yy >= 0 && yy < R && xx >= 0 && xx < R) ? __ldg(nc + yy * R + xx) : 0.f; };
                s_tile[lane][jj] = at(y0, x0) * (1.f - tx) * (1.f - ty) + at(y0, x1) * tx * (1.f - ty) +
                                   at(y1, x0) * (1.f - tx) * ty + at(y1, x1) * tx * ty;
            }
        }
    }
    __syncthreads();
    float* out = tab.out[l] + (long long)n * R * R;
    if (j0 + lane < R) {
#pragma unroll
        for (int q = 0; q < SN_TILE / 8; ++q) {
            const int ii = warp + q * 8;
            if (i0 + ii < R) out[(i0 + ii) * R + j0 + lane] = s_tile[ii][lane];
        }
    }
}

}  // namespace nbe

extern "C" int nbe_styles_demod_f32(const float* ws, int N, int num_ws, int w_dim, int n_layers,
                                    const void* const* affine_w, const void* const* affine_b, const void* const* wsq,
                                    void* const* styles, void* const* dcoef, const int* cin, const int* cout, const int* w_index,
                                    const float* post_scale, const int* post_from, nbe_stream_t stream) {
    return nbe_styles_demod_input_f32(ws, N, num_ws, w_dim, n_layers, affine_w, affine_b, wsq, styles, dcoef, cin, cout, w_index,
                                      post_scale, post_from, -1, nullptr, nullptr, 0, 0, 0, stream);
}

extern "C" int nbe_styles_demod_input_f32(const float* ws, int N, int num_ws, int w_dim, int n_layers,
                                          const void* const* affine_w, const void* const* affine_b, const void* const* wsq,
                                          void* const* styles, void* const* dcoef, const int* cin, const int* cout, const int* w_index,
                                          const float* post_scale, const int* post_from,
                                          int in_layer, const float* in_const, void* in_out, int in_h, int in_w, int in_pitch,
                                          nbe_stream_t stream) {
    NBE_REQUIRE(ws && N >= 0 && n_layers >= 1 && n_layers <= SD_MAX_LAYERS && w_dim >= 1 && num_ws >= 1, "styles_demod: bad arguments");
    NBE_REQUIRE(affine_w && affine_b && wsq && styles && dcoef && cin && cout && w_index && post_scale && post_from, "styles_demod: null table");
    if (N == 0) return NBE_OK;
    NBE_REQUIRE(N <= INT32_MAX / 2, "styles_demod: batch too large");
    StylesTable tab;
    tab.n_layers = n_layers;
    int max_cin = 0;
    for (int l = 0; l < n_layers; ++l) {
        NBE_REQUIRE(affine_w[l] && affine_b[l] && styles[l] && cin[l] >= 1 && w_index[l] >= 0 && w_index[l] < num_ws, "styles_demod: bad layer %d", l);
        tab.affine_w[l] = (const float*)affine_w[l]; tab.affine_b[l] = (const float*)affine_b[l]; tab.wsq[l] = (const float*)wsq[l];
        tab.styles[l] = (float*)styles[l]; tab.dcoef[l] = (float*)dcoef[l]; tab.cin[l] = cin[l]; tab.cout[l] = cout[l];
        tab.w_index[l] = w_index[l]; tab.post_scale[l] = post_scale[l]; tab.post_from[l] = post_from[l];
        if (cin[l] > max_cin) max_cin = cin[l];
    }
    tab.in_layer = -1; tab.in_const = nullptr; tab.in_out = nullptr; tab.in_h = tab.in_w = tab.in_pitch = 0;
    if (in_layer >= 0) {
        NBE_REQUIRE(in_layer < n_layers && in_const && in_out && in_h >= 1 && in_w >= 1 && in_pitch >= in_w && ((uintptr_t)in_out & 1) == 0,
                    "styles_demod: bad constant-input arguments");
        NBE_REQUIRE(post_scale[in_layer] == 1.f || post_from[in_layer] >= cin[in_layer], "styles_demod: the constant input is modulated by un-scaled styles (layer %d has a post scale)", in_layer);
        tab.in_layer = in_layer; tab.in_const = in_const; tab.in_out = (__nv_bfloat16*)in_out; tab.in_h = in_h; tab.in_w = in_w; tab.in_pitch = in_pitch;
    }
    const int groups = (N + SD_NB - 1) / SD_NB;
    dim3 grid(groups, n_layers, groups * n_layers < 2 * kNumSMs ? 8 : 1);
    const size_t smem = (size_t)SD_NB * (w_dim + max_cin) * sizeof(float);
    NBE_REQUIRE(smem <= 48 * 1024, "styles_demod: layer too wide");
    styles_demod_kernel<<<grid, SD_THREADS, smem, (cudaStream_t)stream>>>(ws, N, num_ws, w_dim, tab);
    return launched("styles_demod_kernel");
}

extern "C" int nbe_shifted_noise_all_f32(const int64_t* positions, int N, int mod, int n_layers,
                                         const void* const* noise_const, const void* const* lin, void* const* out, const int* res,
                                         nbe_stream_t stream) {
    NBE_REQUIRE(positions && N >= 0 && mod >= 2 && n_layers >= 1 && n_layers <= SN_MAX_LAYERS && noise_const && lin && out && res, "shifted_noise_all: bad arguments");
    if (N == 0) return NBE_OK;
    NoiseTable tab;
    tab.n_layers = n_layers;
    tab.start[0] = 0;
    for (int l = 0; l < n_layers; ++l) {
        NBE_REQUIRE(noise_const[l] && lin[l] && out[l] && res[l] >= 2 , "shifted_noise_all: bad layer %d", l);
        tab.noise_const[l] = (const float*)noise_const[l]; tab.lin[l] = (const float*)lin[l]; tab.out[l] = (float*)out[l]; tab.res[l] = res[l];
        NBE_REQUIRE((long long)N * res[l] * res[l] <= INT32_MAX, "shifted_noise_all: layer %d too large", l);
        { const long long tpr = (res[l] + SN_TILE - 1) / SN_TILE; tab.start[l + 1] = tab.start[l] + (long long)N * tpr * tpr; }
    }
    NBE_REQUIRE(tab.start[n_layers] <= INT32_MAX, "shifted_noise_all: too many blocks");
    shifted_noise_all_kernel<<<(int)tab.start[n_layers], 256, 0, (cudaStream_t)stream>>>(positions, N, mod, tab);
    return launched("shifted_noise_all_kernel");
}
