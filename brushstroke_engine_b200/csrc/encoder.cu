// Geometry-encoder helpers for the tensor-core path (NHWC bf16, reflect padding kept explicit in the buffers so the
// 3x3 convolutions become *valid* implicit GEMMs for nbe_conv_tc_bf16_ex):
//   * enc_conv7x7_kernel      : first layer, 1 -> 64 channels, 7x7, reflect pad 3, folded BN bias + LeakyReLU (CUDA cores:
//                               K = 49 is too thin for the tensor pipe), writes the interior of a 1-px padded NHWC buffer
//   * reflect_border_kernel   : fills the 1-px border of a padded NHWC buffer by reflection (torch 'reflect' semantics)
//   * bilinear2x_pad_kernel   : bilinear x2 (align_corners = True) + reflect pad 1 in one pass
// Reference: forger/experimental/autoenc/simple_autoencoder.py:95-126 (SingleConvolution / ScaleUp), 155-199, 251-261.
#include "common.cuh"
#include <algorithm>

namespace nbe {

__device__ __forceinline__ int reflect_idx(int i, int n) {        // torch 'reflect': -1 -> 1, n -> n-2
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

constexpr int E7_T = 16;                                             // 16 x 16 output pixels per CTA

__global__ void __launch_bounds__(256)
enc_conv7x7_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   __nv_bfloat16* __restrict__ y, int H, int W, int Cout, int y_cs, float neg_slope, int preproc) {
    __shared__ float s_in[E7_T + 6][E7_T + 6 + 1];
    extern __shared__ float s_w[];                                   // [49][Cout] then bias[Cout]
    float* s_b = s_w + 49 * Cout;
    const int n = blockIdx.z, ty0 = blockIdx.y * E7_T, tx0 = blockIdx.x * E7_T;
    for (int i = threadIdx.x; i < 49 * Cout; i += 256) {
        const int t = i / Cout, o = i - t * Cout;
        s_w[i] = w[o * 49 + t];
    }
    for (int i = threadIdx.x; i < Cout; i += 256) s_b[i] = bias[i];
    for (int i = threadIdx.x; i < (E7_T + 6) * (E7_T + 6); i += 256) {
        const int r = i / (E7_T + 6), c = i - r * (E7_T + 6);
        const int iy = reflect_idx(ty0 + r - 3, H), ix = reflect_idx(tx0 + c - 3, W);
        float v = x[((long long)n * H + iy) * W + ix];
        if (preproc == 1) v = 1.f - v;                               // 'inverse'      (base.py:46-49)
        else if (preproc == 2) v = (1.f - v) * 2.f - 1.f;            // '-11inverse'   (base.py:42-45)
        s_in[r][c] = v;
    }
    __syncthreads();
    const int ly = threadIdx.x >> 4, lx = threadIdx.x & 15;
    const int oy = ty0 + ly, ox = tx0 + lx;
    float in[49];
#pragma unroll
    for (int kh = 0; kh < 7; ++kh)
#pragma unroll
        for (int kw = 0; kw < 7; ++kw) in[kh * 7 + kw] = s_in[ly + kh][lx + kw];
    if (oy >= H || ox >= W) return;
    // output goes to the interior of a [N, H+2, W+2, y_cs] buffer
    __nv_bfloat16* yp = y + (((long long)n * (H + 2) + oy + 1) * (W + 2) + ox + 1) * y_cs;
    for (int o0 = 0; o0 < Cout; o0 += 8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = s_b[o0 + j];
#pragma unroll
        for (int t = 0; t < 49; ++t) {
            const float4 w0 = *reinterpret_cast<const float4*>(&s_w[t * Cout + o0]);
            const float4 w1 = *reinterpret_cast<const float4*>(&s_w[t * Cout + o0 + 4]);
            acc[0] = fmaf(in[t], w0.x, acc[0]); acc[1] = fmaf(in[t], w0.y, acc[1]);
            acc[2] = fmaf(in[t], w0.z, acc[2]); acc[3] = fmaf(in[t], w0.w, acc[3]);
            acc[4] = fmaf(in[t], w1.x, acc[4]); acc[5] = fmaf(in[t], w1.y, acc[5]);
            acc[6] = fmaf(in[t], w1.z, acc[6]); acc[7] = fmaf(in[t], w1.w, acc[7]);
        }
        int4 out;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = acc[2 * j], b = acc[2 * j + 1];
            a = a > 0.f ? a : a * neg_slope;
            b = b > 0.f ? b : b * neg_slope;
            o2[j] = __floats2bfloat162_rn(a, b);
        }
        *reinterpret_cast<int4*>(yp + o0) = out;
    }
}

// One thread per (border pixel, 8-channel vector).  Border pixels of an Hp x Wp image: 2*Wp + 2*(Hp-2).
__global__ void __launch_bounds__(256)
reflect_border_kernel(__nv_bfloat16* __restrict__ buf, int N, int Hp, int Wp, int C, int cs) {
    pdl_trigger();
    pdl_wait();
    const int CV = C / 8;
    const int nb = 2 * Wp + 2 * (Hp - 2);
    const long long total = (long long)N * nb * CV;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        long long t = idx / CV;
        const int b = (int)(t % nb);
        const int n = (int)(t / nb);
        int py, px;
        if (b < Wp) { py = 0; px = b; }
        else if (b < 2 * Wp) { py = Hp - 1; px = b - Wp; }
        else { const int r = b - 2 * Wp; py = 1 + (r >> 1); px = (r & 1) ? Wp - 1 : 0; }
        // interior coordinates are padded coordinates - 1; reflect in interior space
        const int sy = reflect_idx(py - 1, Hp - 2) + 1, sx = reflect_idx(px - 1, Wp - 2) + 1;
        const int4 v = *reinterpret_cast<const int4*>(buf + (((long long)n * Hp + sy) * Wp + sx) * cs + cv * 8);
        *reinterpret_cast<int4*>(buf + (((long long)n * Hp + py) * Wp + px) * cs + cv * 8) = v;
    }
}

// out[n, py, px, :] (padded (2h+2) x (2w+2)) = bilinear_x2_align_corners(x)[reflect(py-1), reflect(px-1)]
__global__ void __launch_bounds__(256)
bilinear2x_pad_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int h, int w, int C,
                      int xs_c, int out_cs) {
    pdl_trigger();
    pdl_wait();
    const int CV = C / 8;
    const int Hp = 2 * h + 2, Wp = 2 * w + 2;
    const long long total = (long long)N * Hp * Wp * CV;
    const float sy_scale = (float)(h - 1) / (float)(2 * h - 1), sx_scale = (float)(w - 1) / (float)(2 * w - 1);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        long long t = idx / CV;
        const int px = (int)(t % Wp); t /= Wp;
        const int py = (int)(t % Hp);
        const int n = (int)(t / Hp);
        const int Y = reflect_idx(py - 1, 2 * h), X = reflect_idx(px - 1, 2 * w);
        const float fy = sy_scale * (float)Y, fx = sx_scale * (float)X;      // ATen: scale * dst_index (align_corners)
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        const __nv_bfloat16* base = x + (long long)n * h * w * xs_c + cv * 8;
        const int4 r00 = *reinterpret_cast<const int4*>(base + ((long long)y0 * w + x0) * xs_c);
        const int4 r01 = *reinterpret_cast<const int4*>(base + ((long long)y0 * w + x1) * xs_c);
        const int4 r10 = *reinterpret_cast<const int4*>(base + ((long long)y1 * w + x0) * xs_c);
        const int4 r11 = *reinterpret_cast<const int4*>(base + ((long long)y1 * w + x1) * xs_c);
        const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&r00);
        const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&r01);
        const __nv_bfloat162* c = reinterpret_cast<const __nv_bfloat162*>(&r10);
        const __nv_bfloat162* d = reinterpret_cast<const __nv_bfloat162*>(&r11);
        int4 o;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 va = __bfloat1622float2(a[k]), vb = __bfloat1622float2(b[k]);
            const float2 vc = __bfloat1622float2(c[k]), vd = __bfloat1622float2(d[k]);
            o2[k] = __floats2bfloat162_rn(w00 * va.x + w01 * vb.x + w10 * vc.x + w11 * vd.x,
                                          w00 * va.y + w01 * vb.y + w10 * vc.y + w11 * vd.y);
        }
        *reinterpret_cast<int4*>(out + (((long long)n * Hp + py) * Wp + px) * out_cs + cv * 8) = o;
    }
}

// ---- FP32 parity mode (NCHW float32) ---------------------------------------------------------------------
// y[nc, i, j] = x[nc, r(i - pad, H), r(j - pad, W)],  r = reflection without repeating the edge (padding_mode='reflect',
// simple_autoencoder.py:98), optionally of the bilinear x2 up-sampling of x (align_corners=True, ScaleUp :117): the
// up-sampled map is never materialised.
__global__ void __launch_bounds__(256)
reflect_pad_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, long long NC, int H, int W, int pad, int up) {
    const int SH = up ? 2 * H : H, SW = up ? 2 * W : W;            // size of the (virtual) source map
    const int OH = SH + 2 * pad, OW = SW + 2 * pad;
    const long long total = NC * OH * OW;
    const float sy = (SH > 1) ? (float)(H - 1) / (float)(SH - 1) : 0.f, sx = (SW > 1) ? (float)(W - 1) / (float)(SW - 1) : 0.f;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % OW);
        const long long t = idx / OW;
        const int i = (int)(t % OH);
        const long long nc = t / OH;
        int u = i - pad, v = j - pad;
        u = u < 0 ? -u : u; u = u >= SH ? 2 * SH - 2 - u : u;
        v = v < 0 ? -v : v; v = v >= SW ? 2 * SW - 2 - v : v;
        const float* xp = x + nc * H * W;
        float r;
        if (!up) r = xp[(long long)u * W + v];
        else {
            // torch's upsample_bilinear2d, align_corners=True: src = dst * (in - 1) / (out - 1)
            const float fy = sy * (float)u, fx = sx * (float)v;
            const int y0 = (int)fy, x0 = (int)fx;
            const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
            const float ly = fy - (float)y0, lx = fx - (float)x0;
            const float hy = 1.f - ly, hx = 1.f - lx;
            r = hy * (hx * xp[(long long)y0 * W + x0] + lx * xp[(long long)y0 * W + x1]) +
                ly * (hx * xp[(long long)y1 * W + x0] + lx * xp[(long long)y1 * W + x1]);
        }
        y[idx] = r;
    }
}

// y[n, p, c] = (x[n, p, c] * scale[c] + shift[c]) * next_scale[n, c]: eval-mode BatchNorm AFTER the activation
// (simple_autoencoder.py:100-103, the --neg_slope variant) for the feature maps that leave the encoder, optionally times the
// consuming generator layer's styles.  8 channels (16 bytes) per thread; source and destination have their own pitches.
__global__ void __launch_bounds__(256)
affine_nhwc_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, long long x_row_pitch, long long x_img_pitch,
                   __nv_bfloat16* __restrict__ y, int y_cs, long long y_row_pitch, long long y_img_pitch,
                   int N, int H, int W, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                   const float* __restrict__ next_scale) {
    const int CV = C / 8;
    const long long total = (long long)N * H * W * CV;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        long long t = idx / CV;
        const int px = (int)(t % W); t /= W;
        const int py = (int)(t % H);
        const int n = (int)(t / H);
        int4 v = *reinterpret_cast<const int4*>(x + ((long long)n * x_img_pitch + (long long)py * x_row_pitch + px) * x_cs + cv * 8);
        __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(&v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = cv * 8 + k;
            float a = fmaf(__bfloat162float(e[k]), scale[c], shift[c]);
            if (next_scale) a *= next_scale[(long long)n * C + c];
            e[k] = __float2bfloat16_rn(a);
        }
        *reinterpret_cast<int4*>(y + ((long long)n * y_img_pitch + (long long)py * y_row_pitch + px) * y_cs + cv * 8) = v;
    }
}

__global__ void __launch_bounds__(256)
affine_nchw_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long total, int C, int HW,
                       const float* __restrict__ scale, const float* __restrict__ shift) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((idx / HW) % C);
        y[idx] = x[idx] * scale[c] + shift[c];                    // two roundings, like torch's batch_norm in eval mode restated as scale / shift
    }
}

}  // namespace nbe

using namespace nbe;

extern "C" int nbe_enc_conv7x7_bf16(const float* x, const float* w, const float* bias, void* y, int N, int H, int W, int Cout,
                                    int y_cs, float neg_slope, int preproc, nbe_stream_t stream) {
    NBE_REQUIRE(x && w && bias && y && N >= 0 && H >= 4 && W >= 4, "enc_conv7x7: bad arguments");
    NBE_REQUIRE(Cout % 8 == 0 && Cout <= 128 && y_cs % 8 == 0 && y_cs >= Cout, "enc_conv7x7: Cout must be a multiple of 8 (<= 128)");
    NBE_REQUIRE(preproc >= 0 && preproc <= 2, "enc_conv7x7: unknown preprocessing %d", preproc);
    if (N == 0) return NBE_OK;
    NBE_REQUIRE(N <= 65535, "enc_conv7x7: batch too large for one launch");
    dim3 grid((W + E7_T - 1) / E7_T, (H + E7_T - 1) / E7_T, N);
    const size_t smem = (size_t)(49 + 1) * Cout * sizeof(float);
    enc_conv7x7_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, w, bias, (__nv_bfloat16*)y, H, W, Cout, y_cs, neg_slope, preproc);
    return launched("enc_conv7x7_kernel");
}

extern "C" int nbe_reflect_border_nhwc_bf16(void* buf, int N, int Hp, int Wp, int C, int cs, nbe_stream_t stream) {
    NBE_REQUIRE(buf && N >= 0 && Hp >= 4 && Wp >= 4 && C >= 8 && C % 8 == 0 && cs % 8 == 0 && cs >= C, "reflect_border: bad arguments");
    if (N == 0) return NBE_OK;
    const long long total = (long long)N * (2 * Wp + 2 * (Hp - 2)) * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
    launch_pdl(reflect_border_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, (__nv_bfloat16*)buf, N, Hp, Wp, C, cs);
    return launched("reflect_border_kernel");
}

extern "C" int nbe_bilinear2x_pad_nhwc_bf16(const void* x, void* out, int N, int h, int w, int C, int xs_c, int out_cs,
                                            nbe_stream_t stream) {
    NBE_REQUIRE(x && out && N >= 0 && h >= 2 && w >= 2 && C >= 8 && C % 8 == 0, "bilinear2x_pad: bad arguments");
    NBE_REQUIRE(xs_c % 8 == 0 && xs_c >= C && out_cs % 8 == 0 && out_cs >= C, "bilinear2x_pad: bad channel strides");
    NBE_REQUIRE((((uintptr_t)x | (uintptr_t)out) & 15) == 0, "bilinear2x_pad: tensors must be 16-byte aligned");
    if (N == 0) return NBE_OK;
    const long long total = (long long)N * (2 * h + 2) * (2 * w + 2) * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
    launch_pdl(bilinear2x_pad_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)out, N, h, w, C, xs_c, out_cs);
    return launched("bilinear2x_pad_kernel");
}

extern "C" int nbe_reflect_pad_nchw_f32(const float* x, float* y, int64_t NC, int H, int W, int pad, int upsample2x,
                                        nbe_stream_t stream) {
    NBE_REQUIRE(x && y && NC >= 0 && H >= 1 && W >= 1 && pad >= 0, "reflect_pad_nchw: bad arguments");
    const int SH = upsample2x ? 2 * H : H, SW = upsample2x ? 2 * W : W;
    NBE_REQUIRE(pad < SH && pad < SW, "reflect_pad_nchw: padding %d must be smaller than the map (%d x %d)", pad, SH, SW);
    if (NC == 0) return NBE_OK;
    const int64_t total = NC * (SH + 2 * pad) * (SW + 2 * pad);
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
    reflect_pad_nchw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, NC, H, W, pad, upsample2x ? 1 : 0);
    return launched("reflect_pad_nchw_kernel");
}

extern "C" int nbe_affine_nhwc_bf16(const void* x, int x_cs, int64_t x_row_pitch, int64_t x_img_pitch,
                                    void* y, int y_cs, int64_t y_row_pitch, int64_t y_img_pitch,
                                    int N, int H, int W, int C, const float* scale, const float* shift, const float* next_scale,
                                    nbe_stream_t stream) {
    NBE_REQUIRE(x && y && scale && shift && N >= 0 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "affine_nhwc: bad arguments");
    NBE_REQUIRE(x_cs % 8 == 0 && x_cs >= C && y_cs % 8 == 0 && y_cs >= C, "affine_nhwc: channel strides must be multiples of 8");
    NBE_REQUIRE(x_row_pitch >= W && x_img_pitch >= x_row_pitch * H && y_row_pitch >= W && y_img_pitch >= y_row_pitch * H, "affine_nhwc: bad pitches");
    NBE_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "affine_nhwc: tensors must be 16-byte aligned");
    if (N == 0) return NBE_OK;
    const long long total = (long long)N * H * W * (C / 8);
    const long long blocks = std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
    affine_nhwc_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, x_cs, x_row_pitch, x_img_pitch,
        (__nv_bfloat16*)y, y_cs, y_row_pitch, y_img_pitch, N, H, W, C, scale, shift, next_scale);
    return launched("affine_nhwc_kernel");
}

extern "C" int nbe_affine_nchw_f32(const float* x, float* y, int N, int C, int HW, const float* scale, const float* shift,
                                   nbe_stream_t stream) {
    NBE_REQUIRE(x && y && scale && shift && N >= 0 && C >= 1 && HW >= 1, "affine_nchw: bad arguments");
    if (N == 0) return NBE_OK;
    const long long total = (long long)N * C * HW;
    const long long blocks = std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
    affine_nchw_f32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, total, C, HW, scale, shift);
    return launched("affine_nchw_f32_kernel");
}
