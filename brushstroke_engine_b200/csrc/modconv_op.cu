// modulated_conv2d behind the reference's operator signature, as ONE C entry point
// (SG2/training/networks.py:31-88: modulated_conv2d(x, weight, styles, noise, up, down, padding, resample_filter,
//  demodulate, flip_weight, fused_modconv); the convolution itself is conv2d_resample, SG2/torch_utils/ops/conv2d_resample.py:59-154).
//
// The operator contract is NCHW in / NCHW out in the dtype of x.  For 16-bit activations (the reference's fp16 layers; bf16
// here) the entry point runs the tensor-core kernels of the generator's fast path:
//   pack   NCHW T -> zero-gapped NHWC bf16, multiplied by styles[n,i]          (input-side modulation, networks.py:67)
//   up=1   nbe_conv3x3_flat_bf16 (tcgen05 CTA pairs)  with dcoef / noise in its epilogue
//   up=2   nbe_convT3x3s2_flat_bf16 (transposed conv at 9 taps per INPUT pixel) + nbe_fir_act_nhwc_bf16 (4x4 FIR, gain 4,
//          dcoef, noise) -- conv2d_resample.py:124-142
//   unpack NHWC bf16 -> NCHW T
// and for float32 activations the true-FP32 direct convolution nbe_conv2d_f32 (FIR-first for up = 2, conv2d_resample.py:149-154),
// which is what the <= 1e-4 parity mode needs.  All scratch memory (weight squares, demodulation coefficients, repacked
// weights, NHWC staging) lives in ONE caller-provided workspace whose size nbe_modulated_conv2d_workspace() reports.
#include "common.cuh"
#include <algorithm>

namespace nbe {

static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

// ---- NCHW 16-bit -> zero-gapped NHWC bf16 -------------------------------------------------------------------------
// dst[((n*H + h)*P + w)*cs + c] = src[n,c,h,w] * scale[n,c]   for w < W, c < C;   0 for gap columns W <= w < P and padding
// channels C <= c < cs (so the buffer never has to be cleared).  One 64 (channels) x 64 (columns) tile per CTA through
// shared memory: global reads are 8-byte vectors along w (128 contiguous bytes per channel row), global writes 16-byte
// vectors along c (128 contiguous bytes per pixel); the transposition itself is done with 2-byte shared-memory accesses
// (33-word row pitch: at most 2-way bank conflicts).  HBM-bound: 2 bytes read + 2 bytes written per element.
constexpr int PK_T = 64;                                            // tile edge
constexpr int PK_PITCH = PK_T + 2;                                  // elements per shared-memory row (33 words)

template <class T> struct Pair16;
template <> struct Pair16<__nv_bfloat16> {
    static __device__ __forceinline__ float2 unpack(uint32_t w) { return bf16x2_to_f2(w); }
    static __device__ __forceinline__ uint32_t pack(float a, float b) { const __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&v); }
};
template <> struct Pair16<__half> {
    static __device__ __forceinline__ float2 unpack(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
    static __device__ __forceinline__ uint32_t pack(float a, float b) { const __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&v); }
};

template <class T>
__global__ void __launch_bounds__(256)
pack_flat_kernel(const T* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int H, int W, int P, int cs,
                 const float* __restrict__ scale, int vec_ok) {
    __shared__ __nv_bfloat16 tile[PK_T][PK_PITCH];                  // [channel][column]
    const int nh = blockIdx.z, n = nh / H, h = nh - n * H;
    const int c0 = blockIdx.y * PK_T, w0 = blockIdx.x * PK_T;
    const int t = threadIdx.x;
    {   // ---- load: thread = (channel row r of 16, group of 4 columns)
        // (pointers advance by constant strides: the 64-bit index arithmetic per pass was a quarter of this kernel's instructions,
        // ncu: 24 instructions per element, issue slots 64 % busy at 46 % of the DRAM throughput)
        const int r = t >> 4, g = (t & 15) * 4;
        const int w = w0 + g;
        const T* sp = src + (((int64_t)n * C + c0 + r) * H + h) * W + w;
        const int64_t sp_step = (int64_t)16 * H * W;
        const float* scp = scale ? scale + (int64_t)n * C + c0 + r : nullptr;
#pragma unroll
        for (int pass = 0; pass < 4; ++pass, sp += sp_step) {
            const int cl = r + 16 * pass, c = c0 + cl;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (c < C && w < W) {
                if (vec_ok && w + 3 < W) {
                    const uint2 raw = *reinterpret_cast<const uint2*>(sp);
                    const float2 a = Pair16<T>::unpack(raw.x), b = Pair16<T>::unpack(raw.y);
                    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (w + k < W) v[k] = Cvt<T>::ld(sp[k]);
                }
                if (scale) {
                    const float sc = scp[16 * pass];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] *= sc;
                }
            }
            *reinterpret_cast<__nv_bfloat162*>(&tile[cl][g]) = __floats2bfloat162_rn(v[0], v[1]);
            *reinterpret_cast<__nv_bfloat162*>(&tile[cl][g + 2]) = __floats2bfloat162_rn(v[2], v[3]);
        }
    }
    __syncthreads();
    {   // ---- store: thread = (column of 32, group of 8 channels) -> one 16-byte vector
        const int cg = (t & 7) * 8, pl = t >> 3;
        const int c = c0 + cg;
        __nv_bfloat16* dp = dst + (((int64_t)n * H + h) * P + w0 + pl) * cs + c;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass, dp += 32 * cs) {
            const int wl = pl + 32 * pass, w = w0 + wl;
            if (w < P && c < cs) {
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned short lo = *reinterpret_cast<const unsigned short*>(&tile[cg + 2 * k][wl]);
                    const unsigned short hi = *reinterpret_cast<const unsigned short*>(&tile[cg + 2 * k + 1][wl]);
                    o[k] = (uint32_t)lo | ((uint32_t)hi << 16);
                }
                *reinterpret_cast<uint4*>(dp) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        // the tiles cover the W image columns; gap columns the last tile does not reach (W a multiple of 64: the gap would otherwise
        // cost a whole extra tile per row -- a third of all CTAs at 128^2) are zeroed by the last tile's threads
        if (blockIdx.x == gridDim.x - 1) {
            const int g0 = w0 + PK_T;                                // first column this tile's 64 columns do not cover
            const int wz = g0 + (t >> 3);
            if (wz < P && c < cs)
                *reinterpret_cast<uint4*>(dst + (((int64_t)n * H + h) * P + wz) * cs + c) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// ---- NHWC bf16 (row pitch P pixels, channel stride cs) -> NCHW 16-bit: the same tile, the other way round -------------------
template <class T>
__global__ void __launch_bounds__(256)
unpack_flat_kernel(const __nv_bfloat16* __restrict__ src, T* __restrict__ dst, int C, int H, int W, int P, int cs, int vec_ok) {
    __shared__ __nv_bfloat16 tile[PK_T][PK_PITCH];                  // [channel][column]
    const int nh = blockIdx.z, n = nh / H, h = nh - n * H;
    const int c0 = blockIdx.y * PK_T, w0 = blockIdx.x * PK_T;
    const int t = threadIdx.x;
    {
        const int cg = (t & 7) * 8, pl = t >> 3;
        const int c = c0 + cg;
        const __nv_bfloat16* sp = src + (((int64_t)n * H + h) * P + w0 + pl) * cs + c;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass, sp += 32 * cs) {
            const int wl = pl + 32 * pass, w = w0 + wl;
            uint4 raw = make_uint4(0u, 0u, 0u, 0u);
            if (w < W && c < C) {                                   // C % 8 == 0 on this path: whole vectors
                raw = *reinterpret_cast<const uint4*>(sp);
            }
            const uint32_t o[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                *reinterpret_cast<unsigned short*>(&tile[cg + 2 * k][wl]) = (unsigned short)(o[k] & 0xffffu);
                *reinterpret_cast<unsigned short*>(&tile[cg + 2 * k + 1][wl]) = (unsigned short)(o[k] >> 16);
            }
        }
    }
    __syncthreads();
    {
        const int r = t >> 4, g = (t & 15) * 4;
        const int w = w0 + g;
        T* dp = dst + (((int64_t)n * C + c0 + r) * H + h) * W + w;
        const int64_t dp_step = (int64_t)16 * H * W;
#pragma unroll
        for (int pass = 0; pass < 4; ++pass, dp += dp_step) {
            const int cl = r + 16 * pass, c = c0 + cl;
            if (c < C && w < W) {
                const float2 a = bf16x2_to_f2(*reinterpret_cast<const uint32_t*>(&tile[cl][g]));
                const float2 b = bf16x2_to_f2(*reinterpret_cast<const uint32_t*>(&tile[cl][g + 2]));
                if (vec_ok && w + 3 < W) {
                    *reinterpret_cast<uint2*>(dp) = make_uint2(Pair16<T>::pack(a.x, a.y), Pair16<T>::pack(b.x, b.y));
                } else {
                    const float v[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (w + k < W) dp[k] = Cvt<T>::st(v[k]);
                }
            }
        }
    }
}

template <class T>
static int launch_pack(const void* x, void* dst, int N, int C, int H, int W, int P, int cs, const float* scale, cudaStream_t s) {
    NBE_REQUIRE((int64_t)N * H <= 65535 * 1024LL, "modulated_conv2d: too many rows");
    // grid.z carries n*H (<= 65535): split the batch if needed
    const int max_n = std::max(1, 65535 / H);
    for (int n0 = 0; n0 < N; n0 += max_n) {
        const int nn = std::min(max_n, N - n0);
        NBE_REQUIRE(P - (W + PK_T - 1) / PK_T * PK_T <= 32, "modulated_conv2d: row pitch too large for the packing kernel");
        dim3 grid((W + PK_T - 1) / PK_T, (cs + PK_T - 1) / PK_T, nn * H);
        const T* xp = (const T*)x + (int64_t)n0 * C * H * W;
        const int vec_ok = (W % 4 == 0) && (((uintptr_t)xp & 7) == 0);
        pack_flat_kernel<T><<<grid, 256, 0, s>>>(xp, (__nv_bfloat16*)dst + (int64_t)n0 * H * P * cs,
                                                  C, H, W, P, cs, scale ? scale + (int64_t)n0 * C : nullptr, vec_ok);
        int st = launched("pack_flat_kernel");
        if (st) return st;
    }
    return NBE_OK;
}

template <class T>
static int launch_unpack(const void* src, void* y, int N, int C, int H, int W, int P, int cs, cudaStream_t s) {
    const int max_n = std::max(1, 65535 / H);
    for (int n0 = 0; n0 < N; n0 += max_n) {
        const int nn = std::min(max_n, N - n0);
        dim3 grid((W + PK_T - 1) / PK_T, (C + PK_T - 1) / PK_T, nn * H);
        T* yp = (T*)y + (int64_t)n0 * C * H * W;
        const int vec_ok = (W % 4 == 0) && (((uintptr_t)yp & 7) == 0);
        unpack_flat_kernel<T><<<grid, 256, 0, s>>>((const __nv_bfloat16*)src + (int64_t)n0 * H * P * cs, yp, C, H, W, P, cs, vec_ok);
        int st = launched("unpack_flat_kernel");
        if (st) return st;
    }
    return NBE_OK;
}

struct McPlan {
    bool tc;                       // tensor-core path (16-bit activations, 3x3, Cout % 128 == 0)
    int OH, OW;
    int x_cs, P, in_rows;          // staged input: [N, in_rows, P, x_cs] bf16
    int cin_pad;
    int64_t off_wsq, off_dcoef, off_wq, off_xin, off_t, off_y, total;
};

static bool plan(McPlan& pl, int dtype, int N, int Cin, int H, int W, int Cout, int K, int up, int padding) {
    pl.tc = (dtype == NBE_BF16 || dtype == NBE_F16) && K == 3 && Cout % 128 == 0 && Cin >= 1 &&
            ((up == 1 && (padding == 0 || padding == 1)) || (up == 2 && padding >= 0));
    if (up == 1) { pl.OH = H + 2 * padding - K + 1; pl.OW = W + 2 * padding - K + 1; }
    else         { pl.OH = 2 * H + 2 * padding - 2; pl.OW = 2 * W + 2 * padding - 2; }       // (2H+1) + 2 pad - 3
    if (pl.OH < 1 || pl.OW < 1) return false;
    int64_t o = 0;
    pl.off_wsq = o;   o += align256((int64_t)Cout * Cin * 4);
    pl.off_dcoef = o; o += align256((int64_t)N * Cout * 4);
    if (pl.tc) {
        pl.cin_pad = (Cin + 63) / 64 * 64;
        pl.x_cs = (Cin + 7) / 8 * 8;
        // up = 1, 'same': zero gap column; 'valid' (padding 0): the rows are their own halo.  up = 2: zero gap column.
        pl.P = (up == 1 && padding == 0) ? W : W + 1;
        pl.in_rows = H;
        pl.off_wq = o;  o += align256((int64_t)9 * Cout * pl.cin_pad * 2);
        pl.off_xin = o; o += align256((int64_t)N * H * pl.P * pl.x_cs * 2);
        pl.off_t = o;
        if (up == 2) o += align256((int64_t)N * (2 * H + 1) * (2 * W + 2) * Cout * 2);
        pl.off_y = o;   o += align256((int64_t)N * pl.OH * pl.OW * Cout * 2);
    } else if (dtype == NBE_F32) {
        pl.off_wq = pl.off_xin = pl.off_y = o;
        pl.off_t = o;                                              // FIR-first intermediate U [N,Cin,2H+2p+1.., ..] float32
        if (up == 2) o += align256((int64_t)N * Cin * (pl.OH + K - 1) * (pl.OW + K - 1) * 4);
    } else {
        return false;
    }
    pl.total = o;
    return true;
}

}  // namespace nbe

using namespace nbe;

extern "C" int64_t nbe_modulated_conv2d_workspace(int dtype, int N, int Cin, int H, int W, int Cout, int K, int up, int padding) {
    McPlan pl;
    if (N < 0 || Cin < 1 || H < 1 || W < 1 || Cout < 1 || K < 1 || (up != 1 && up != 2) || padding < 0) return -1;
    if (dtype != NBE_F32 && dtype != NBE_F16 && dtype != NBE_BF16) return -1;
    if (up == 2 && K != 3) return -1;
    if (!plan(pl, dtype, N, Cin, H, W, Cout, K, up, padding)) return -1;
    if (!pl.tc && dtype != NBE_F32) return -1;                      // 16-bit shapes outside the tensor-core kernels: caller converts to float32
    return pl.total > 0 ? pl.total : 256;
}

extern "C" int nbe_modulated_conv2d(const void* x, int dtype, const float* weight, const float* styles,
                                    const float* noise, int64_t noise_sn, void* y,
                                    int N, int Cin, int H, int W, int Cout, int K, int up, int padding,
                                    const float* resample_filter, int demodulate, int flip_weight,
                                    void* workspace, int64_t workspace_bytes, nbe_stream_t stream) {
    NBE_REQUIRE(x && weight && y, "modulated_conv2d: null tensor");
    NBE_REQUIRE(N >= 0 && Cin >= 1 && H >= 1 && W >= 1 && Cout >= 1 && K >= 1, "modulated_conv2d: bad shape");
    NBE_REQUIRE(up == 1 || up == 2, "modulated_conv2d: up must be 1 or 2 (down-sampling and other factors: compose conv2d_resample)");
    NBE_REQUIRE(padding >= 0, "modulated_conv2d: negative padding");
    NBE_REQUIRE(up == 1 || (K == 3 && resample_filter), "modulated_conv2d: up = 2 needs a 3x3 kernel and a 4x4 resample filter");
    NBE_REQUIRE(!demodulate || styles, "modulated_conv2d: demodulation needs styles");
    McPlan pl;
    if (!plan(pl, dtype, N, Cin, H, W, Cout, K, up, padding))
        return fail(NBE_EUNSUPPORTED, "modulated_conv2d: dtype %d / empty output not supported", dtype);
    if (!pl.tc && dtype != NBE_F32)
        return fail(NBE_EUNSUPPORTED, "modulated_conv2d: 16-bit activations need K = 3, Cout %% 128 == 0 and padding 1 (or 0) for the "
                                      "tensor-core kernels (got K %d, Cout %d, padding %d); convert to float32", K, Cout, padding);
    NBE_REQUIRE(workspace && workspace_bytes >= pl.total && (((uintptr_t)workspace) & 255) == 0,
                "modulated_conv2d: workspace of %lld bytes (256-byte aligned) required", (long long)pl.total);
    if (N == 0) return NBE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    float* wsq = (float*)(ws + pl.off_wsq);
    float* dcoef = demodulate ? (float*)(ws + pl.off_dcoef) : nullptr;
    int st;
    if (demodulate) {
        if ((st = nbe_weight_sqsum_f32(weight, wsq, Cout, Cin, K * K, stream))) return st;
        if ((st = nbe_demod_coefs_f32(styles, wsq, dcoef, N, Cin, Cout, stream))) return st;
    }
    const float ngain = noise ? 1.f : 0.f;
    if (!pl.tc) {
        // ---- float32: true-FP32 direct convolution; up = 2 FIR-first (per-channel scaling commutes with the per-channel FIR)
        const float* xin = (const float*)x;
        int cH = H, cW = W, cpad = padding;
        if (up == 2) {
            float* u = (float*)(ws + pl.off_t);
            const int UH = pl.OH + K - 1, UW = pl.OW + K - 1;     // = 2H + (p+2) + (p+1) - 3
            st = nbe_upfirdn2d(x, resample_filter, u, N, Cin, H, W, (int64_t)Cin * H * W, (int64_t)H * W, W, 1,
                               UH, UW, (int64_t)Cin * UH * UW, (int64_t)UH * UW, UW, 1, 4, 4, 2, 2, 1, 1,
                               padding + 2, padding + 1, padding + 2, padding + 1, 0, 4.f, NBE_F32, stream);
            if (st) return st;
            xin = u; cH = UH; cW = UW; cpad = 0;
        }
        return nbe_conv2d_f32(xin, weight, (float*)y, N, Cin, cH, cW, Cout, K, cpad, 1, 1, flip_weight ? 0 : 1,
                              styles, dcoef, noise, noise_sn, ngain, nullptr, 0, 0.f, 1.f, -1.f, stream);
    }
    // ---- 16-bit activations: tensor cores
    void* wq = ws + pl.off_wq;
    void* xin = ws + pl.off_xin;
    void* yf = ws + pl.off_y;
    // up = 1: flip_weight = True is a correlation (taps as stored); up = 2: conv_transpose2d scatters the taps as stored when
    // flip_weight = False (conv2d_resample.py:131-138 passes `not flip_weight` to a wrapper that flips when it gets False)
    const int flip = (up == 1) ? !flip_weight : (flip_weight ? 1 : 0);
    if ((st = nbe_prepare_weights_bf16(weight, wq, Cout, Cin, 3, flip, stream))) return st;
    st = (dtype == NBE_BF16) ? launch_pack<__nv_bfloat16>(x, xin, N, Cin, H, W, pl.P, pl.x_cs, styles, s)
                             : launch_pack<__half>(x, xin, N, Cin, H, W, pl.P, pl.x_cs, styles, s);
    if (st) return st;
    if (up == 1) {
        st = nbe_conv3x3_flat_bf16(xin, wq, yf, N, pl.OH, pl.OW, Cin, pl.x_cs, pl.P, padding == 0 ? 1 : 0, Cout, Cout,
                                   pl.OW, (int64_t)pl.OH * pl.OW, dcoef, noise, noise_sn, ngain, nullptr, 1.f, 1.f, -1.f, nullptr, stream);
        if (st) return st;
    } else {
        void* t = ws + pl.off_t;
        const int TH = 2 * H + 1, TW = 2 * W + 1, TP = 2 * W + 2;
        if ((st = nbe_convT3x3s2_flat_bf16(xin, wq, t, N, H, W, Cin, pl.x_cs, pl.P, Cout, Cout, TP, (int64_t)TH * TP, nullptr, stream))) return st;
        st = nbe_fir_act_nhwc_bf16(t, resample_filter, yf, N, pl.OH, pl.OW, Cout, TH, TW, padding, Cout, TP, (int64_t)TH * TP,
                                   Cout, pl.OW, (int64_t)pl.OH * pl.OW, 4.f, dcoef, noise, noise_sn, ngain, nullptr, 1.f, 1.f, -1.f,
                                   nullptr, stream);
        if (st) return st;
    }
    return (dtype == NBE_BF16) ? launch_unpack<__nv_bfloat16>(yf, y, N, Cout, pl.OH, pl.OW, pl.OW, Cout, s)
                               : launch_unpack<__half>(yf, y, N, Cout, pl.OH, pl.OW, pl.OW, Cout, s);
}
