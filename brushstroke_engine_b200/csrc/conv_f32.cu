// True-FP32 direct convolution (CUDA cores, FFMA) with the modulated-conv prologue/epilogue fused:
// the FP32 parity mode of modulated_conv2d (<= 1e-4 against the CPU reference needs real fp32
// operands *and* accumulation; tensor cores would round operands to TF32/BF16).
//
// conv_f32_tiled_kernel<K, S>: CTA = 64 output channels x (8 x 32) output pixels of one image,
// 256 threads, each thread an 8-channel x 8-pixel register tile (64 accumulators).  Input channels
// are streamed CT_CI (2..8) at a time through shared memory (input halo tile + the matching weight slab,
// out-channel innermost so a warp's weight reads are 128-bit broadcasts).  Per (ci, kh) a thread
// issues ~5 shared loads for 8*K*8 FFMAs.
// conv_f32_generic_kernel: any K / stride / groups, one thread per output (correctness path).
#include "common.cuh"

namespace nbe {

struct ConvParams {
    const float* x; const float* w; float* y;
    int N, Cin, H, W, Cout, OH, OW, K, pad, stride, groups, flip;
    const float* xscale; const float* dcoef; const float* noise; int64_t noise_sn; float noise_gain;
    const float* bias; int act; float alpha, gain, clamp;
};

__device__ __forceinline__ float conv_epilogue(const ConvParams& p, float acc, int n, int o, int oy, int ox) {
    if (p.dcoef) acc *= p.dcoef[(int64_t)n * p.Cout + o];
    if (p.noise) acc += p.noise[n * p.noise_sn + (int64_t)oy * p.OW + ox] * p.noise_gain;
    if (p.act) {
        if (p.bias) acc += p.bias[o];
        acc = apply_act(acc, p.act, p.alpha) * p.gain;
        if (p.clamp >= 0.f) acc = fminf(fmaxf(acc, -p.clamp), p.clamp);
    }
    return acc;
}

__global__ void __launch_bounds__(256)
conv_f32_generic_kernel(ConvParams p) {
    const int64_t total = (int64_t)p.N * p.Cout * p.OH * p.OW;
    const int cin_g = p.Cin / p.groups, cout_g = p.Cout / p.groups;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int ox = (int)(idx % p.OW);
        int64_t t = idx / p.OW;
        int oy = (int)(t % p.OH); t /= p.OH;
        int o = (int)(t % p.Cout);
        int n = (int)(t / p.Cout);
        const int g = o / cout_g;
        float acc = 0.f;
        for (int i = 0; i < cin_g; ++i) {
            const int ci = g * cin_g + i;
            const float sc = p.xscale ? p.xscale[(int64_t)n * p.Cin + ci] : 1.f;
            const float* xp = p.x + ((int64_t)n * p.Cin + ci) * p.H * p.W;
            const float* wp = p.w + ((int64_t)o * cin_g + i) * p.K * p.K;
            for (int kh = 0; kh < p.K; ++kh) {
                int iy = oy * p.stride + kh - p.pad;
                if (iy < 0 || iy >= p.H) continue;
                for (int kw = 0; kw < p.K; ++kw) {
                    int ix = ox * p.stride + kw - p.pad;
                    if (ix < 0 || ix >= p.W) continue;
                    float wv = p.flip ? wp[(p.K - 1 - kh) * p.K + (p.K - 1 - kw)] : wp[kh * p.K + kw];
                    acc = fmaf(wv, xp[(int64_t)iy * p.W + ix] * sc, acc);
                }
            }
        }
        p.y[idx] = conv_epilogue(p, acc, n, o, oy, ox);
    }
}

constexpr int CT_O = 64;      // out channels per CTA
constexpr int CT_H = 8;       // output rows per CTA
constexpr int CT_W = 32;      // output cols per CTA

template <int K, int S, int CT_CI>
__global__ void __launch_bounds__(256)
conv_f32_tiled_kernel(ConvParams p, int tiles_x, int tiles_y, int tiles_o) {
    constexpr int IN_H = (CT_H - 1) * S + K;
    constexpr int IN_W = (CT_W - 1) * S + K;
    constexpr int IN_WP = (IN_W + 3) & ~3;
    constexpr int KK = K * K;
    __shared__ __align__(16) float s_in[CT_CI][IN_H][IN_WP];
    __shared__ __align__(16) float s_w[CT_CI][KK][CT_O];

    int tile = blockIdx.x;
    const int tx = tile % tiles_x; tile /= tiles_x;
    const int ty = tile % tiles_y; tile /= tiles_y;
    const int to = tile % tiles_o; tile /= tiles_o;
    const int n = tile;
    const int oy0 = ty * CT_H, ox0 = tx * CT_W, o0 = to * CT_O;
    const int iy0 = oy0 * S - p.pad, ix0 = ox0 * S - p.pad;

    // thread -> (channel group cg of 8, pixel strip: row pr, 8 consecutive columns starting at pc*8)
    const int cg = threadIdx.x >> 5;                  // 0..7, uniform per warp -> weight loads broadcast
    const int lane = threadIdx.x & 31;
    const int pr = lane >> 2;                         // 0..7
    const int pc = lane & 3;                          // 0..3

    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

    for (int c0 = 0; c0 < p.Cin; c0 += CT_CI) {
        __syncthreads();
        for (int i = threadIdx.x; i < CT_CI * IN_H * IN_W; i += 256) {
            int c = i / (IN_H * IN_W);
            int r = (i / IN_W) % IN_H;
            int col = i % IN_W;
            int ci = c0 + c, iy = iy0 + r, ix = ix0 + col;
            float v = 0.f;
            if (ci < p.Cin && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
                v = p.x[(((int64_t)n * p.Cin + ci) * p.H + iy) * p.W + ix];
                if (p.xscale) v *= p.xscale[(int64_t)n * p.Cin + ci];
            }
            s_in[c][r][col] = v;
        }
        for (int i = threadIdx.x; i < CT_O * CT_CI * KK; i += 256) {
            int o = i / (CT_CI * KK);
            int rem = i - o * (CT_CI * KK);
            int c = rem / KK, t = rem - c * KK;
            float v = 0.f;
            if (o0 + o < p.Cout && c0 + c < p.Cin) {
                int tt = p.flip ? (KK - 1 - t) : t;
                v = p.w[((int64_t)(o0 + o) * p.Cin + (c0 + c)) * KK + tt];
            }
            s_w[c][t][o] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < CT_CI; ++c) {
#pragma unroll
            for (int kh = 0; kh < K; ++kh) {
                float xin[(8 - 1) * S + K];
#pragma unroll
                for (int j = 0; j < (8 - 1) * S + K; ++j) xin[j] = s_in[c][pr * S + kh][pc * 8 * S + j];
#pragma unroll
                for (int kw = 0; kw < K; ++kw) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&s_w[c][kh * K + kw][cg * 8]);
                    const float4 w1 = *reinterpret_cast<const float4*>(&s_w[c][kh * K + kw][cg * 8 + 4]);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int a = 0; a < 8; ++a)
#pragma unroll
                        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(wv[a], xin[b * S + kw], acc[a][b]);
                }
            }
        }
    }
    const int oy = oy0 + pr;
    if (oy < p.OH) {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int o = o0 + cg * 8 + a;
            if (o >= p.Cout) continue;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int ox = ox0 + pc * 8 + b;
                if (ox < p.OW)
                    p.y[(((int64_t)n * p.Cout + o) * p.OH + oy) * p.OW + ox] = conv_epilogue(p, acc[a][b], n, o, oy, ox);
            }
        }
    }
}

template <int K, int S, int CT_CI>
static int launch_tiled(const ConvParams& p, cudaStream_t s) {
    const int tiles_x = (p.OW + CT_W - 1) / CT_W, tiles_y = (p.OH + CT_H - 1) / CT_H, tiles_o = (p.Cout + CT_O - 1) / CT_O;
    const int64_t blocks = (int64_t)tiles_x * tiles_y * tiles_o * p.N;
    if (blocks > INT32_MAX) return fail(NBE_EINVAL, "conv2d_f32: grid too large");
    conv_f32_tiled_kernel<K, S, CT_CI><<<(int)blocks, 256, 0, s>>>(p, tiles_x, tiles_y, tiles_o);
    return launched("conv_f32_tiled_kernel");
}

__global__ void weight_sqsum_kernel(const float* __restrict__ w, float* __restrict__ wsq, int n, int KK) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < KK; ++k) { float v = w[(int64_t)i * KK + k]; s = fmaf(v, v, s); }
    wsq[i] = s;
}

// one warp per (n, o): d = rsqrt(sum_i s^2 * wsq + 1e-8)
__global__ void demod_coefs_kernel(const float* __restrict__ styles, const float* __restrict__ wsq, float* __restrict__ d,
                                   int N, int Cin, int Cout) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= N * Cout) return;
    int n = warp / Cout, o = warp - n * Cout;
    float s = 0.f;
    for (int i = lane; i < Cin; i += 32) {
        float st = styles[(int64_t)n * Cin + i];
        s = fmaf(st * st, wsq[(int64_t)o * Cin + i], s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) d[warp] = rsqrtf(s + 1e-8f);
}

}  // namespace nbe

extern "C" int nbe_conv2d_f32(const float* x, const float* w, float* y,
                              int N, int Cin, int H, int W, int Cout, int K, int pad, int stride, int groups, int flip,
                              const float* xscale, const float* dcoef,
                              const float* noise, int64_t noise_sn, float noise_gain,
                              const float* bias, int act, float alpha, float gain, float clamp,
                              nbe_stream_t stream) {
    using namespace nbe;
    NBE_REQUIRE(x && w && y, "conv2d_f32: null tensor");
    NBE_REQUIRE(N >= 0 && Cin >= 1 && Cout >= 1 && H >= 1 && W >= 1 && K >= 1 && stride >= 1 && groups >= 1 && pad >= 0,
                "conv2d_f32: bad shape");
    NBE_REQUIRE(Cin % groups == 0 && Cout % groups == 0, "conv2d_f32: channels not divisible by groups");
    NBE_REQUIRE(act >= 0 && act <= NBE_ACT_SWISH, "conv2d_f32: bad activation");
    ConvParams p;
    p.x = x; p.w = w; p.y = y; p.N = N; p.Cin = Cin; p.H = H; p.W = W; p.Cout = Cout; p.K = K; p.pad = pad;
    p.stride = stride; p.groups = groups; p.flip = flip;
    p.OH = (H + 2 * pad - K) / stride + 1;
    p.OW = (W + 2 * pad - K) / stride + 1;
    NBE_REQUIRE(p.OH >= 1 && p.OW >= 1, "conv2d_f32: empty output");
    p.xscale = xscale; p.dcoef = dcoef; p.noise = noise; p.noise_sn = noise_sn; p.noise_gain = noise_gain;
    p.bias = bias; p.act = act; p.alpha = alpha; p.gain = gain; p.clamp = clamp;
    if (N == 0) return NBE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (groups == 1) {
        if (K == 3 && stride == 1) return launch_tiled<3, 1, 8>(p, s);
        if (K == 1 && stride == 1) return launch_tiled<1, 1, 8>(p, s);
        if (K == 3 && stride == 2) return launch_tiled<3, 2, 4>(p, s);
        if (K == 7 && stride == 1) return launch_tiled<7, 1, 2>(p, s);
    }
    const int64_t total = (int64_t)N * Cout * p.OH * p.OW;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
    conv_f32_generic_kernel<<<(int)blocks, 256, 0, s>>>(p);
    return launched("conv_f32_generic_kernel");
}

extern "C" int nbe_weight_sqsum_f32(const float* w, float* wsq, int Cout, int Cin, int KK, nbe_stream_t stream) {
    using namespace nbe;
    NBE_REQUIRE(w && wsq && Cout >= 1 && Cin >= 1 && KK >= 1, "weight_sqsum: bad arguments");
    const int n = Cout * Cin;
    weight_sqsum_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, wsq, n, KK);
    return launched("weight_sqsum_kernel");
}

extern "C" int nbe_demod_coefs_f32(const float* styles, const float* wsq, float* d, int N, int Cin, int Cout,
                                   nbe_stream_t stream) {
    using namespace nbe;
    NBE_REQUIRE(styles && wsq && d && N >= 0 && Cin >= 1 && Cout >= 1, "demod_coefs: bad arguments");
    if (N == 0) return NBE_OK;
    const int64_t threads = (int64_t)N * Cout * 32;
    demod_coefs_kernel<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(styles, wsq, d, N, Cin, Cout);
    return launched("demod_coefs_kernel");
}
