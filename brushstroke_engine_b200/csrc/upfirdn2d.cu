// upfirdn2d forward.
//
//   y[n,c,oy,ox] = gain * sum_{a<fh,b<fw} ft[a,b] * xhat[oy*downy + a - pady0, ox*downx + b - padx0]
//   xhat[u,v] = x[n,c,u/upy,v/upx] when u,v >= 0, divisible by the up factors and in range, else 0
//   ft = f flipped in both axes unless `flip`                       
//
// Two kernels:
//  * upfirdn2d_tiled_kernel<T, UP>: the cases the generator / public helpers hit -- 4x4 filter, down = 1,
//    up in {1, 2}, dense NCHW.  One CTA = one 32x64 output tile of one (n,c) plane; the input footprint is
//    staged in shared memory as fp32 with coalesced row loads, every thread produces a 1x8 (UP=1) or
//    2x4 (UP=2) register strip so each staged value is re-used from registers, stores are row-contiguous.
//    HBM-bound: algorithmic bytes = numel(x) + numel(y) elements.
//  * upfirdn2d_generic_kernel<T>: any filter size / up / down / padding / strides (incl. channels_last),
//    one thread per output element.
#include "common.cuh"

namespace nbe {

struct UpfirdnParams {
    const void* x; const float* f; void* y;
    int N, C, H, W, OH, OW, fh, fw, upx, upy, downx, downy, padx0, pady0, flip;
    int64_t xs_n, xs_c, xs_h, xs_w, ys_n, ys_c, ys_h, ys_w;
    float gain;
};

template <class T>
__global__ void __launch_bounds__(256)
upfirdn2d_generic_kernel(UpfirdnParams p) {
    const int64_t total = (int64_t)p.N * p.C * p.OH * p.OW;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int ox = (int)(idx % p.OW);
        int64_t t = idx / p.OW;
        int oy = (int)(t % p.OH); t /= p.OH;
        int c = (int)(t % p.C);
        int n = (int)(t / p.C);
        const T* xp = (const T*)p.x + n * p.xs_n + c * p.xs_c;
        float acc = 0.f;
        for (int a = 0; a < p.fh; ++a) {
            int u = oy * p.downy + a - p.pady0;
            if (u < 0 || (u % p.upy) != 0) continue;
            int iy = u / p.upy;
            if (iy >= p.H) continue;
            for (int b = 0; b < p.fw; ++b) {
                int v = ox * p.downx + b - p.padx0;
                if (v < 0 || (v % p.upx) != 0) continue;
                int ix = v / p.upx;
                if (ix >= p.W) continue;
                float fv = p.flip ? p.f[a * p.fw + b] : p.f[(p.fh - 1 - a) * p.fw + (p.fw - 1 - b)];
                acc += fv * Cvt<T>::ld(xp[iy * p.xs_h + ix * p.xs_w]);
            }
        }
        ((T*)p.y)[n * p.ys_n + c * p.ys_c + oy * p.ys_h + ox * p.ys_w] = Cvt<T>::st(acc * p.gain);
    }
}

// ---- tiled 4x4 kernel -------------------------------------------------------------------------
constexpr int TILE_OH = 32;
constexpr int TILE_OW = 64;

template <class T, int UP>
__global__ void __launch_bounds__(256)
upfirdn2d_tiled_kernel(UpfirdnParams p, int tiles_x, int tiles_y) {
    // input footprint of a TILE_OH x TILE_OW output tile
    constexpr int IN_H = (UP == 1) ? TILE_OH + 3 : TILE_OH / 2 + 2;
    constexpr int IN_W = (UP == 1) ? TILE_OW + 3 : TILE_OW / 2 + 2;
    constexpr int IN_WP = IN_W + 1;                              // +1: odd pitch, conflict-free column walks
    __shared__ float s_in[IN_H][IN_WP];
    __shared__ float s_f[16];

    int tile = blockIdx.x;
    const int tx = tile % tiles_x; tile /= tiles_x;
    const int ty = tile % tiles_y; tile /= tiles_y;
    const int plane = tile;                                      // n * C + c
    const int oy0 = ty * TILE_OH, ox0 = tx * TILE_OW;
    const T* xp = (const T*)p.x + (int64_t)plane * p.H * p.W;
    T* yp = (T*)p.y + (int64_t)plane * p.OH * p.OW;

    if (threadIdx.x < 16) {
        int a = threadIdx.x >> 2, b = threadIdx.x & 3;
        s_f[threadIdx.x] = (p.flip ? p.f[a * 4 + b] : p.f[(3 - a) * 4 + (3 - b)]) * p.gain;
    }
    // First input row/col touched by this tile.  UP == 1: u = oy + a - pad.  UP == 2: iy = (oy + a - pad) / 2
    // for the taps with even (oy + a - pad); floor-div of the smallest candidate.
    int iy0, ix0;
    if (UP == 1) { iy0 = oy0 - p.pady0; ix0 = ox0 - p.padx0; }
    else {
        // floor((oy0 - pad) / 2) and one extra row of slack handled by IN_H = TILE/2 + 2
        int u0 = oy0 - p.pady0, v0 = ox0 - p.padx0;
        iy0 = (u0 >= 0) ? (u0 + 1) / 2 : -((-u0) / 2);           // ceil(u0 / 2): first even-aligned input row >= u0/2
        ix0 = (v0 >= 0) ? (v0 + 1) / 2 : -((-v0) / 2);
    }
    for (int i = threadIdx.x; i < IN_H * IN_W; i += 256) {
        int r = i / IN_W, c = i - r * IN_W;
        int iy = iy0 + r, ix = ix0 + c;
        float v = 0.f;
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v = Cvt<T>::ld(xp[(int64_t)iy * p.W + ix]);
        s_in[r][c] = v;
    }
    __syncthreads();

    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = s_f[i];

    if (UP == 1) {
        // thread -> column lx (0..63), rows ly0..ly0+7 ; 256 threads = 64 cols x 4 row groups
        const int lx = threadIdx.x & 63;
        const int ly0 = (threadIdx.x >> 6) * 8;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int r = 0; r < 11; ++r) {
            float v0 = s_in[ly0 + r][lx], v1 = s_in[ly0 + r][lx + 1], v2 = s_in[ly0 + r][lx + 2], v3 = s_in[ly0 + r][lx + 3];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                int o = r - a;                                   // output row (within strip) fed by input row r via tap a
                if (o >= 0 && o < 8)
                    acc[o] += f[a * 4 + 0] * v0 + f[a * 4 + 1] * v1 + f[a * 4 + 2] * v2 + f[a * 4 + 3] * v3;
            }
        }
        const int ox = ox0 + lx;
        if (ox < p.OW) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int oy = oy0 + ly0 + i;
                if (oy < p.OH) yp[(int64_t)oy * p.OW + ox] = Cvt<T>::st(acc[i]);
            }
        }
    } else {
        // Each thread: 2 output columns (ox even/odd pair) x 4 output rows.  256 threads = 32 col pairs x 8 row groups.
        const int lxp = threadIdx.x & 31;
        const int ly0 = (threadIdx.x >> 5) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int oy = oy0 + ly0 + i;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int ox = ox0 + lxp * 2 + j;
                float acc = 0.f;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    int u = oy + a - p.pady0;
                    if (u & 1) continue;
                    int r = (u >> 1) - iy0;                      // arithmetic shift = floor for negatives
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        int v = ox + b - p.padx0;
                        if (v & 1) continue;
                        int c = (v >> 1) - ix0;
                        acc += f[a * 4 + b] * s_in[r][c];
                    }
                }
                if (oy < p.OH && ox < p.OW) yp[(int64_t)oy * p.OW + ox] = Cvt<T>::st(acc);
            }
        }
    }
}

template <class T>
static int run_typed(const UpfirdnParams& p, bool tiled_ok, cudaStream_t s) {
    if (tiled_ok) {
        const int tiles_x = (p.OW + TILE_OW - 1) / TILE_OW, tiles_y = (p.OH + TILE_OH - 1) / TILE_OH;
        const int64_t blocks = (int64_t)tiles_x * tiles_y * p.N * p.C;
        if (blocks <= INT32_MAX) {
            if (p.upx == 1) upfirdn2d_tiled_kernel<T, 1><<<(int)blocks, 256, 0, s>>>(p, tiles_x, tiles_y);
            else            upfirdn2d_tiled_kernel<T, 2><<<(int)blocks, 256, 0, s>>>(p, tiles_x, tiles_y);
            return launched("upfirdn2d_tiled_kernel");
        }
    }
    const int64_t total = (int64_t)p.N * p.C * p.OH * p.OW;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
    upfirdn2d_generic_kernel<T><<<(int)blocks, 256, 0, s>>>(p);
    return launched("upfirdn2d_generic_kernel");
}

}  // namespace nbe

extern "C" int nbe_upfirdn2d(const void* x, const float* f, void* y,
                             int N, int C, int H, int W, int64_t xs_n, int64_t xs_c, int64_t xs_h, int64_t xs_w,
                             int OH, int OW, int64_t ys_n, int64_t ys_c, int64_t ys_h, int64_t ys_w,
                             int fh, int fw, int upx, int upy, int downx, int downy,
                             int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                             int dtype, nbe_stream_t stream) {
    using namespace nbe;
    NBE_REQUIRE(x && f && y, "upfirdn2d: null tensor");
    NBE_REQUIRE(fh >= 1 && fw >= 1, "upfirdn2d: f must be at least 1x1");
    NBE_REQUIRE(upx >= 1 && upy >= 1, "upfirdn2d: upsampling factor must be at least 1");
    NBE_REQUIRE(downx >= 1 && downy >= 1, "upfirdn2d: downsampling factor must be at least 1");
    NBE_REQUIRE(N >= 0 && C >= 0 && H >= 1 && W >= 1, "upfirdn2d: bad input shape");
    const int eow = (W * upx + padx0 + padx1 - fw + downx) / downx;       // upfirdn2d.cpp:32-33
    const int eoh = (H * upy + pady0 + pady1 - fh + downy) / downy;
    NBE_REQUIRE(eow >= 1 && eoh >= 1, "upfirdn2d: output must be at least 1x1");
    NBE_REQUIRE(OH == eoh && OW == eow, "upfirdn2d: output shape mismatch (%d,%d) vs expected (%d,%d)", OH, OW, eoh, eow);
    NBE_REQUIRE((int64_t)N * C * OH * OW <= INT32_MAX && (int64_t)N * C * H * W <= INT32_MAX, "upfirdn2d: tensor too large");
    if (N == 0 || C == 0) return NBE_OK;
    UpfirdnParams p;
    p.x = x; p.f = f; p.y = y; p.N = N; p.C = C; p.H = H; p.W = W; p.OH = OH; p.OW = OW; p.fh = fh; p.fw = fw;
    p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0; p.flip = flip;
    p.xs_n = xs_n; p.xs_c = xs_c; p.xs_h = xs_h; p.xs_w = xs_w; p.ys_n = ys_n; p.ys_c = ys_c; p.ys_h = ys_h; p.ys_w = ys_w;
    p.gain = gain;
    const bool dense_x = xs_w == 1 && xs_h == W && xs_c == (int64_t)H * W && xs_n == (int64_t)C * H * W;
    const bool dense_y = ys_w == 1 && ys_h == OW && ys_c == (int64_t)OH * OW && ys_n == (int64_t)C * OH * OW;
    const bool tiled_ok = dense_x && dense_y && fh == 4 && fw == 4 && downx == 1 && downy == 1 && upx == upy &&
                          (upx == 1 || upx == 2);
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case NBE_F32:  return run_typed<float>(p, tiled_ok, s);
        case NBE_F16:  return run_typed<__half>(p, tiled_ok, s);
        case NBE_BF16: return run_typed<__nv_bfloat16>(p, tiled_ok, s);
        default: return fail(NBE_EUNSUPPORTED, "upfirdn2d: unsupported dtype %d (float32/float16/bfloat16 only)", dtype);
    }
}
