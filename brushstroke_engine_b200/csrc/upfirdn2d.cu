// upfirdn2d forward.
//
//   y[n,c,oy,ox] = gain * sum_{a<fh,b<fw} ft[a,b] * xhat[oy*downy + a - pady0, ox*downx + b - padx0]
//   xhat[u,v] = x[n,c,u/upy,v/upx] when u,v >= 0, divisible by the up factors and in range, else 0
//   ft = f flipped in both axes unless `flip`                       
//
// Two kernels:
//  * upfirdn2d_rows_up{1,2}_kernel<T>: the cases the generator / public helpers hit -- 4x4 filter, down = 1,
//    up in {1, 2}, dense NCHW.  A lane owns 16 bytes of consecutive output columns and walks 16 output rows with a
//    register window of input rows (no shared memory, no barriers); stores are 16-byte streaming vectors.
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

// ---- register-window 4x4 kernels (dense NCHW, down = 1) ------------------------------------------
// One lane owns VPT consecutive output columns (16 bytes of output) and walks ROWS output rows of one (n,c) plane
// keeping the input rows it needs in registers: no shared memory, no barriers, ~1.5 global loads (L1 hits for the
// overlap between neighbouring lanes / rows) + 16 (UP=1) or 4 (UP=2) FMAs + 1/VPT vector store per output.
constexpr int RW_ROWS = 16;

template <class T> struct VecOut;                                       // VPT outputs packed into one 16-byte store
template <> struct VecOut<float> { static constexpr int VPT = 4; };
template <> struct VecOut<__half> { static constexpr int VPT = 8; };
template <> struct VecOut<__nv_bfloat16> { static constexpr int VPT = 8; };

template <class T, int VPT>
__device__ __forceinline__ void store_row(T* dst, const float (&acc)[VPT], int n_valid) {
    if (n_valid == VPT && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        int4 v;
        T* e = reinterpret_cast<T*>(&v);
#pragma unroll
        for (int k = 0; k < VPT; ++k) e[k] = Cvt<T>::st(acc[k]);
        st_stream16(dst, v);
    } else {
#pragma unroll
        for (int k = 0; k < VPT; ++k) if (k < n_valid) dst[k] = Cvt<T>::st(acc[k]);
    }
}

template <class T>
__global__ void __launch_bounds__(128)
upfirdn2d_rows_up1_kernel(UpfirdnParams p, int col_blocks, int row_blocks) {
    constexpr int VPT = VecOut<T>::VPT;
    constexpr int WIN = VPT + 3;
    __shared__ float s_f[16];
    if (threadIdx.x < 16) {
        int a = threadIdx.x >> 2, b = threadIdx.x & 3;
        s_f[threadIdx.x] = (p.flip ? p.f[a * 4 + b] : p.f[(3 - a) * 4 + (3 - b)]) * p.gain;
    }
    __syncthreads();
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = s_f[i];
    int blk = blockIdx.x;
    const int cb = blk % col_blocks; blk /= col_blocks;
    const int rb = blk % row_blocks; blk /= row_blocks;
    const int plane = blk;
    const int ox0 = (cb * 128 + threadIdx.x) * VPT;
    if (ox0 >= p.OW) return;
    const int oy0 = rb * RW_ROWS;
    const T* xp = (const T*)p.x + (int64_t)plane * p.H * p.W;
    T* yp = (T*)p.y + (int64_t)plane * p.OH * p.OW;
    const int ix0 = ox0 - p.padx0;
    const int n_valid = min(VPT, p.OW - ox0);
    float win[4][WIN];
    auto load_row = [&](float (&dst)[WIN], int iy) {
        const bool row_ok = iy >= 0 && iy < p.H;
        const T* rp = xp + (int64_t)iy * p.W;
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
            const int ix = ix0 + j;
            dst[j] = (row_ok && ix >= 0 && ix < p.W) ? Cvt<T>::ld(rp[ix]) : 0.f;
        }
    };
    const int iy0 = oy0 - p.pady0;
    load_row(win[0], iy0); load_row(win[1], iy0 + 1); load_row(win[2], iy0 + 2);
#pragma unroll
    for (int r = 0; r < RW_ROWS; ++r) {
        const int oy = oy0 + r;
        if (oy >= p.OH) break;
        load_row(win[(r + 3) & 3], iy0 + r + 3);
        float acc[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) acc[k] = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int k = 0; k < VPT; ++k) acc[k] = fmaf(f[a * 4 + b], win[(r + a) & 3][k + b], acc[k]);
        store_row<T, VPT>(yp + (int64_t)oy * p.OW + ox0, acc, n_valid);
    }
}

template <class T>
__global__ void __launch_bounds__(128)
upfirdn2d_rows_up2_kernel(UpfirdnParams p, int col_blocks, int row_blocks) {
    constexpr int VPT = VecOut<T>::VPT;
    constexpr int WIN = VPT / 2 + 2;
    __shared__ float s_f[16];
    if (threadIdx.x < 16) {
        int a = threadIdx.x >> 2, b = threadIdx.x & 3;
        s_f[threadIdx.x] = (p.flip ? p.f[a * 4 + b] : p.f[(3 - a) * 4 + (3 - b)]) * p.gain;
    }
    __syncthreads();
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = s_f[i];
    int blk = blockIdx.x;
    const int cb = blk % col_blocks; blk /= col_blocks;
    const int rb = blk % row_blocks; blk /= row_blocks;
    const int plane = blk;
    const int ox0 = (cb * 128 + threadIdx.x) * VPT;
    if (ox0 >= p.OW) return;
    const int oy0 = rb * RW_ROWS;
    const T* xp = (const T*)p.x + (int64_t)plane * p.H * p.W;
    T* yp = (T*)p.y + (int64_t)plane * p.OH * p.OW;
    const int n_valid = min(VPT, p.OW - ox0);
    // first input column any of this lane's outputs can touch: floor((ox0 - padx0) / 2) (arithmetic shift = floor)
    const int ixb = (ox0 - p.padx0) >> 1;
    // per output column k and tap parity: b0 = first tap with even (ox + b - padx0)
#pragma unroll 1
    for (int r = 0; r < RW_ROWS; ++r) {
        const int oy = oy0 + r;
        if (oy >= p.OH) break;
        const int u0 = oy - p.pady0;
        const int a0 = u0 & 1;                                      // taps a0, a0 + 2 hit even rows of the zero-stuffed grid
        const int iyA = (u0 + a0) >> 1;
        float rowA[WIN], rowB[WIN];
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
            const int ix = ixb + j;
            const bool cok = ix >= 0 && ix < p.W;
            rowA[j] = (cok && iyA >= 0 && iyA < p.H) ? Cvt<T>::ld(xp[(int64_t)iyA * p.W + ix]) : 0.f;
            rowB[j] = (cok && iyA + 1 >= 0 && iyA + 1 < p.H) ? Cvt<T>::ld(xp[(int64_t)(iyA + 1) * p.W + ix]) : 0.f;
        }
        float acc[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int v0 = ox0 + k - p.padx0;
            const int b0 = v0 & 1;
            const int j0 = ((v0 + b0) >> 1) - ixb;                  // 0 .. WIN-2
            // select the four active taps without dynamic indexing (keeps f[] in registers)
            const float fa0 = a0 ? (b0 ? f[5] : f[4]) : (b0 ? f[1] : f[0]);
            const float fa1 = a0 ? (b0 ? f[7] : f[6]) : (b0 ? f[3] : f[2]);
            const float fb0 = a0 ? (b0 ? f[13] : f[12]) : (b0 ? f[9] : f[8]);
            const float fb1 = a0 ? (b0 ? f[15] : f[14]) : (b0 ? f[11] : f[10]);
            float x00 = 0.f, x01 = 0.f, x10 = 0.f, x11 = 0.f;
#pragma unroll
            for (int j = 0; j < WIN - 1; ++j)                       // static indexing of the register window
                if (j == j0) { x00 = rowA[j]; x01 = rowA[j + 1]; x10 = rowB[j]; x11 = rowB[j + 1]; }
            acc[k] = fa0 * x00 + fa1 * x01 + fb0 * x10 + fb1 * x11;
        }
        store_row<T, VPT>(yp + (int64_t)oy * p.OW + ox0, acc, n_valid);
    }
}

template <class T>
static int run_typed(const UpfirdnParams& p, bool tiled_ok, cudaStream_t s) {
    if (tiled_ok) {
        constexpr int VPT = VecOut<T>::VPT;
        const int col_blocks = (p.OW + 128 * VPT - 1) / (128 * VPT), row_blocks = (p.OH + RW_ROWS - 1) / RW_ROWS;
        const int64_t blocks = (int64_t)col_blocks * row_blocks * p.N * p.C;
        if (blocks <= INT32_MAX) {
            if (p.upx == 1) upfirdn2d_rows_up1_kernel<T><<<(int)blocks, 128, 0, s>>>(p, col_blocks, row_blocks);
            else            upfirdn2d_rows_up2_kernel<T><<<(int)blocks, 128, 0, s>>>(p, col_blocks, row_blocks);
            return launched("upfirdn2d_rows_kernel");
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
