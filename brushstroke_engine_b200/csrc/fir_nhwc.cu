// 4x4 FIR + modulated-conv epilogue on NHWC bf16: the second half of an up-sampling SynthesisLayer when the
// transposed convolution runs at its algorithmic cost (nbe_convT3x3s2_flat_bf16):
//   y[n,oy,ox,c] = post( fgain * sum_{a,b} ft[a,b] * T[n, oy+a-pad, ox+b-pad, c] )     (upfirdn2d after conv_transpose2d,
//   SG2/torch_utils/ops/conv2d_resample.py:139 ; T is zero outside [0,TH) x [0,TW))
//   post(v) = clamp(lrelu(v * scale[n,c] + noise[n,oy,ox] * noise_gain + bias[c]) * gain) * next_scale[n,c]
//   (SG2/training/networks.py:71-75,386-390).
// One thread = one output column x 4 channels (8 bytes; keeps the register footprint small enough for 4 CTAs/SM, which is
// what hides the load latency), walking 8 output rows with a sliding window of horizontally
// filtered rows (rank-1 filters: 4 + 4 FMAs per output; general filters: 16).  Consecutive threads walk the channel
// vectors of a pixel, then the pixels of a row, so every global access is a whole 32-byte sector and the 4x overlap
// between neighbouring windows is served by L1.  HBM-bound: (TH*TW + OH*OW) * C * 2 bytes per image.
#include "tc_common.cuh"
#include <mutex>
#include <type_traits>
#include <algorithm>
#include <cstdlib>

namespace nbe {

constexpr int FIR_RPT = 8;

struct FirParams {
    const __nv_bfloat16* t; __nv_bfloat16* y; const float* f;
    int N, OH, OW, C, TH, TW, pad;
    int t_cs; long long t_row_pitch, t_img_pitch;
    int y_cs; long long y_row_pitch, y_img_pitch;
    float fgain;
    const float* scale; const float* noise; long long noise_sn; float noise_gain; const float* bias;
    float alpha, gain, clamp; const float* next_scale;
    int row_groups;
    int debug;                                                      // NBE_FIR_DEBUG (timing experiments, WRONG results): 1 = no global stores, 2 = no arithmetic (first tile pixel stored)
};

constexpr int FIR_CPT = 4;                                          // channels per thread

__device__ __forceinline__ void unpack4(const uint2& raw, float (&v)[FIR_CPT]) {
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int k = 0; k < 2; ++k) { const float2 f2 = __bfloat1622float2(h2[k]); v[2 * k] = f2.x; v[2 * k + 1] = f2.y; }
}

__global__ void __launch_bounds__(256, 4)
fir_act_nhwc_kernel(const FirParams p) {
    __shared__ float s_f[16];
    extern __shared__ float s_epi[];                               // [3][C]: scale, bias, next_scale of this CTA's image
    if (threadIdx.x < 16) {
        const int a = threadIdx.x >> 2, b = threadIdx.x & 3;
        s_f[threadIdx.x] = p.f[(3 - a) * 4 + (3 - b)] * p.fgain;      // flip_filter = False
    }
    {
        const int n_ = blockIdx.x / p.row_groups;
        for (int c = threadIdx.x; c < p.C; c += 256) {
            s_epi[c] = p.scale ? p.scale[(long long)n_ * p.C + c] : 1.f;
            s_epi[p.C + c] = p.bias ? p.bias[c] : 0.f;
            s_epi[2 * p.C + c] = p.next_scale ? p.next_scale[(long long)n_ * p.C + c] : 1.f;
        }
    }
    __syncthreads();
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = s_f[i];
    bool sep = f[0] != 0.f;
#pragma unroll
    for (int a = 1; a < 4; ++a)
#pragma unroll
        for (int b = 1; b < 4; ++b) sep = sep && fabsf(f[a * 4 + b] * f[0] - f[a * 4] * f[b]) <= 1e-6f * fabsf(f[a * 4 + b] * f[0]) + 1e-30f;
    float fx[4], fy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { fx[i] = f[i]; fy[i] = sep ? f[i * 4] / f[0] : 0.f; }

    const int CV = p.C / FIR_CPT;
    const int n = blockIdx.x / p.row_groups;
    const int oy0 = (blockIdx.x - n * p.row_groups) * FIR_RPT;
    const int idx = blockIdx.y * blockDim.x + threadIdx.x;
    if (idx >= p.OW * CV) return;
    const int cv = idx % CV, ox = idx / CV;
    const __nv_bfloat16* tp = p.t + (long long)n * p.t_img_pitch * p.t_cs + cv * FIR_CPT;
    const int ix0 = ox - p.pad;

    const float* sc = s_epi + cv * FIR_CPT;                        // per-channel epilogue vectors (shared memory)
    const float* bs = s_epi + p.C + cv * FIR_CPT;
    const float* ns = s_epi + 2 * p.C + cv * FIR_CPT;
    const float pos_gain = p.gain, neg_gain = p.gain * p.alpha;

    auto load_px = [&](int iy, int ix, float (&v)[FIR_CPT]) {
        if (iy >= 0 && iy < p.TH && ix >= 0 && ix < p.TW) {
            const uint2 raw = *reinterpret_cast<const uint2*>(tp + ((long long)iy * p.t_row_pitch + ix) * p.t_cs);
            unpack4(raw, v);
        } else {
#pragma unroll
            for (int k = 0; k < FIR_CPT; ++k) v[k] = 0.f;
        }
    };
    auto finish = [&](int oy, float (&acc)[FIR_CPT]) {
        float nz = 0.f;
        if (p.noise) nz = p.noise[(long long)n * p.noise_sn + (long long)oy * p.OW + ox] * p.noise_gain;
        uint2 outv;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&outv);
        float r[FIR_CPT];
#pragma unroll
        for (int k = 0; k < FIR_CPT; ++k) {
            float a = acc[k] * sc[k] + nz + bs[k];
            a *= (a > 0.f) ? pos_gain : neg_gain;
            if (p.clamp >= 0.f) a = fminf(fmaxf(a, -p.clamp), p.clamp);
            r[k] = a * ns[k];
        }
#pragma unroll
        for (int k = 0; k < FIR_CPT / 2; ++k) o2[k] = __floats2bfloat162_rn(r[2 * k], r[2 * k + 1]);
        *reinterpret_cast<uint2*>(p.y + (((long long)n * p.y_img_pitch + (long long)oy * p.y_row_pitch + ox) * p.y_cs + cv * FIR_CPT)) = outv;
    };

    const int iy0 = oy0 - p.pad;
    if (sep) {
        float h[4][FIR_CPT];
        auto hrow = [&](float (&dst)[FIR_CPT], int iy) {
            float v0[FIR_CPT], v1[FIR_CPT], v2[FIR_CPT], v3[FIR_CPT];
            load_px(iy, ix0, v0); load_px(iy, ix0 + 1, v1); load_px(iy, ix0 + 2, v2); load_px(iy, ix0 + 3, v3);
#pragma unroll
            for (int k = 0; k < FIR_CPT; ++k) dst[k] = fx[0] * v0[k] + fx[1] * v1[k] + fx[2] * v2[k] + fx[3] * v3[k];
        };
        hrow(h[0], iy0); hrow(h[1], iy0 + 1); hrow(h[2], iy0 + 2);
#pragma unroll
        for (int r = 0; r < FIR_RPT; ++r) {
            const int oy = oy0 + r;
            if (oy >= p.OH) break;
            hrow(h[(r + 3) & 3], iy0 + r + 3);
            float acc[FIR_CPT];
#pragma unroll
            for (int k = 0; k < FIR_CPT; ++k)
                acc[k] = fy[0] * h[r & 3][k] + fy[1] * h[(r + 1) & 3][k] + fy[2] * h[(r + 2) & 3][k] + fy[3] * h[(r + 3) & 3][k];
            finish(oy, acc);
        }
    } else {
#pragma unroll 1
        for (int r = 0; r < FIR_RPT; ++r) {
            const int oy = oy0 + r;
            if (oy >= p.OH) break;
            float acc[FIR_CPT];
#pragma unroll
            for (int k = 0; k < FIR_CPT; ++k) acc[k] = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float v[FIR_CPT];
                    load_px(oy - p.pad + a, ix0 + b, v);
#pragma unroll
                    for (int k = 0; k < FIR_CPT; ++k) acc[k] = fmaf(f[a * 4 + b], v[k], acc[k]);
                }
            finish(oy, acc);
        }
    }
}

// ---- TMA-tiled variant for C == 128 ------------------------------------------------------------------
// The register-window kernel above is latency-bound (every thread waits on its own global loads).  Here a persistent CTA
// streams (8+3) x (16+3) pixel x 128-channel tiles of T through a double-buffered shared-memory ring with 4-D TMA boxes
// (out-of-range rows / columns are zero-filled by TMA = the FIR's padding), the next tile is in flight while the
// current one is filtered, and every shared-memory read is a conflict-free 8-byte vector (a warp reads the 256 contiguous
// bytes of one pixel).  The tile's 8 x 16 noise values ride on the same mbarrier through a second (fp32) tensor map, and every
// CTA walks one contiguous run of tiles so that the per-image epilogue vectors are reloaded once per image, not per tile.
constexpr int FT_OW = 16, FT_OH = 8;
constexpr int FT_IW = FT_OW + 3, FT_IH = FT_OH + 3;
constexpr int FT_TILE_BYTES = FT_IH * FT_IW * 256;                  // 128 bf16 channels per pixel

template <bool PACKED>
__device__ __forceinline__ void fir_tiled_body(const CUtensorMap& tmap_t, const CUtensorMap& tmap_t8, const CUtensorMap& tmap_nz,
                                               const FirParams& p, int tiles_x, int tiles_y, int total_tiles, int noise_mode, int strip_ok) {
    extern __shared__ uint8_t smem_raw[];
    pdl_trigger();
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint8_t* bufs = smem;                                           // [2][FT_TILE_BYTES]
    constexpr int BUF_STRIDE = (FT_TILE_BYTES + 127) & ~127;
    float* s_nz = reinterpret_cast<float*>(bufs + 2 * BUF_STRIDE);  // [2][FT_OH][FT_OW] noise tiles (TMA destinations, 512 B each)
    float* s_epi = s_nz + 2 * FT_OH * FT_OW;                        // [3][128]
    float* s_f = s_epi + 3 * 128;                                   // [16]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_f + 16);         // [2]

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&full[0]), 1); mbar_init(smem_u32(&full[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_t) : "memory");
        if (noise_mode == 1) asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_nz) : "memory");
    }
    if (noise_mode == 2) s_nz[threadIdx.x] = 0.f;                   // no noise input: both tiles stay zero (2 x 128 floats)
    pdl_wait();                                                     // T, the filter, the epilogue vectors and y belong to other kernels of the chain
    if (threadIdx.x < 16) {
        const int a = threadIdx.x >> 2, b = threadIdx.x & 3;
        s_f[threadIdx.x] = p.f[(3 - a) * 4 + (3 - b)] * p.fgain;
    }
    __syncthreads();
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = s_f[i];
    bool sep = f[0] != 0.f;
#pragma unroll
    for (int a = 1; a < 4; ++a)
#pragma unroll
        for (int b = 1; b < 4; ++b) sep = sep && fabsf(f[a * 4 + b] * f[0] - f[a * 4] * f[b]) <= 1e-6f * fabsf(f[a * 4 + b] * f[0]) + 1e-30f;
    float fx[4], fy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { fx[i] = f[i]; fy[i] = sep ? f[i * 4] / f[0] : 0.f; }

    // thread = 4 channels (8 bytes) x 2 adjacent output pixels x FT_OH rows: the pair shares 5 input columns per row (2.5
    // unpacked vectors per output instead of 4), a warp reads 256 contiguous bytes of one pixel (conflict-free)
    const int cq = threadIdx.x & 31, px = (threadIdx.x >> 5) * 2;
    // lrelu(a) * gain == max(a*gain, a*gain*alpha) for gain > 0, 0 <= alpha <= 1: gain is folded into scale/bias/noise
    const bool fold = p.gain > 0.f && p.alpha >= 0.f && p.alpha <= 1.f;
    const float g_pre = fold ? p.gain : 1.f, g_post = fold ? 1.f : p.gain;
    const float clamp_hi = p.clamp >= 0.f ? p.clamp : INFINITY;
    const float ngain = p.noise_gain * g_pre;
    // noise_mode 1: the tile's 8 x 16 noise values arrive by TMA with the tile itself (same mbarrier; out-of-range elements are
    // zero-filled); 0: staged by the threads (noise tensors TMA cannot describe); 2: no noise
    // every CTA takes one contiguous run of tiles: the image (and with it the per-channel epilogue vectors in s_epi) changes
    // once per 128 tiles instead of at every tile, so the loop has no exposed global-memory round trip and one barrier per tile
    const int per_cta = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    // strip_ok == 2, CYCLIC strip order: CTA k takes the strips k, k + grid, k + 2 grid, ... -- neighbouring CTAs then filter
    // neighbouring strips of one image at the same time, and the 3 halo columns two strips share are read from DRAM once and from
    // L2 the second time (in contiguous runs the two strips are ~40 MB of traffic apart).  `tile` is then a CTA-local counter.
    const bool cyclic = PACKED && strip_ok == 2 && p.debug != 2;
    const int total_strips = tiles_x * (total_tiles / (tiles_x * tiles_y));
    const int my_strips = cyclic ? (total_strips - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int t_begin = cyclic ? 0 : blockIdx.x * per_cta, t_end = cyclic ? my_strips * tiles_y : min(total_tiles, t_begin + per_cta);
    // STRIP order (packed separable path): a run walks DOWN a 16-pixel-wide strip of its image, so the last three horizontally
    // filtered rows of a tile are the first three of the next one and stay in REGISTERS: only the first tile of a strip (or of
    // the run) loads its 3 halo rows, every other tile loads 8 rows instead of 11 -- the tile loads are what this pass waits for
    // (0.44 of its 0.49 ms at 128^2 remain with the arithmetic compiled out), and they shrink by 27 %
    if (PACKED && !sep) __trap();                                   // unreachable: the kernel sends other filters to fir_tiled_generic
    const bool strip = PACKED && sep && strip_ok && p.debug != 2;
    auto decode = [&](int tile, int& tx, int& ty, int& n) {
        int t = tile;
        if (cyclic) { ty = t % tiles_y; t = (int)blockIdx.x + (t / tiles_y) * (int)gridDim.x; tx = t % tiles_x; t /= tiles_x; }
        else if (strip) { ty = t % tiles_y; t /= tiles_y; tx = t % tiles_x; t /= tiles_x; }
        else       { tx = t % tiles_x; t /= tiles_x; ty = t % tiles_y; t /= tiles_y; }
        n = t;
    };
    auto issue = [&](int tile, int b) {
        int tx, ty, t;
        decode(tile, tx, ty, t);
        const bool fresh = !strip || tile == t_begin || ty == 0;
        const uint32_t bar = smem_u32(&full[b]);
        mbar_expect_tx(bar, (fresh ? FT_TILE_BYTES : FT_OH * FT_IW * 256) + (noise_mode == 1 ? FT_OH * FT_OW * 4 : 0));
        if (fresh) tma_load_4d(smem_u32(bufs + b * BUF_STRIDE), &tmap_t, bar, 0, tx * FT_OW - p.pad, ty * FT_OH - p.pad, t);
        else       tma_load_4d(smem_u32(bufs + b * BUF_STRIDE + 3 * FT_IW * 256), &tmap_t8, bar, 0, tx * FT_OW - p.pad, ty * FT_OH - p.pad + 3, t);
        if (noise_mode == 1) tma_load_3d(smem_u32(s_nz + b * FT_OH * FT_OW), &tmap_nz, bar, tx * FT_OW, ty * FT_OH, p.noise_sn ? t : 0);
    };
    if (threadIdx.x == 0 && t_begin < t_end) issue(t_begin, 0);
    int it = 0, cur_n = -1;
    float2 h[4][2][2];                                              // packed path: horizontally filtered rows, carried from tile to tile in strip order
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const int b = it & 1;
        if (threadIdx.x == 0 && tile + 1 < t_end) issue(tile + 1, b ^ 1);   // buffer b^1 was released by the __syncthreads of the previous iteration
        int tx, ty, n;
        decode(tile, tx, ty, n);
        const bool fresh = !strip || tile == t_begin || ty == 0;
        if (n != cur_n) {
            for (int c = threadIdx.x; c < 128; c += 256) {
                s_epi[c] = (p.scale ? p.scale[(long long)n * 128 + c] : 1.f) * g_pre;
                s_epi[128 + c] = (p.bias ? p.bias[c] : 0.f) * g_pre;
                s_epi[256 + c] = p.next_scale ? p.next_scale[(long long)n * 128 + c] : 1.f;
            }
            cur_n = n;
            __syncthreads();
        }
        const int ox = tx * FT_OW + px, oy0 = ty * FT_OH;
        // tiles that lie completely inside the output (all of them at 2^k resolutions) skip the per-store bounds checks
        const bool full_tile = oy0 + FT_OH <= p.OH && tx * FT_OW + FT_OW <= p.OW;
        const float* s_nzb = s_nz + b * FT_OH * FT_OW;
        if (noise_mode == 0) {
            if (threadIdx.x < FT_OH * FT_OW) {
                const int r = threadIdx.x / FT_OW, c = threadIdx.x - r * FT_OW;
                const int oy = oy0 + r, oxx = tx * FT_OW + c;
                float v = 0.f;
                if (oy < p.OH && oxx < p.OW) v = __ldg(p.noise + (long long)n * p.noise_sn + (long long)oy * p.OW + oxx);
                s_nz[b * FT_OH * FT_OW + threadIdx.x] = v;
            }
            __syncthreads();
        }
        float sc[4], bs[4], ns[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { sc[k] = s_epi[cq * 4 + k]; bs[k] = s_epi[128 + cq * 4 + k]; ns[k] = s_epi[256 + cq * 4 + k]; }
        const float alpha = p.alpha;
        // output pointer of (oy0, ox), advanced by one row per output row: no 64-bit index math inside the loops
        __nv_bfloat16* yrow = p.y + (((long long)n * p.y_img_pitch + (long long)oy0 * p.y_row_pitch + ox) * p.y_cs + cq * 4);
        const long long yrow_step = p.y_row_pitch * p.y_cs;
        const int ycs = p.y_cs;
        mbar_wait(smem_u32(&full[b]), (uint32_t)((it >> 1) & 1));
        const uint32_t tb = smem_u32(bufs + b * BUF_STRIDE) + (uint32_t)(cq * 8 + px * 256);   // 32-bit shared address: immediate offsets
        auto ldi = [&](uint32_t off, float (&v)[4]) {
            uint32_t x, y2;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y2) : "r"(tb + off));
            v[0] = __uint_as_float(x << 16); v[1] = __uint_as_float(x & 0xffff0000u);
            v[2] = __uint_as_float(y2 << 16); v[3] = __uint_as_float(y2 & 0xffff0000u);
        };
        auto body = [&](auto full_c) {
            constexpr bool FULL = decltype(full_c)::value;
            auto finish = [&](int r, int j, float (&acc)[4]) {
                if (!FULL && (oy0 + r >= p.OH || ox + j >= p.OW)) return;
                const float nzg = s_nzb[r * FT_OW + px + j] * ngain;
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float a = fmaf(acc[k], sc[k], nzg + bs[k]);
                    if (fold) a = fmaxf(a, a * alpha);
                    else a *= ((a > 0.f) ? 1.f : alpha) * g_post;
                    a = fminf(fmaxf(a, -clamp_hi), clamp_hi);
                    o[k] = a * ns[k];
                }
                uint2 outv;
                *reinterpret_cast<__nv_bfloat162*>(&outv.x) = __floats2bfloat162_rn(o[0], o[1]);
                *reinterpret_cast<__nv_bfloat162*>(&outv.y) = __floats2bfloat162_rn(o[2], o[3]);
                __stcs(reinterpret_cast<uint2*>(yrow + j * ycs), outv);
            };
            if (PACKED && sep && p.debug == 2) {
                // timing experiment: the tile is waited for and one vector per output is read and stored -- loads and stores, no FIR
#pragma unroll
                for (int r = 0; r < FT_OH; ++r) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint32_t x, y2;
                        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y2) : "r"(tb + (uint32_t)(((r + 1) * FT_IW + j + 1) * 256)));
                        uint2 outv; outv.x = x; outv.y = y2;
                        __stcs(reinterpret_cast<uint2*>(yrow + j * ycs), outv);
                    }
                    yrow += yrow_step;
                }
            } else if (PACKED && sep) {
                // the same arithmetic on register pairs: FFMA2 / FMUL2 / FADD2 halve the FP32-pipe instruction count
                // (6.75 instead of 13.5 per output), which is what bounds this pass
                const float2 fx2[4] = {{fx[0], fx[0]}, {fx[1], fx[1]}, {fx[2], fx[2]}, {fx[3], fx[3]}};
                const float2 fy2[4] = {{fy[0], fy[0]}, {fy[1], fy[1]}, {fy[2], fy[2]}, {fy[3], fy[3]}};
                const float2 sc2[2] = {{sc[0], sc[1]}, {sc[2], sc[3]}}, bs2[2] = {{bs[0], bs[1]}, {bs[2], bs[3]}};
                const float2 ns2[2] = {{ns[0], ns[1]}, {ns[2], ns[3]}}, alpha2 = {alpha, alpha};
                auto finish2 = [&](int r, int j, float2 (&acc)[2]) {
                    if (!FULL && (oy0 + r >= p.OH || ox + j >= p.OW)) return;
                    const float nzg = s_nzb[r * FT_OW + px + j] * ngain;
                    const float2 nz2 = {nzg, nzg};
                    float2 o[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float2 a = fma2(acc[q], sc2[q], add2(nz2, bs2[q]));
                        if (fold) {
                            const float2 m = mul2(a, alpha2);
                            a.x = fmaxf(a.x, m.x); a.y = fmaxf(a.y, m.y);
                        } else {
                            a.x *= ((a.x > 0.f) ? 1.f : alpha) * g_post; a.y *= ((a.y > 0.f) ? 1.f : alpha) * g_post;
                        }
                        a.x = fminf(fmaxf(a.x, -clamp_hi), clamp_hi); a.y = fminf(fmaxf(a.y, -clamp_hi), clamp_hi);
                        o[q] = mul2(a, ns2[q]);
                    }
                    uint2 outv;
                    *reinterpret_cast<__nv_bfloat162*>(&outv.x) = __floats2bfloat162_rn(o[0].x, o[0].y);
                    *reinterpret_cast<__nv_bfloat162*>(&outv.y) = __floats2bfloat162_rn(o[1].x, o[1].y);
                    if (p.debug != 1) __stcs(reinterpret_cast<uint2*>(yrow + j * ycs), outv);
                };
                auto hrow2 = [&](float2 (&dst)[2][2], int r) {
                    float2 v[5][2];
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        uint32_t x, y2;
                        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y2) : "r"(tb + (uint32_t)((r * FT_IW + c) * 256)));
                        v[c][0] = bf16x2_to_f2(x); v[c][1] = bf16x2_to_f2(y2);
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            dst[j][q] = fma2(fx2[3], v[j + 3][q], fma2(fx2[2], v[j + 2][q], fma2(fx2[1], v[j + 1][q], mul2(fx2[0], v[j][q]))));
                };
                if (fresh) { hrow2(h[0], 0); hrow2(h[1], 1); hrow2(h[2], 2); }      // else: rows 8, 9, 10 of the tile above, already in h[0..2]
#pragma unroll
                for (int r = 0; r < FT_OH; ++r) {
                    hrow2(h[(r + 3) & 3], r + 3);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float2 acc[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            acc[q] = fma2(fy2[3], h[(r + 3) & 3][j][q], fma2(fy2[2], h[(r + 2) & 3][j][q],
                                     fma2(fy2[1], h[(r + 1) & 3][j][q], mul2(fy2[0], h[r & 3][j][q]))));
                        finish2(r, j, acc);
                    }
                    yrow += yrow_step;
                }
            } else if (!PACKED && sep) {
                float h[4][2][4];
                auto hrow = [&](float (&dst)[2][4], int r) {
                    float v[5][4];
#pragma unroll
                    for (int c = 0; c < 5; ++c) ldi((uint32_t)((r * FT_IW + c) * 256), v[c]);
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int k = 0; k < 4; ++k) dst[j][k] = fx[0] * v[j][k] + fx[1] * v[j + 1][k] + fx[2] * v[j + 2][k] + fx[3] * v[j + 3][k];
                };
                hrow(h[0], 0); hrow(h[1], 1); hrow(h[2], 2);
#pragma unroll
                for (int r = 0; r < FT_OH; ++r) {
                    hrow(h[(r + 3) & 3], r + 3);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float acc[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            acc[k] = fy[0] * h[r & 3][j][k] + fy[1] * h[(r + 1) & 3][j][k] + fy[2] * h[(r + 2) & 3][j][k] + fy[3] * h[(r + 3) & 3][j][k];
                        finish(r, j, acc);
                    }
                    yrow += yrow_step;
                }
            } else if (!PACKED) {
#pragma unroll
                for (int r = 0; r < FT_OH; ++r) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int bb = 0; bb < 4; ++bb) {
                                float v[4];
                                ldi((uint32_t)(((r + a) * FT_IW + j + bb) * 256), v);
#pragma unroll
                                for (int k = 0; k < 4; ++k) acc[k] = fmaf(f[a * 4 + bb], v[k], acc[k]);
                            }
                        finish(r, j, acc);
                    }
                    yrow += yrow_step;
                }
            }
        };
        if (full_tile) body(std::true_type{}); else body(std::false_type{});
        __syncthreads();                                            // everyone is done with buffer b (and s_epi) before it is refilled
    }
}

// the general path (scalar arithmetic, any 4x4 filter) as an out-of-line function: the packed kernel falls back to it for a
// filter that is not rank-1 without paying for its code paths in registers (inlined next to the packed path they cost it
// 472 bytes of spills and 15 % of its speed)
__device__ __noinline__ void fir_tiled_generic(const CUtensorMap& tmap_t, const CUtensorMap& tmap_t8, const CUtensorMap& tmap_nz,
                                               const FirParams& p, int tiles_x, int tiles_y, int total_tiles, int noise_mode) {
    fir_tiled_body<false>(tmap_t, tmap_t8, tmap_nz, p, tiles_x, tiles_y, total_tiles, noise_mode, 0);
}

template <bool PACKED>
__global__ void __launch_bounds__(256, 2)
fir_act_tiled_kernel(const __grid_constant__ CUtensorMap tmap_t, const __grid_constant__ CUtensorMap tmap_t8, const __grid_constant__ CUtensorMap tmap_nz,
                     const FirParams p, int tiles_x, int tiles_y, int total_tiles, int noise_mode, int strip_ok) {
    if (PACKED) {
        // rank-1 test on the flipped, scaled filter (the same test the body repeats): decided before anything is set up
        pdl_trigger();
        pdl_wait();                                                 // (the filter may be the output of the kernel before this one)
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __ldg(p.f + (15 - i)) * p.fgain;
        bool sep = f[0] != 0.f;
#pragma unroll
        for (int a = 1; a < 4; ++a)
#pragma unroll
            for (int b = 1; b < 4; ++b) sep = sep && fabsf(f[a * 4 + b] * f[0] - f[a * 4] * f[b]) <= 1e-6f * fabsf(f[a * 4 + b] * f[0]) + 1e-30f;
        if (!sep) {
            fir_tiled_generic(tmap_t, tmap_t8, tmap_nz, p, tiles_x, tiles_y, total_tiles, noise_mode);
            return;
        }
    }
    fir_tiled_body<PACKED>(tmap_t, tmap_t8, tmap_nz, p, tiles_x, tiles_y, total_tiles, noise_mode, strip_ok);
}


}  // namespace nbe

using namespace nbe;

extern "C" int nbe_fir_act_nhwc_bf16(const void* t, const float* f, void* y, int N, int OH, int OW, int C, int TH, int TW, int pad,
                                     int t_cs, int64_t t_row_pitch, int64_t t_img_pitch,
                                     int y_cs, int64_t y_row_pitch, int64_t y_img_pitch, float fgain,
                                     const float* scale, const float* noise, int64_t noise_sn, float noise_gain, const float* bias,
                                     float alpha, float gain, float clamp, const float* next_scale, nbe_stream_t stream) {
    NBE_REQUIRE(t && f && y && N >= 0 && OH >= 1 && OW >= 1 && C >= 8 && C % 8 == 0, "fir_act_nhwc: bad arguments");
    NBE_REQUIRE(TH >= 1 && TW >= 1 && pad >= 0, "fir_act_nhwc: bad input extent");
    NBE_REQUIRE(OH == TH + 2 * pad - 3 && OW == TW + 2 * pad - 3, "fir_act_nhwc: output %dx%d does not match input %dx%d, pad %d, 4x4 filter", OH, OW, TH, TW, pad);
    NBE_REQUIRE(t_cs % 8 == 0 && t_cs >= C && y_cs % 8 == 0 && y_cs >= C, "fir_act_nhwc: channel strides must be multiples of 8");
    NBE_REQUIRE(t_row_pitch >= TW && t_img_pitch >= t_row_pitch * TH && y_row_pitch >= OW && y_img_pitch >= y_row_pitch * OH, "fir_act_nhwc: bad pitches");
    NBE_REQUIRE((((uintptr_t)t | (uintptr_t)y) & 15) == 0, "fir_act_nhwc: tensors must be 16-byte aligned");
    if (N == 0) return NBE_OK;
    FirParams p;
    p.t = (const __nv_bfloat16*)t; p.y = (__nv_bfloat16*)y; p.f = f; p.N = N; p.OH = OH; p.OW = OW; p.C = C; p.TH = TH; p.TW = TW; p.pad = pad;
    p.t_cs = t_cs; p.t_row_pitch = t_row_pitch; p.t_img_pitch = t_img_pitch; p.y_cs = y_cs; p.y_row_pitch = y_row_pitch; p.y_img_pitch = y_img_pitch;
    p.fgain = fgain; p.scale = scale; p.noise = noise; p.noise_sn = noise_sn; p.noise_gain = noise_gain; p.bias = bias;
    p.alpha = alpha; p.gain = gain; p.clamp = clamp; p.next_scale = next_scale;
    static const int fir_debug = getenv("NBE_FIR_DEBUG") ? atoi(getenv("NBE_FIR_DEBUG")) : 0;
    p.debug = fir_debug;
    static const bool force_simple = getenv("NBE_FIR_SIMPLE") != nullptr;
    if (C == 128 && t_cs == 128 && !force_simple) {
        // T is described to TMA with its VALID extent, so halo reads outside [0,TH) x [0,TW) come back as zeros
        CUtensorMap tm;
        cuuint64_t dims[4] = {128, (cuuint64_t)TW, (cuuint64_t)TH, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)t_cs * 2, (cuuint64_t)t_row_pitch * t_cs * 2, (cuuint64_t)t_img_pitch * t_cs * 2};
        cuuint32_t box[4] = {128, FT_IW, FT_IH, 1};
        int st = make_tmap(&tm, t, 4, dims, strides, box, "FIR input", 1, /*swizzle=*/0);
        if (st) return st;
        CUtensorMap tm8;                                                // the same tensor, boxes of FT_OH rows: tiles that continue a strip
        cuuint32_t box8[4] = {128, FT_IW, FT_OH, 1};
        st = make_tmap(&tm8, t, 4, dims, strides, box8, "FIR input (strip)", 1, /*swizzle=*/0);
        if (st) return st;
        static const int strip_env = getenv("NBE_FIR_NO_STRIP") == nullptr;              // A/B switch: every tile loads its own halo rows
        static const int cyclic_env = getenv("NBE_FIR_CYCLIC") ? atoi(getenv("NBE_FIR_CYCLIC")) : 1;   // A/B switch: 0 = contiguous runs only
        const int tiles_x = (OW + FT_OW - 1) / FT_OW, tiles_y = (OH + FT_OH - 1) / FT_OH;
        const int64_t total = (int64_t)tiles_x * tiles_y * N;
        int strip_ok = strip_env;
        {   // cyclic strip order where whole strips balance as well as tile runs do (128^2 at batch 256: 7 strips = 112 tiles vs 111)
            const int64_t g = std::min<int64_t>(total, (int64_t)kNumSMs * 2), strips = (int64_t)tiles_x * N;
            const int64_t run = (total + g - 1) / g, cyc = (strips + g - 1) / g * tiles_y;
            if (strip_env && cyclic_env && tiles_x > 1 && strips >= g && cyc * 50 <= run * 51) strip_ok = 2;
        }
        NBE_REQUIRE(total <= INT32_MAX, "fir_act_nhwc: too many tiles");
        const size_t smem = 128 + 2 * ((FT_TILE_BYTES + 127) & ~127) + (2 * FT_OH * FT_OW + 3 * 128 + 16) * sizeof(float) + 64;
        static std::once_flag once;
        static cudaError_t err = cudaSuccess;
        std::call_once(once, [] {
            err = cudaFuncSetAttribute(fir_act_tiled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            if (err == cudaSuccess) err = cudaFuncSetAttribute(fir_act_tiled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        });
        if (err != cudaSuccess) return fail(NBE_ECUDA, "fir_act_nhwc: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
        // noise [N or 1][OH][OW] fp32 as a 3-D tensor map (8 x 16 boxes), when TMA can describe it
        static const bool noise_by_threads = getenv("NBE_FIR_NOISE_THREADS") != nullptr;     // A/B switch
        int noise_mode = noise ? 0 : 2;
        CUtensorMap tn = tm;
        if (noise && !noise_by_threads && ((uintptr_t)noise & 15) == 0 && OW % 4 == 0 &&
            (noise_sn == 0 || (noise_sn % 4 == 0 && noise_sn >= (int64_t)OH * OW))) {
            cuuint64_t ndims[3] = {(cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)(noise_sn ? N : 1)};
            cuuint64_t nstr[2] = {(cuuint64_t)OW * 4, (cuuint64_t)(noise_sn ? noise_sn : (int64_t)OH * OW) * 4};
            cuuint32_t nbox[3] = {FT_OW, FT_OH, 1};
            st = make_tmap(&tn, noise, 3, ndims, nstr, nbox, "FIR noise", 1, /*swizzle=*/0, /*f32=*/1);
            if (st) return st;
            noise_mode = 1;
        }
        static const int grid_env = getenv("NBE_FIR_GRID") ? atoi(getenv("NBE_FIR_GRID")) : 0;     // timing experiment: fewer CTAs
        int grid = grid_env > 0 ? std::min(grid_env, kNumSMs * 2) : kNumSMs * 2;
        if (total < grid) grid = (int)total;
        static const bool scalar_fp32 = getenv("NBE_FIR_SCALAR") != nullptr;      // A/B switch: the unpacked FP32 arithmetic
        if (scalar_fp32) launch_pdl(fir_act_tiled_kernel<false>, dim3(grid), dim3(256), smem, (cudaStream_t)stream, tm, tm8, tn, p, tiles_x, tiles_y, (int)total, noise_mode, strip_ok);
        else launch_pdl(fir_act_tiled_kernel<true>, dim3(grid), dim3(256), smem, (cudaStream_t)stream, tm, tm8, tn, p, tiles_x, tiles_y, (int)total, noise_mode, strip_ok);
        return launched("fir_act_tiled_kernel");
    }
    p.row_groups = (OH + FIR_RPT - 1) / FIR_RPT;
    const int64_t gx = (int64_t)N * p.row_groups;
    NBE_REQUIRE(gx <= INT32_MAX, "fir_act_nhwc: too many row groups");
    const int per_row = OW * (C / FIR_CPT);
    dim3 grid((unsigned)gx, (per_row + 255) / 256);
    NBE_REQUIRE(grid.y <= 65535u, "fir_act_nhwc: rows too wide");
    fir_act_nhwc_kernel<<<grid, 256, 3 * C * sizeof(float), (cudaStream_t)stream>>>(p);
    return launched("fir_act_nhwc_kernel");
}
