// upfirdn2d forward.
//
//   y[n,c,oy,ox] = gain * sum_{a<fh,b<fw} ft[a,b] * xhat[oy*downy + a - pady0, ox*downx + b - padx0]
//   xhat[u,v] = x[n,c,u/upy,v/upx] when u,v >= 0, divisible by the up factors and in range, else 0
//   ft = f flipped in both axes unless `flip`                       
//
// Two kernels:
//  * upfirdn2d_staged_kernel<T, UP>: the cases the generator / public helpers hit -- 4x4 filter, down = 1,
//    up in {1, 2}, dense NCHW (see the comment above the kernel).  HBM-bound: algorithmic bytes = numel(x) + numel(y) elements.
//  * upfirdn2d_generic_kernel<T>: any filter size / up / down / padding / strides (incl. channels_last),
//    one thread per output element.
#include "common.cuh"
#include <type_traits>
#include <cstdlib>

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

// ---- staged 4x4 kernels (dense NCHW, down = 1, up in {1, 2}) ---------------------------------------
// A work item is either a strip of output rows of one (n,c) plane or a group of whole small planes.  Either way the
// input it needs is ONE contiguous range of the flat tensor, so it reaches shared memory as one 1-D bulk copy of aligned
// 16-byte chunks (no per-element index arithmetic, misaligned odd-length rows like 2R+1 do not matter).  up = 1: CTAs are
// persistent and double-buffered (the next item's copy is in flight while the current one is filtered); up = 2: one item
// per CTA.  Each thread produces VPT consecutive columns x RPT rows from a sliding register window; rank-1 filters
// (everything setup_filter builds from a 1-D tap list) take the separable path (4 + 4 taps per output instead of 16; as
// FFMA2 on column pairs for up = 1 with paddings <= 3).  Stores are 16-byte streaming vectors.
// HBM-bound: algorithmic bytes = (numel(x) + numel(y)) * sizeof(T).
constexpr int ST_THREADS = 128;
constexpr int ST_RPT = 8;                                               // output rows per thread

template <class T> struct VecOut;                                       // VPT outputs packed into one 16-byte store
template <> struct VecOut<float> { static constexpr int VPT = 8; };          // two 16-byte stores
template <> struct VecOut<__half> { static constexpr int VPT = 8; };
template <> struct VecOut<__nv_bfloat16> { static constexpr int VPT = 8; };

template <class T, int VPT>
__device__ __forceinline__ void store_row(T* dst, const float (&acc)[VPT], int n_valid) {
    constexpr int PER16 = 16 / (int)sizeof(T);                          // elements per 16-byte store
    if (n_valid == VPT && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int v0 = 0; v0 < VPT; v0 += PER16) {
            int4 v;
            T* e = reinterpret_cast<T*>(&v);
#pragma unroll
            for (int k = 0; k < PER16; ++k) e[k] = Cvt<T>::st(acc[v0 + k]);
            st_stream16(dst + v0, v);
        }
    } else {
#pragma unroll
        for (int k = 0; k < VPT; ++k) if (k < n_valid) dst[k] = Cvt<T>::st(acc[k]);
    }
}

// raw register image of one element (for the AND-mask fix-ups) and its exact conversion to float
template <class T> struct Raw;
template <> struct Raw<float> {
    static __device__ __forceinline__ uint32_t ld(const float* p) { return __float_as_uint(*p); }
    static __device__ __forceinline__ float to_float(uint32_t r) { return __uint_as_float(r); }
    static __device__ __forceinline__ void unpack2(uint32_t, float& lo, float& hi) { lo = hi = 0.f; }   // (16-bit types only)
};
template <> struct Raw<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t ld(const __nv_bfloat16* p) { return *reinterpret_cast<const unsigned short*>(p); }
    static __device__ __forceinline__ float to_float(uint32_t r) { return __uint_as_float(r << 16); }
    static __device__ __forceinline__ void unpack2(uint32_t v, float& lo, float& hi) { lo = __uint_as_float(v << 16); hi = __uint_as_float(v & 0xffff0000u); }
};
template <> struct Raw<__half> {
    static __device__ __forceinline__ uint32_t ld(const __half* p) { return *reinterpret_cast<const unsigned short*>(p); }
    static __device__ __forceinline__ float to_float(uint32_t r) { return __half2float(__ushort_as_half((unsigned short)r)); }
    static __device__ __forceinline__ void unpack2(uint32_t v, float& lo, float& hi) {
        const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&v));
        lo = f2.x; hi = f2.y;
    }
};

struct StagedGeom {
    int cg;             // column groups per row  = ceil(OW / VPT)
    int rg;             // row groups per CTA-plane = ceil(rows_per_cta / RPT)
    int ppc;            // whole planes per CTA (>= 1); > 1 only when a plane fits one CTA
    int strip;          // output rows per CTA when ppc == 1
    int strips;         // strips per plane
    int smem_bytes;     // one input buffer: guard chunk + staged range + tail guard
    int buf_bytes;      // smem_bytes rounded up to 128
    int lds32;          // 16-bit types on the packed path: aligned 32-bit shared-memory loads + funnel shift
    int packed;         // up = 1, windows hang at most 3 elements over a row end, OW % VPT == 0: FFMA2 path with mask fix-ups
    int cg_sh, rg_sh, strips_sh;   // log2 of cg / rg / strips when they are powers of two (-1 otherwise): the per-thread index
                                   // arithmetic was a third of this kernel's instructions while its divisors were runtime values
};
__device__ __forceinline__ void divmod_sh(int v, int d, int sh, int& q, int& r) {
    if (sh >= 0) { q = v >> sh; r = v & (d - 1); } else { q = v / d; r = v - q * d; }
}

template <int VPT, int WIN, int PI>
__device__ __forceinline__ void up2_row(float (&acc)[VPT], const float (&rowA)[WIN], const float (&rowB)[WIN], const float (&f)[16], int a0) {
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        constexpr int dummy = 0; (void)dummy;
        const int b0 = (PI + k) & 1;                                    // compile-time after unrolling
        const int j0 = (k + PI + b0) >> 1;
        const float fa0 = a0 ? f[4 + b0] : f[b0];
        const float fa1 = a0 ? f[6 + b0] : f[2 + b0];
        const float fb0 = a0 ? f[12 + b0] : f[8 + b0];
        const float fb1 = a0 ? f[14 + b0] : f[10 + b0];
        acc[k] = fa0 * rowA[j0] + fa1 * rowA[j0 + 1] + fb0 * rowB[j0] + fb1 * rowB[j0 + 1];
    }
}

template <class T, int UP, int VPT>
__global__ void __launch_bounds__(ST_THREADS)
upfirdn2d_staged_kernel(UpfirdnParams p, StagedGeom g, int n_planes, int n_items) {
    // Persistent CTA, two input buffers: the contiguous input range of the NEXT work item is in flight (one 1-D bulk copy,
    // cp.async.bulk -> mbarrier) while the current one is filtered, so no thread ever waits on a global load of its own.
    extern __shared__ __align__(128) uint8_t st_smem[];                 // [2][g.buf_bytes] input buffers, then 2 mbarriers
    __shared__ float s_f[16];
    __shared__ __align__(16) uint32_t s_zero[8];                        // up = 2: what a window row outside the staged range reads
    uint64_t* bars = reinterpret_cast<uint64_t*>(st_smem + 2 * (size_t)g.buf_bytes);
    const T* xbase = (const T*)p.x;
    const long long plane_elems = (long long)p.H * p.W;
    const uintptr_t t_lo = reinterpret_cast<uintptr_t>(xbase), t_hi = reinterpret_cast<uintptr_t>(xbase + (long long)n_planes * plane_elems);

    struct Item { int plane0, oy0, rows, planes, r_lo, r_hi; uintptr_t gb_lo, a_lo; int n_chunks; };
    auto item_geom = [&](int item, Item& I) {
        // ---- which planes / rows does this item cover
        if (g.ppc > 1) { I.plane0 = item * g.ppc; I.oy0 = 0; I.rows = p.OH; }
        else { int si; divmod_sh(item, g.strips, g.strips_sh, I.plane0, si); I.oy0 = si * g.strip; I.rows = min(g.strip, p.OH - I.oy0); }
        I.planes = min(g.ppc, n_planes - I.plane0);
        // ---- input rows needed (per plane): UP=1: [oy0 - pad, oy0 - pad + rows + 3) ; UP=2: [(oy0 - pad) >> 1, ((oy0 + rows + 2 - pad) >> 1) + 1)
        int r_lo, r_hi;
        if (UP == 1) { r_lo = I.oy0 - p.pady0; r_hi = r_lo + I.rows + 3; }
        else { r_lo = (I.oy0 - p.pady0) >> 1; r_hi = ((I.oy0 + I.rows + 2 - p.pady0) >> 1) + 1; }
        r_lo = max(r_lo, 0); r_hi = min(r_hi, p.H);
        if (g.ppc > 1) { r_lo = 0; r_hi = p.H; }
        I.r_lo = r_lo; I.r_hi = r_hi;
        const long long e_lo = (long long)I.plane0 * plane_elems + (long long)r_lo * p.W;             // first element needed
        const long long e_hi = (g.ppc > 1) ? (long long)(I.plane0 + I.planes) * plane_elems
                                           : (long long)I.plane0 * plane_elems + (long long)max(r_hi, r_lo) * p.W;
        // ---- the range [e_lo, e_hi) in aligned 16-byte chunks; buffer byte 16 <-> global byte a_lo (chunk 0 is a guard)
        I.gb_lo = reinterpret_cast<uintptr_t>(xbase + e_lo);
        const uintptr_t gb_hi = reinterpret_cast<uintptr_t>(xbase + e_hi);
        I.a_lo = I.gb_lo & ~uintptr_t(15);
        I.n_chunks = (int)((gb_hi - I.a_lo + 15) >> 4);
    };
    // all threads call this: thread 0 issues the bulk copy of the chunks that lie completely inside the tensor, the (at most
    // two) chunks straddling its first / last bytes are assembled element by element with zeros outside
    auto stage = [&](int item, int b, Item& I) {
        item_geom(item, I);
        uint8_t* buf = st_smem + (size_t)b * g.buf_bytes;
        int c0 = 0, c1 = I.n_chunks;
        if (I.a_lo < t_lo) c0 = min(c1, 1);                              // a_lo > t_lo - 16
        if (c1 > c0 && I.a_lo + ((uintptr_t)c1 << 4) > t_hi) --c1;       // the last chunk reaches past the tensor
        if (threadIdx.x == 0) {
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[b]);
            const uint32_t bytes = (uint32_t)(c1 - c0) << 4;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
            if (bytes)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"((uint32_t)__cvta_generic_to_shared(buf + 16 + ((size_t)c0 << 4))), "l"(I.a_lo + ((uintptr_t)c0 << 4)), "r"(bytes), "r"(bar)
                             : "memory");
        }
        const int n_edge = c0 + (I.n_chunks - c1);                       // 0, 1 or 2
        if ((int)threadIdx.x < n_edge) {
            const int i = ((int)threadIdx.x == 0 && c0 == 1) ? 0 : c1;
            const uintptr_t ga = I.a_lo + ((uintptr_t)i << 4);
            for (int k = 0; k < 16 / (int)sizeof(T); ++k) {
                const uintptr_t ea = ga + k * sizeof(T);
                reinterpret_cast<T*>(buf)[(1 + i) * (16 / (int)sizeof(T)) + k] =
                    (ea >= t_lo && ea + sizeof(T) <= t_hi) ? *reinterpret_cast<const T*>(ea) : Cvt<T>::st(0.f);
            }
        }
    };
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bars[i])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 16) {
        int a = threadIdx.x >> 2, b = threadIdx.x & 3;
        s_f[threadIdx.x] = (p.flip ? p.f[a * 4 + b] : p.f[(3 - a) * 4 + (3 - b)]) * p.gain;
    }
    if (threadIdx.x < 8) s_zero[threadIdx.x] = 0u;
    __syncthreads();
    Item I_next{};
    if ((int)blockIdx.x < n_items) stage(blockIdx.x, 0, I_next);
    // the (at most two) edge chunks of the first item are written by threads 0 / 1 with ordinary stores: the mbarrier only
    // covers the bulk copy, so the other threads need this barrier before they read them (later items are staged one
    // iteration ahead and ordered by the barrier that ends every iteration).  Found by compute-sanitizer (round 2):
    // under its timing the last chunk of a plane was read before it was written.
    __syncthreads();
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = s_f[i];
    // rank-1 test: f[a][b] * f[0][0] == f[a][0] * f[0][b]  ->  f = fy (x) fx with fy[a] = f[a][0] / f[0][0], fx[b] = f[0][b]
    bool sep = f[0] != 0.f;
#pragma unroll
    for (int a = 1; a < 4; ++a)
#pragma unroll
        for (int b = 1; b < 4; ++b) sep = sep && fabsf(f[a * 4 + b] * f[0] - f[a * 4] * f[b]) <= 1e-6f * fabsf(f[a * 4 + b] * f[0]) + 1e-30f;
    float fx[4], fy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { fx[i] = f[i]; fy[i] = sep ? f[i * 4] / f[0] : 0.f; }

    // chunk 0 of a buffer is a guard: edge threads over-read up to 3 elements to the left of the staged range (and to the
    // right, into the tail guard) and then clear what lies outside the row -- no per-element predicate on the loads
    auto compute = [&](const Item& I, const uint8_t* buf) {
    const int plane0 = I.plane0, oy0 = I.oy0, rows = I.rows, planes = I.planes, r_lo = I.r_lo, r_hi = I.r_hi;
    const T* sx = reinterpret_cast<const T*>(buf + 16 + (I.gb_lo - I.a_lo));                             // smem view of element e_lo

    // ---- thread -> (plane_local, row group, column group)
    int t, cgi, rgi, pl;
    divmod_sh((int)threadIdx.x, g.cg, g.cg_sh, t, cgi);
    divmod_sh(t, g.rg, g.rg_sh, pl, rgi);
    if (pl >= planes) return;
    const int ox0 = cgi * VPT;
    const int ry0 = oy0 + rgi * ST_RPT;
    if (ox0 >= p.OW || ry0 >= oy0 + rows) return;
    const int n_valid = min(VPT, p.OW - ox0);
    const T* sp = sx + (long long)pl * plane_elems - (long long)r_lo * p.W;                             // sp[iy * W + ix] for iy in [r_lo, r_hi)
    T* yp = (T*)p.y + ((long long)(plane0 + pl) * p.OH) * p.OW;
    const int row_end = min(ry0 + ST_RPT, oy0 + rows);

    if (UP == 1) {
        constexpr int WIN = VPT + 3;
        const int ix0 = ox0 - p.padx0;
        const int iy0 = ry0 - p.pady0;
        const int W = p.W;
        // loads are unconditional (guards on both sides of the staged range); only threads whose window hangs over a row
        // end zero the 1..3 elements outside it, rows outside the staged range are zero rows
        const int jl = ix0 < 0 ? -ix0 : 0;                                  // first valid j
        const int jh = ix0 + WIN > W ? W - ix0 : WIN;                       // one past the last valid j
        auto load_in = [&](float (&dst)[WIN], int iy) {
            if (iy >= r_lo && iy < r_hi) {
                const T* rp = sp + (long long)iy * W + ix0;
#pragma unroll
                for (int j = 0; j < WIN; ++j) dst[j] = Cvt<T>::ld(rp[j]);
                if (jl > 0 || jh < WIN) {
#pragma unroll
                    for (int j = 0; j < WIN; ++j) if (j < jl) dst[j] = 0.f;
#pragma unroll
                    for (int j = 0; j < WIN; ++j) if (j >= jh) dst[j] = 0.f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < WIN; ++j) dst[j] = 0.f;
            }
        };
        if (sep && g.packed) {
            // column pairs (k, k+1) share one FFMA2 / FMUL2: half the FP32-pipe instructions of the scalar path below, same
            // operation order per output (16-bit outputs bit-identical; fp32 within 1 ulp of the scalar build, whose contraction is the compiler's)
            constexpr int VP = VPT / 2;
            const float2 fx2[4] = {{fx[0], fx[0]}, {fx[1], fx[1]}, {fx[2], fx[2]}, {fx[3], fx[3]}};
            const float2 fy2[4] = {{fy[0], fy[0]}, {fy[1], fy[1]}, {fy[2], fy[2]}, {fy[3], fy[3]}};
            // edge fix-up (host guarantees g.packed only when every window hangs at most 3 elements over a row end and
            // OW % VPT == 0): only the first / last 3 window slots can lie outside the row, at compile-time slots, so they are
            // cleared with six row-invariant AND masks on the raw bits -- no compares inside the row loop
            uint32_t mk[6] = {jl > 0 ? 0u : ~0u, jl > 1 ? 0u : ~0u, jl > 2 ? 0u : ~0u,
                              jh <= WIN - 3 ? 0u : ~0u, jh <= WIN - 2 ? 0u : ~0u, jh <= WIN - 1 ? 0u : ~0u};
#pragma unroll
            for (int i = 0; i < 6; ++i) asm volatile("" : "+r"(mk[i]));         // keep them as values (not re-derived predicates)
            const bool edge = jl > 0 || jh < WIN;
            float2 h2[4][VP];
            auto hrow2 = [&](float2 (&dst)[VP], int iy) {
                if (iy < r_lo || iy >= r_hi) {                                  // a zero row (padding above / below the image)
#pragma unroll
                    for (int m = 0; m < VP; ++m) dst[m] = make_float2(0.f, 0.f);
                    return;
                }
                const T* rp = sp + (long long)iy * W + ix0;
                uint32_t raw[WIN];                                              // float bits (or raw 16-bit values) of the window
                if (sizeof(T) == 2 && g.lds32) {
                    // 16-bit types: the lanes of a warp sit 16 bytes apart, so every shared-memory load of a warp costs 4
                    // wavefronts whatever its width -- read the window as 6 aligned 32-bit words instead of 11 halfwords and
                    // realign with a funnel shift (rows of odd pitch start on either halfword of a word)
                    constexpr int NW = (WIN + 1) / 2;
                    const uint32_t a = (uint32_t)__cvta_generic_to_shared(rp);    // 32-bit shared address: immediate offsets
                    const uint32_t sh = (a & 2u) ? 16u : 0u;
                    const uint32_t wa = a & ~3u;
                    uint32_t w[NW];
#pragma unroll
                    for (int k = 0; k < NW; ++k) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[k]) : "r"(wa + 4u * k));
#pragma unroll
                    for (int k = 0; k < NW; ++k) {
                        const uint32_t v = __funnelshift_r(w[k], w[k + 1 < NW ? k + 1 : k], sh);   // elements 2k (low half), 2k+1 (high half)
                        float lo, hi;
                        Raw<T>::unpack2(v, lo, hi);
                        raw[2 * k] = __float_as_uint(lo);
                        if (2 * k + 1 < WIN) raw[2 * k + 1] = __float_as_uint(hi);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < WIN; ++j) raw[j] = __float_as_uint(Raw<T>::to_float(Raw<T>::ld(rp + j)));
                }
                if (edge) {                                                     // +0.0f is all-zero bits: masks work on float bits
                    raw[0] &= mk[0]; raw[1] &= mk[1]; raw[2] &= mk[2];
                    raw[WIN - 3] &= mk[3]; raw[WIN - 2] &= mk[4]; raw[WIN - 1] &= mk[5];
                }
                float in[WIN];
#pragma unroll
                for (int j = 0; j < WIN; ++j) in[j] = __uint_as_float(raw[j]);
#pragma unroll
                for (int m = 0; m < VP; ++m) {
                    const float2 e0 = make_float2(in[2 * m], in[2 * m + 1]), o0 = make_float2(in[2 * m + 1], in[2 * m + 2]);
                    const float2 e1 = make_float2(in[2 * m + 2], in[2 * m + 3]), o1 = make_float2(in[2 * m + 3], in[2 * m + 4]);
                    dst[m] = fma2(fx2[3], o1, fma2(fx2[2], e1, fma2(fx2[1], o0, mul2(fx2[0], e0))));
                }
            };
            hrow2(h2[0], iy0); hrow2(h2[1], iy0 + 1); hrow2(h2[2], iy0 + 2);
            T* yrow = yp + (long long)ry0 * p.OW + ox0;
#pragma unroll
            for (int r = 0; r < ST_RPT; ++r) {
                if (ry0 + r >= row_end) break;
                hrow2(h2[(r + 3) & 3], iy0 + r + 3);
                float acc[VPT];
#pragma unroll
                for (int m = 0; m < VP; ++m) {
                    const float2 a2 = fma2(fy2[3], h2[(r + 3) & 3][m], fma2(fy2[2], h2[(r + 2) & 3][m],
                                      fma2(fy2[1], h2[(r + 1) & 3][m], mul2(fy2[0], h2[r & 3][m]))));
                    acc[2 * m] = a2.x; acc[2 * m + 1] = a2.y;
                }
                store_row<T, VPT>(yrow, acc, n_valid);
                yrow += p.OW;
            }
        } else if (sep) {
            float h[4][VPT];                                                // horizontally filtered input rows (sliding window)
            auto hrow = [&](float (&dst)[VPT], int iy) {
                float in[WIN];
                load_in(in, iy);
#pragma unroll
                for (int k = 0; k < VPT; ++k) dst[k] = fx[0] * in[k] + fx[1] * in[k + 1] + fx[2] * in[k + 2] + fx[3] * in[k + 3];
            };
            hrow(h[0], iy0); hrow(h[1], iy0 + 1); hrow(h[2], iy0 + 2);
            T* yrow = yp + (long long)ry0 * p.OW + ox0;
#pragma unroll
            for (int r = 0; r < ST_RPT; ++r) {
                if (ry0 + r >= row_end) break;
                hrow(h[(r + 3) & 3], iy0 + r + 3);
                float acc[VPT];
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    acc[k] = fy[0] * h[r & 3][k] + fy[1] * h[(r + 1) & 3][k] + fy[2] * h[(r + 2) & 3][k] + fy[3] * h[(r + 3) & 3][k];
                store_row<T, VPT>(yrow, acc, n_valid);
                yrow += p.OW;
            }
        } else {
            float win[4][WIN];
            load_in(win[0], iy0); load_in(win[1], iy0 + 1); load_in(win[2], iy0 + 2);
#pragma unroll
            for (int r = 0; r < ST_RPT; ++r) {
                const int oy = ry0 + r;
                if (oy >= row_end) break;
                load_in(win[(r + 3) & 3], iy0 + r + 3);
                float acc[VPT];
#pragma unroll
                for (int k = 0; k < VPT; ++k) acc[k] = 0.f;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                        for (int k = 0; k < VPT; ++k) acc[k] = fmaf(f[a * 4 + b], win[(r + a) & 3][k + b], acc[k]);
                store_row<T, VPT>(yp + (long long)oy * p.OW + ox0, acc, n_valid);
            }
        }
    } else {
        // up = 2: output row oy reads input rows iyA = (oy - pady0 + a0) >> 1 and iyA + 1 with the taps of row parity a0; the
        // ST_RPT output rows of a thread touch ST_RPT / 2 + 2 input rows, each loaded ONCE into a sliding register window
        constexpr int WIN = VPT / 2 + 2;
        constexpr int NR = ST_RPT / 2 + 2;
        const int ixb = (ox0 - p.padx0) >> 1;
        const int W = p.W;
        const int u00 = ry0 - p.pady0;
        const int iy_first = (u00 + (u00 & 1)) >> 1;                       // first input row of the thread's first output row
        const int jl = ixb < 0 ? -ixb : 0;
        const int jh = ixb + WIN > W ? W - ixb : WIN;
        float rows[NR][WIN];
        // Window loads without branches (ncu, round 2: this path was ISSUE-bound -- 17 instructions per output, two thirds of
        // them compares / selects / branches around the loads and stores): a row outside the staged range reads the zero
        // words in s_zero instead, window slots outside the row are cleared by row-invariant AND masks on the raw bits, and
        // 16-bit types read aligned 32-bit words (a window of 6 halfwords = 4 words) realigned by a funnel shift.
        uint32_t mk[WIN];
#pragma unroll
        for (int j = 0; j < WIN; ++j) mk[j] = (j >= jl && j < jh) ? ~0u : 0u;
        const uint32_t zero_a = (uint32_t)__cvta_generic_to_shared(s_zero);
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const int iy = iy_first + q;
            const bool ok = iy >= r_lo && iy < r_hi;
            const uint32_t a_row = (uint32_t)__cvta_generic_to_shared(sp + (long long)(ok ? iy : r_lo) * W + ixb);
            if (sizeof(T) == 2) {
                constexpr int NW = WIN / 2 + 1;
                const uint32_t sh = (a_row & 2u) ? 16u : 0u;
                const uint32_t wa = ok ? (a_row & ~3u) : zero_a;
                uint32_t w[NW];
#pragma unroll
                for (int k = 0; k < NW; ++k) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[k]) : "r"(wa + 4u * k));
#pragma unroll
                for (int k = 0; k < WIN / 2; ++k) {
                    const uint32_t v = __funnelshift_r(w[k], w[k + 1], sh);      // elements 2k (low half), 2k + 1 (high half)
                    float lo, hi;
                    Raw<T>::unpack2(v, lo, hi);
                    rows[q][2 * k] = __uint_as_float(__float_as_uint(lo) & mk[2 * k]);
                    rows[q][2 * k + 1] = __uint_as_float(__float_as_uint(hi) & mk[2 * k + 1]);
                }
            } else {
                const uint32_t wa = ok ? a_row : zero_a;
#pragma unroll
                for (int j = 0; j < WIN; ++j) {
                    uint32_t v;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(wa + 4u * j));
                    rows[q][j] = __uint_as_float(v & mk[j]);
                }
            }
        }
        // column parity pi = (ox0 - padx0) & 1 is uniform across the CTA (ox0 is a multiple of VPT): with it fixed,
        // the tap parity b0 = (pi + k) & 1 and the window slot j0 = (k + pi + b0) >> 1 are compile-time per k.
        const int pi = (ox0 - p.padx0) & 1;
        const int par0 = u00 & 1;                                          // row parity of the first output row (uniform per launch)
        T* yrow = yp + (long long)ry0 * p.OW + ox0;
        // whole row groups of whole, 16-byte aligned column groups (every thread at the generator's sizes) store without the
        // per-row bounds / alignment checks
        const bool fast = ry0 + ST_RPT <= row_end && n_valid == VPT && (reinterpret_cast<uintptr_t>(yrow) & 15) == 0 &&
                          (((size_t)p.OW * sizeof(T)) & 15) == 0;
        auto out_rows = [&](auto par0_c, auto pi_c, auto fast_c) {
            constexpr int PAR0 = decltype(par0_c)::value, PI = decltype(pi_c)::value;
            constexpr bool FAST = decltype(fast_c)::value;
#pragma unroll
            for (int r = 0; r < ST_RPT; ++r) {
                if (!FAST && ry0 + r >= row_end) break;
                float acc[VPT];
                // q = (u0 + a0) / 2 - iy_first with u0 = u00 + r, a0 = u0 & 1: compile-time per (r, par0)
                const int a0 = (r + PAR0) & 1, q = PAR0 == 0 ? (r + a0) >> 1 : (r + 1 + a0) / 2 - 1;
                up2_row<VPT, WIN, PI>(acc, rows[q], rows[q + 1], f, a0);
                if (FAST) {
                    constexpr int PER16 = 16 / (int)sizeof(T);
#pragma unroll
                    for (int v0 = 0; v0 < VPT; v0 += PER16) {
                        int4 v;
                        T* e = reinterpret_cast<T*>(&v);
#pragma unroll
                        for (int k = 0; k < PER16; ++k) e[k] = Cvt<T>::st(acc[v0 + k]);
                        st_stream16(yrow + v0, v);
                    }
                } else {
                    store_row<T, VPT>(yrow, acc, n_valid);
                }
                yrow += p.OW;
            }
        };
        using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>;
        if (fast) {
            if (par0 == 0) { if (pi == 0) out_rows(I0{}, I0{}, std::true_type{}); else out_rows(I0{}, I1{}, std::true_type{}); }
            else           { if (pi == 0) out_rows(I1{}, I0{}, std::true_type{}); else out_rows(I1{}, I1{}, std::true_type{}); }
        } else {
            if (par0 == 0) { if (pi == 0) out_rows(I0{}, I0{}, std::false_type{}); else out_rows(I0{}, I1{}, std::false_type{}); }
            else           { if (pi == 0) out_rows(I1{}, I0{}, std::false_type{}); else out_rows(I1{}, I1{}, std::false_type{}); }
        }
    }
    };   // compute

    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int b = it & 1;
        const Item I = I_next;                                            // (computed when the item was staged)
        if (item + (int)gridDim.x < n_items) stage(item + gridDim.x, b ^ 1, I_next);   // buffer b^1 was released by the barrier that ended the previous iteration
        {   // bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[b]), parity = (uint32_t)((it >> 1) & 1);
            uint32_t done = 0;
            const long long t0 = clock64();
            while (true) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                if (done) break;
                if (clock64() - t0 > 4000000000LL) { printf("nbe upfirdn2d: mbarrier wait timed out (block %d)\n", (int)blockIdx.x); __trap(); }
            }
        }
        compute(I, st_smem + (size_t)b * g.buf_bytes);
        __syncthreads();                                               // everyone is done with buffer b before it is refilled
    }
}

template <class T>
static bool staged_geometry(const UpfirdnParams& p, StagedGeom& g, int VPT) {
    g.cg = (p.OW + VPT - 1) / VPT;
    if (g.cg > ST_THREADS) return false;
    const int rg_plane = (p.OH + ST_RPT - 1) / ST_RPT;
    const int es = (int)sizeof(T);
    if (g.cg * rg_plane * 2 <= ST_THREADS) {                          // several whole planes per CTA
        g.ppc = ST_THREADS / (g.cg * rg_plane);
        g.rg = rg_plane; g.strip = p.OH; g.strips = 1;
        while (g.ppc > 1 && (long long)g.ppc * p.H * p.W * es + 80 > 48 * 1024) --g.ppc;
        g.smem_bytes = (int)((long long)g.ppc * p.H * p.W * es + 80);
    } else {
        g.ppc = 1;
        g.rg = ST_THREADS / g.cg;
        if (g.rg > rg_plane) g.rg = rg_plane;
        g.strip = g.rg * ST_RPT;
        g.strips = (p.OH + g.strip - 1) / g.strip;
        const int in_rows = (p.upy == 1) ? g.strip + 3 : g.strip / 2 + 3;
        g.smem_bytes = (int)((long long)in_rows * p.W * es + 80);
    }
    g.buf_bytes = (g.smem_bytes + 127) & ~127;
    auto lg = [](int v) { int sh = 0; while ((1 << sh) < v) ++sh; return (1 << sh) == v ? sh : -1; };
    g.cg_sh = lg(g.cg); g.rg_sh = lg(g.rg); g.strips_sh = lg(g.strips);
    return g.smem_bytes <= 48 * 1024;                                  // two buffers per CTA
}

template <class T>
static int run_typed(const UpfirdnParams& p, bool tiled_ok, cudaStream_t s) {
    StagedGeom g;
    // outputs per thread and row: 8 (one or two 16-byte stores); 4-byte types filtering at the input resolution keep 4 --
    // half the shared memory per CTA, twice the CTAs per SM to hide the staging latency
    static const char* vpt_env = getenv("NBE_UPF_VPT");
    int vpt = VecOut<T>::VPT;
    if (sizeof(T) == 4 && p.upx == 1) vpt = 4;
    if (sizeof(T) == 4 && vpt_env) vpt = atoi(vpt_env) == 4 ? 4 : 8;
    static const bool scalar_fp32 = getenv("NBE_UPF_SCALAR") != nullptr;          // A/B switch: the unpacked arithmetic / generic fix-ups
    if (tiled_ok && staged_geometry<T>(p, g, vpt)) {
        const int padx1 = p.OW - p.W - p.padx0 + 3;                               // up = 1: OW = W + padx0 + padx1 - 3
        g.packed = !scalar_fp32 && p.upx == 1 && p.padx0 <= 3 && padx1 <= 3 && p.OW % vpt == 0;
        static const bool lds16 = getenv("NBE_UPF_LDS16") != nullptr;            // A/B switch: halfword shared-memory loads
        g.lds32 = !lds16;
        const int n_planes = p.N * p.C;
        const int64_t blocks = g.ppc > 1 ? (n_planes + g.ppc - 1) / g.ppc : (int64_t)n_planes * g.strips;
        if (blocks <= INT32_MAX) {
            void (*kern)(UpfirdnParams, StagedGeom, int, int);
            if (sizeof(T) == 4 && vpt == 4) kern = (p.upx == 1) ? upfirdn2d_staged_kernel<T, 1, (sizeof(T) == 4 ? 4 : 8)> : upfirdn2d_staged_kernel<T, 2, (sizeof(T) == 4 ? 4 : 8)>;
            else kern = (p.upx == 1) ? upfirdn2d_staged_kernel<T, 1, 8> : upfirdn2d_staged_kernel<T, 2, 8>;
            const int dyn = 2 * g.buf_bytes + 64;                          // two input buffers + mbarriers
            if (dyn > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
                if (e != cudaSuccess) return fail(NBE_ECUDA, "upfirdn2d: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            }
            // persistent grid: as many CTAs as fit an SM (shared memory; at most 6, 12 for 16-bit up = 2), each keeps its next item in flight
            int per_sm = (227 * 1024) / (dyn + 1024);
            const char* cap_env = getenv("NBE_UPF_PER_SM");                 // A/B: persistent CTAs per SM
            const int cap = cap_env ? atoi(cap_env) : ((p.upx == 2 && sizeof(T) == 2) ? 12 : 6);
            per_sm = per_sm < 1 ? 1 : (per_sm > cap ? cap : per_sm);
            // one item per CTA (the loop runs once; residency is then set by registers / shared memory only): FP32 up = 2 writes 4x
            // what it reads, gains nothing from prefetching its small input and loses residency to the persistent grid
            // (tools/ab_up2.py, 64^2 -> 128^2, batch 256 x 128: 0.565 ms one-shot vs 0.62-0.66 ms persistent).  The 16-bit types were
            // ISSUE-bound (ncu: 17 instructions per output, ALU pipe 71 %); with the branch-free window loads and shift / mask
            // index arithmetic the persistent grid is the faster one for them (bf16: 0.258 ms persistent x12 vs 0.314 ms one-shot;
            // 0.382 ms before).  NBE_UPF_ONESHOT=0/1 overrides.
            const char* one_env = getenv("NBE_UPF_ONESHOT");
            const bool oneshot = one_env ? atoi(one_env) != 0 : (p.upx == 2 && sizeof(T) == 4);
            const int grid = (int)(oneshot || blocks < (int64_t)kNumSMs * per_sm ? blocks : (int64_t)kNumSMs * per_sm);
            kern<<<grid, ST_THREADS, dyn, s>>>(p, g, n_planes, (int)blocks);
            return launched("upfirdn2d_staged_kernel");
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
    // the staged kernels size their shared-memory guards (and their edge fix-ups) for small paddings -- the generator's
    // (1,1,1,1), filter2d's / upsample2d's (2,1,2,1), and the swept range of tests/test_ops_gpu.py; larger paddings
    // (a window that starts >= 4 columns left of the row) and vertical crops take the generic kernel
    const bool pads_ok = padx0 >= -2 && padx0 <= 3 && padx1 >= -1 && padx1 <= 3 && pady0 >= 0 && pady0 <= 3 && pady1 >= 0 && pady1 <= 3;
    const bool tiled_ok = dense_x && dense_y && fh == 4 && fw == 4 && downx == 1 && downy == 1 && upx == upy &&
                          (upx == 1 || upx == 2) && pads_ok;
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case NBE_F32:  return run_typed<float>(p, tiled_ok, s);
        case NBE_F16:  return run_typed<__half>(p, tiled_ok, s);
        case NBE_BF16: return run_typed<__nv_bfloat16>(p, tiled_ok, s);
        default: return fail(NBE_EUNSUPPORTED, "upfirdn2d: unsupported dtype %d (float32/float16/bfloat16 only)", dtype);
    }
}
