// Geometry encoder, first layer (simple_autoencoder.py:155-166: Conv2d(1, 64, k=7, padding=3, padding_mode='reflect') + BN +
// LeakyReLU) as a BLOCK-TOEPLITZ implicit GEMM on tcgen05, with no im2col written by threads.
//
// The layer has ONE input channel, so its GEMM has K = 49: too thin to tile the usual way, and building the im2col rows with
// thread loads (enc_conv7x7_tc_kernel) costs ~500 instructions per pixel tile and a barrier chain per tile (0.23 ms at batch
// 256 against a 0.09 ms write floor).  Here a GEMM row is not a pixel but a GROUP OF 8 ADJACENT PIXELS of one image row:
//
//   D[(y, j), (p, c)] = sum_{kh < 7, t < 16}  P[y + kh][8 j + t] * Wt[(p, c), (kh, t)],     Wt[(p, c), (kh, t)] = w[c][kh][t - p]
//
// with P the reflect-padded (3 px), pre-processed bf16 image, p = pixel within the group, c = output channel and Wt zero where
// t - p is outside 0..6.  Row (y, j) of the A operand is then 7 runs of 16 CONSECUTIVE bf16 of P starting at column 8 j -- a
// 16-byte aligned address -- so a whole 128-row A tile (8 image rows x 16 groups = 8 x 128 pixels) is ONE 5-D TMA box over P
// whose dimensions (t, j, y, kh) have the strides (2 B, 16 B, row, row): the boxes of neighbouring groups and of neighbouring
// kernel rows simply overlap in global memory (a tensor map only asks every stride to be a multiple of the one before).
// A run of 16 bf16 is exactly one K = 16 MMA block, so both operands use the 32-BYTE swizzle: per kernel row kh the A tile is
// 128 rows x 32 B and the weight tile 512 rows x 32 B.  K = 7 x 16 = 112; N = 8 px x 64 ch = 512 accumulator columns = two
// N = 256 halves (pixels 0..3 / 4..7 of every group) that alternate between the two halves of TMEM, so the epilogue of one half
// overlaps the MMAs of the next.  2.3x the algorithmic FLOPs, which is nothing (14 MMAs per 1024 pixels); the kernel is bound
// by its 553 MB of output.
//
// Epilogue: TMEM lane = group, 64 consecutive columns = the 64 channels of one pixel = one 128-byte NHWC line: a thread adds the
// bias, applies LeakyReLU, rounds to bf16 and puts the line into a 128-byte-swizzled staging tile [8 rows x 16 groups][128 B];
// one TMA store per (tile, pixel-in-group) writes it out -- a box that takes every 8th pixel of 8 image rows (traversal
// stride 8).  (Storing the lines straight from the owning lanes, 32 bytes per lane, touches 32 different lines per instruction:
// measured 77 % L1TEX utilisation and 2.9 TB/s.)  The 1-pixel reflect border the next (stride-2, padding_mode='reflect') layer
// reads is written with plain stores by the threads that own the mirrored pixels, so no separate border pass follows.
#include "tc_common.cuh"
#include <mutex>
#include <algorithm>

namespace nbe {

constexpr int TP_THREADS = 320;                                      // warps: 0 TMA, 1 MMA, 2..9 epilogue
constexpr int TP_ASTAGES = 1;                                       // the kernel is epilogue-paced: the next tile's 28 KB land long before they are needed
constexpr int TP_STAGE_BYTES = 128 * 128;                            // staging tile of one epilogue group: 128 pixels x 128 B
constexpr int TP_A_KH = 128 * 32;                                    // one kernel row of an A tile: 128 rows x 32 B
constexpr int TP_B_KH = 512 * 32;                                    // one kernel row of the weights: 512 rows x 32 B
constexpr int TP_A_BYTES = 7 * TP_A_KH;
constexpr int TP_B_BYTES = 7 * TP_B_KH;
// K-major SWIZZLE_32B descriptor, high word: SBO = 256 B (8 rows x 32 B) | version 1 | layout type 6
constexpr uint32_t kDescHiSw32 = (uint32_t)(256 >> 4) | (1u << 14) | (6u << 29);
__device__ __forceinline__ void umma_bf16_sw32(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(kDescHiSw32), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// 32-byte store: one whole sector per lane
__device__ __forceinline__ void st_global_32B(void* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// x [N,H,W] float32 -> P [N,H+6,Wp] bf16: reflect padding by 3, BaseGeoEncoder preprocessing (base.py:32-58), zeros beyond W+6.
// One thread = 8 consecutive columns of one padded row (one 16-byte store; Wp is a multiple of 8): the element-per-thread form
// of this kernel took 30 us for 27 MB (three integer divisions per element), on the encoder's critical path.
__global__ void __launch_bounds__(256)
enc7_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ P, int N, int H, int W, int Wp, int preproc) {
    pdl_trigger();
    pdl_wait();
    const int vpr = Wp >> 3;                                              // 16-byte vectors per padded row
    const int64_t total = (int64_t)N * (H + 6) * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % vpr);
        const int64_t t = i / vpr;
        const int r = (int)(t % (H + 6));
        const int n = (int)(t / (H + 6));
        int iy = r - 3;
        iy = iy < 0 ? -iy : iy; iy = iy >= H ? 2 * H - 2 - iy : iy;
        const float* xr = x + ((int64_t)n * H + iy) * W;
        int4 out;
        __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(&out);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = cv * 8 + k;
            float v = 0.f;
            if (c < W + 6) {
                int ix = c - 3;
                ix = ix < 0 ? -ix : ix; ix = ix >= W ? 2 * W - 2 - ix : ix;
                v = __ldg(xr + ix);
                if (preproc == 1) v = 1.f - v;
                else if (preproc == 2) v = (1.f - v) * 2.f - 1.f;
            }
            e[k] = __float2bfloat16_rn(v);
        }
        *reinterpret_cast<int4*>(P + ((int64_t)n * (H + 6) + r) * Wp + cv * 8) = out;
    }
}

// w [Cout,49] float32 (BN folded) -> Wt [7 kh][512 rows = (p, c)][16 t] bf16
__global__ void enc7_toeplitz_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wt, int Cout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 7 * 512 * 16) return;
    const int t = i & 15, n = (i >> 4) & 511, kh = i >> 13;
    const int p = n >> 6, c = n & 63;
    const int kw = t - p;
    float v = 0.f;
    if (c < Cout && kw >= 0 && kw < 7) v = w[c * 49 + kh * 7 + kw];
    wt[i] = __float2bfloat16_rn(v);
}

struct ToepParams {
    __nv_bfloat16* y;                                                 // [N, H+2, W+2, y_cs]
    const float* bias;
    int N, H, W, y_cs, tiles_x, tiles_y, total;
    float neg_slope;
    uint32_t idesc;
};

__global__ void __launch_bounds__(TP_THREADS, 1)
enc7_toeplitz_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_w,
                     const __grid_constant__ CUtensorMap tmap_y, const ToepParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    pdl_trigger();
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_b = smem;                                           // [7 kh][512][32 B]
    uint8_t* smem_a = smem + TP_B_BYTES;                              // [TP_ASTAGES][7 kh][128][32 B]
    uint8_t* smem_stage = smem_a + TP_ASTAGES * TP_A_BYTES;          // [2 groups][2 buffers][128 rows][128 B], 128-byte swizzle
    float* s_bias = reinterpret_cast<float*>(smem_stage + 4 * TP_STAGE_BYTES);      // [64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 64);
    uint64_t* a_full = bars;                      // [TP_ASTAGES]
    uint64_t* a_empty = bars + TP_ASTAGES;        // [TP_ASTAGES]
    uint64_t* acc_full = bars + 2 * TP_ASTAGES;   // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2]
    uint64_t* b_full = acc_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

    const int warp = (int)uniform_u32(threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < TP_ASTAGES; ++i) { mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 8); }
        mbar_init(smem_u32(b_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_p) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_y) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    const int n_local = (p.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    pdl_wait();
    if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias[threadIdx.x];
    __syncthreads();

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            mbar_expect_tx(smem_u32(b_full), TP_B_BYTES);
            for (int q = 0; q < 14; ++q) tma_load_2d(smem_u32(smem_b + q * 256 * 32), &tmap_w, smem_u32(b_full), 0, q * 256);
            for (int i = 0; i < n_local; ++i) {
                int t = (int)blockIdx.x + i * (int)gridDim.x;
                const int tx = t % p.tiles_x; t /= p.tiles_x;
                const int ty = t % p.tiles_y; const int n = t / p.tiles_y;
                const int slot = i % TP_ASTAGES;
                const uint32_t par = (uint32_t)(i / TP_ASTAGES) & 1u;
                mbar_wait_fast(smem_u32(&a_empty[slot]), par ^ 1);
                const uint32_t full = smem_u32(&a_full[slot]);
                mbar_expect_tx(full, TP_A_BYTES);
                // box (t 16, j 16, y 8, kh 7, n 1): per kernel row, 128 rows (y, j) of 16 K elements
                tma_load_5d(smem_u32(smem_a + slot * TP_A_BYTES), &tmap_p, full, 0, tx * 16, ty * 8, 0, n);
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer (whole warp, elected lane issues) ==============================
        const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a)), b_lo0 = umma_desc_lo(smem_u32(smem_b));
        const uint32_t idesc = p.idesc;
        mbar_wait(smem_u32(b_full), 0);
        tcgen05_fence_after();
        uint32_t acc_par[2] = {0, 0};
        for (int i = 0; i < n_local; ++i) {
            const int slot = i % TP_ASTAGES;
            const uint32_t par = (uint32_t)(i / TP_ASTAGES) & 1u;
            mbar_wait_fast(smem_u32(&a_full[slot]), par);
            tcgen05_fence_after();
            const uint32_t a_lo = a_lo0 + (uint32_t)slot * (TP_A_BYTES >> 4);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                mbar_wait_fast(smem_u32(&acc_empty[half]), acc_par[half] ^ 1);
                acc_par[half] ^= 1;
                tcgen05_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(half * 256);
                const uint32_t b_lo = b_lo0 + (uint32_t)half * ((256 * 32) >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int kh = 0; kh < 7; ++kh)
                        umma_bf16_sw32(d, a_lo + (uint32_t)kh * (TP_A_KH >> 4), b_lo + (uint32_t)kh * (TP_B_KH >> 4), idesc, kh != 0);
                    umma_commit(smem_u32(&acc_full[half]));
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(smem_u32(&a_empty[slot]));
            __syncwarp();
        }
    } else {
        // ============================== epilogue (warps 2..9) ==============================
        // TMEM lane quarter qd = warp % 4 (hardware rule); the two groups of four warps take pixels {0, 1} and {2, 3} of a half.
        const int qd = warp & 3, grp = (warp - 2) >> 2;
        const int m = qd * 32 + lane;
        const int yl = m >> 4, j = m & 15;
        const int Hp = p.H + 2, Wp2 = p.W + 2;
        const float slope = p.neg_slope;
        const uint32_t s_bias_u32 = smem_u32(s_bias);
        const bool issuer = ((warp - 2) & 3) == 0 && lane == 0;       // one thread per group issues its TMA stores
        const uint32_t stg0 = smem_u32(smem_stage + grp * 2 * TP_STAGE_BYTES);
        const int sw = m & 7;
        uint32_t buf = 0;                                             // staging tile of the next pixel (two per group, alternating)
        uint32_t acc_phase[2] = {0, 0};
        for (int i = 0; i < n_local; ++i) {
            int t = (int)blockIdx.x + i * (int)gridDim.x;
            const int tx = t % p.tiles_x; t /= p.tiles_x;
            const int ty = t % p.tiles_y; const int n = t / p.tiles_y;
            const int oy = ty * 8 + yl;
            // the reflected border row this pixel row is the source of (-1: none)
            const int my = oy == 1 ? 0 : (oy == p.H - 2 ? Hp - 1 : -1);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                mbar_wait(smem_u32(&acc_full[half]), acc_phase[half]);
                acc_phase[half] ^= 1;
                tcgen05_fence_after();
#pragma unroll 1
                for (int pp = 0; pp < 2; ++pp) {
                    const int pl = grp * 2 + pp;
                    const int ox = (tx * 16 + j) * 8 + half * 4 + pl;
                    uint32_t v0[32], v1[32];
                    const uint32_t ta = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(half * 256 + pl * 64);
                    tmem_ld32_nowait(ta, v0);
                    tmem_ld32_nowait(ta + 32, v1);
                    tmem_ld_wait();
                    if (pp == 1) {                                    // the accumulator half is in registers: hand it back to the MMA warp
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&acc_empty[half])) : "memory");
                    }
                    uint32_t o[4][8];
                    const float2 slope2 = make_float2(slope, slope);
#pragma unroll
                    for (int hv = 0; hv < 2; ++hv) {
                        const uint32_t (&src)[32] = hv ? v1 : v0;
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) {              // 4 channels per step: one 16-byte read of the bias vector
                            float4 b4;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w)
                                         : "r"(s_bias_u32 + (uint32_t)((hv * 32 + q4 * 4) * 4)));
                            const float2 a0 = add2(make_float2(__uint_as_float(src[4 * q4]), __uint_as_float(src[4 * q4 + 1])), make_float2(b4.x, b4.y));
                            const float2 a1 = add2(make_float2(__uint_as_float(src[4 * q4 + 2]), __uint_as_float(src[4 * q4 + 3])), make_float2(b4.z, b4.w));
                            const float2 m0 = mul2(a0, slope2), m1 = mul2(a1, slope2);      // LeakyReLU = max(a, a * slope) for 0 <= slope <= 1
                            const __nv_bfloat162 h0 = __floats2bfloat162_rn(fmaxf(a0.x, m0.x), fmaxf(a0.y, m0.y));
                            const __nv_bfloat162 h1 = __floats2bfloat162_rn(fmaxf(a1.x, m1.x), fmaxf(a1.y, m1.y));
                            o[hv * 2 + (q4 >> 2)][(q4 & 3) * 2] = *reinterpret_cast<const uint32_t*>(&h0);
                            o[hv * 2 + (q4 >> 2)][(q4 & 3) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                        }
                    }
                    // this staging tile was read by the store before the previous one, which the barrier below has already waited for
                    const uint32_t stg = stg0 + buf * TP_STAGE_BYTES;
                    const uint32_t stg_row = stg + (uint32_t)m * 128u;
                    buf ^= 1;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                     :: "r"(stg_row + (uint32_t)(((2 * s) ^ sw) << 4)), "r"(o[s][0]), "r"(o[s][1]), "r"(o[s][2]), "r"(o[s][3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                     :: "r"(stg_row + (uint32_t)(((2 * s + 1) ^ sw) << 4)), "r"(o[s][4]), "r"(o[s][5]), "r"(o[s][6]), "r"(o[s][7]) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    // one barrier per pixel: behind it, all 128 lines are staged AND the previous store has read the other tile
                    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" :: "r"(1 + grp) : "memory");
                    if (issuer) {
                        // interior pixels (x0 + 8 j, ty * 8 + yl): the box walks 128 columns with traversal stride 8
                        tma_store_4d(&tmap_y, stg, 0, tx * 128 + half * 4 + pl, ty * 8, n);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    // reflected border copies of this pixel (rows 1 / H-2 and columns 1 / W-2 are the sources)
                    const int mx = ox == 1 ? 0 : (ox == p.W - 2 ? Wp2 - 1 : -1);
                    if (my >= 0 || mx >= 0) {
                        __nv_bfloat16* img = p.y + (long long)n * Hp * Wp2 * p.y_cs;
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy) {
                            const int ry = dy ? my : oy + 1;
                            if (ry < 0) continue;
#pragma unroll
                            for (int dx = 0; dx < 2; ++dx) {
                                const int rx = dx ? mx : ox + 1;
                                if (rx < 0 || (dy == 0 && dx == 0)) continue;
                                uint8_t* dst = reinterpret_cast<uint8_t*>(img + ((long long)ry * Wp2 + rx) * p.y_cs);
#pragma unroll
                                for (int s = 0; s < 4; ++s) st_global_32B(dst + s * 32, o[s]);
                            }
                        }
                    }
                }
            }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the staging tile outlives its last store
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace nbe

using namespace nbe;

static int toeplitz_image_map(CUtensorMap* tp, const void* scratch, int N, int H, int W) {
    // (t, j, y, kh, n) over P [N][H+6][W+16]: the t / j and the y / kh dimensions overlap in memory; every stride is a multiple
    // of the one before it, as cuTensorMapEncodeTiled asks
    const int Wp = W + 16;
    cuuint64_t dims[5] = {16, (cuuint64_t)W / 8, (cuuint64_t)H, 7, (cuuint64_t)N};
    cuuint64_t strides[4] = {16, (cuuint64_t)Wp * 2, (cuuint64_t)Wp * 2, (cuuint64_t)(H + 6) * Wp * 2};
    cuuint32_t box[5] = {16, 16, 8, 7, 1};
    return make_tmap(tp, scratch, 5, dims, strides, box, "toeplitz image windows", 1, 32);
}

extern "C" int64_t nbe_enc_conv7x7_toeplitz_scratch_bytes(int N, int H, int W) {
    if (N < 0 || H < 1 || W < 1) return -1;
    return (int64_t)N * (H + 6) * (W + 16) * 2;
}

extern "C" int nbe_enc_conv7x7_toeplitz_weights(const float* w, void* wt, int Cout, nbe_stream_t stream) {
    NBE_REQUIRE(w && wt && Cout >= 1 && Cout <= 64, "enc_conv7x7_toeplitz_weights: Cout must be 1..64");
    enc7_toeplitz_weights_kernel<<<(7 * 512 * 16 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wt, Cout);
    return launched("enc7_toeplitz_weights_kernel");
}

extern "C" int nbe_enc_conv7x7_toeplitz_bf16(const float* x, const void* wt, const float* bias, void* y, void* scratch, int64_t scratch_bytes,
                                             int N, int H, int W, int Cout, int y_cs, float neg_slope, int preproc, nbe_stream_t stream) {
    NBE_REQUIRE(x && wt && bias && y && scratch && N >= 0, "enc_conv7x7_toeplitz: bad arguments");
    if (Cout != 64 || y_cs != 64 || H % 8 != 0 || W % 128 != 0 || H < 8)
        return fail(NBE_EUNSUPPORTED, "enc_conv7x7_toeplitz: needs Cout == y_cs == 64, H %% 8 == 0, W %% 128 == 0 (got Cout %d, y_cs %d, H %d, W %d)", Cout, y_cs, H, W);
    NBE_REQUIRE(neg_slope >= 0.f && neg_slope <= 1.f, "enc_conv7x7_toeplitz: LeakyReLU slope must be in [0, 1]");
    NBE_REQUIRE(preproc >= 0 && preproc <= 2, "enc_conv7x7_toeplitz: unknown preprocessing %d", preproc);
    NBE_REQUIRE((((uintptr_t)wt | (uintptr_t)scratch) & 15) == 0 && ((uintptr_t)y & 127) == 0, "enc_conv7x7_toeplitz: tensors must be 16-byte (y: 128-byte) aligned");
    NBE_REQUIRE(scratch_bytes >= nbe_enc_conv7x7_toeplitz_scratch_bytes(N, H, W), "enc_conv7x7_toeplitz: scratch of %lld bytes required",
                (long long)nbe_enc_conv7x7_toeplitz_scratch_bytes(N, H, W));
    if (N == 0) return NBE_OK;
    const int Wp = W + 16;
    {
        const int64_t total = (int64_t)N * (H + 6) * (Wp / 8);
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)kNumSMs * 16);
        launch_pdl(enc7_pad_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, (__nv_bfloat16*)scratch, N, H, W, Wp, preproc);
        int st = launched("enc7_pad_kernel");
        if (st) return st;
    }
    ToepParams p{};
    p.y = (__nv_bfloat16*)y; p.bias = bias; p.N = N; p.H = H; p.W = W; p.y_cs = y_cs;
    p.tiles_x = W / 128; p.tiles_y = H / 8;
    const int64_t total = (int64_t)N * p.tiles_x * p.tiles_y;
    NBE_REQUIRE(total <= INT32_MAX, "enc_conv7x7_toeplitz: too many tiles");
    p.total = (int)total;
    p.neg_slope = neg_slope;
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    CUtensorMap tp, tw;
    {
        int st = toeplitz_image_map(&tp, scratch, N, H, W);
        if (st) return st;
    }
    {
        cuuint64_t dims[2] = {16, 7 * 512};
        cuuint64_t strides[1] = {32};
        cuuint32_t box[2] = {16, 256};
        int st = make_tmap(&tw, wt, 2, dims, strides, box, "toeplitz weights", 1, 32);
        if (st) return st;
    }
    CUtensorMap ty;
    {
        // interior of y [N, H+2, W+2, 64]: (c, x, y, n); a box = every 8th pixel of 128 columns x 8 rows
        cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {128, (cuuint64_t)(W + 2) * 128, (cuuint64_t)(H + 2) * (W + 2) * 128};
        cuuint32_t box[4] = {64, 128, 8, 1};
        const cuuint32_t estr[5] = {1, 8, 1, 1, 1};
        int st = make_tmap_strided(&ty, (const __nv_bfloat16*)y + ((size_t)(W + 2) + 1) * 64, 4, dims, strides, box, estr, "toeplitz output");
        if (st) return st;
    }
    const size_t smem = 1024 + TP_B_BYTES + TP_ASTAGES * TP_A_BYTES + 4 * TP_STAGE_BYTES + 64 * 4 + 128;
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    std::call_once(once, [&] { err = cudaFuncSetAttribute(enc7_toeplitz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
    if (err != cudaSuccess) return fail(NBE_ECUDA, "enc_conv7x7_toeplitz: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    const int grid = (int)std::min<int64_t>(kNumSMs, total);
    launch_pdl(enc7_toeplitz_kernel, dim3(grid), dim3(TP_THREADS), smem, (cudaStream_t)stream, tp, tw, ty, p);
    return launched("enc7_toeplitz_kernel");
}
