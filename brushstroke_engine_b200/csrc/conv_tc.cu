// modulated_conv2d as an implicit GEMM on 5th-gen tensor cores (tcgen05), BF16 operands, FP32 accumulate.
//
//   D[m, o] = sum_{tap, c} A_tap[m, c] * Wq[tap][o][c]           m = pixel (n, oy, ox), o = out channel
//
// * Activations are NHWC bf16.  For filter tap (kh, kw) the A tile of a 128-pixel block is exactly one 4-D TMA box
//   {64 channels, bw, bh, bn} of the activation tensor shifted by (kw - pad, kh - pad): no im2col buffer, and TMA's
//   out-of-bounds zero fill implements the convolution's zero padding (and the channel padding 144 -> 192).
//   The box lands in shared memory as 128 rows x 128 bytes with the 128-byte swizzle = the canonical K-major UMMA
//   operand layout, so the smem descriptor is built once per stage and advanced 32 bytes per K=16 step.
// * Weights are pre-laid out [tap][Cout][Cin_pad] bf16 (nbe_prepare_weights_bf16); one 3-D TMA box per k-block.
//   "Each patch carries its own style": modulation is applied on the *input* side by the producer of x (previous
//   epilogue / upsample kernel) and demodulation in this epilogue, so the B operand is shared by the whole batch.
// * Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
//   warps 2..5 = epilogue (tcgen05.ld 32 lanes x 32 columns at a time -> demod, noise, bias, lrelu, gain, clamp,
//   next-layer modulation -> bf16 -> 16-byte stores).  smem ring of STAGES {A 16 KiB, B Cout*128 B} with
//   full/empty mbarriers; tcgen05.commit releases a stage when the MMAs that read it retire.
// * Two CTAs fit per SM (<= 113 KiB smem, 128 TMEM columns each) so one CTA's epilogue overlaps the other's MMAs.
#include "tc_common.cuh"
#include <algorithm>
#include <mutex>
#include <cstdlib>

namespace nbe {

// ------------------------------------------------------------------------------------------------ kernel
struct ConvTcParams {
    __nv_bfloat16* y;
    int N, OH, OW, Cout, y_cs;
    int bw, bh, bn;                 // pixel box of one M tile (bw * bh * bn == 128)
    int tiles_x, tiles_y;           // tiles per image in x / y
    int KK, K, pad_off;             // taps, kernel size, coordinate offset (= -pad for 'same', 0 for 'valid')
    int in_stride;                  // 1 or 2 (input pixel = out pixel * in_stride + tap)
    long long y_row_pitch, y_img_pitch;   // output addressing in pixels (lets the epilogue write into the interior of a padded buffer)
    int k_chunks;                   // Cin_pad / 64
    const float* dcoef; const float* noise; long long noise_sn; float noise_gain;
    const float* bias; float alpha, gain, clamp; const float* next_scale;
    uint32_t idesc; uint32_t tmem_cols;
};

constexpr int TC_STAGES = 3;
constexpr int TC_A_BYTES = 128 * 128;                               // 128 pixels x 64 bf16
constexpr int TC_THREADS = 192;

// TPC = position tiles per CTA: with 2, every weight tile that streams in feeds two accumulators -- the wide small-map layers
// (encoder 256 -> 256 @16^2: 32 KiB of weights per 16 KiB of activations and K block) are bound by that L2 -> SM traffic
template <int TPC>
__global__ void __launch_bounds__(TC_THREADS)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const ConvTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    pdl_trigger();
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = p.Cout * 128;
    uint8_t* smem_a = smem;                                         // [TC_STAGES][TPC][TC_A_BYTES]
    uint8_t* smem_b = smem + TC_STAGES * TPC * TC_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + TC_STAGES * b_bytes);
    // bars[0..S) full, bars[S..2S) empty, bars[2S] tmem_full ; then the TMEM base address
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);

    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;

    // tile -> (n0, y0, x0); a tile past the end (odd tile count, TPC = 2) loads zeros (n0 >= N) and stores nothing
    int n0s[TPC], y0s[TPC], x0s[TPC];
#pragma unroll
    for (int i = 0; i < TPC; ++i) {
        int tile = blockIdx.x * TPC + i;
        const int txi = tile % p.tiles_x; tile /= p.tiles_x;
        const int tyi = tile % p.tiles_y; tile /= p.tiles_y;
        n0s[i] = tile * p.bn; y0s[i] = tyi * p.bh; x0s[i] = txi * p.bw;
    }
    const int num_kb = p.KK * p.k_chunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[TC_STAGES + s]), 1); }
        mbar_init(smem_u32(&bars[2 * TC_STAGES]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    pdl_wait();

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int tap = kb / p.k_chunks, cc = kb - tap * p.k_chunks;
                const int kh = tap / p.K, kw = tap - kh * p.K;
                mbar_wait(smem_u32(&bars[TC_STAGES + stage]), phase ^ 1);
                const uint32_t full = smem_u32(&bars[stage]);
                mbar_expect_tx(full, TPC * TC_A_BYTES + b_bytes);
#pragma unroll
                for (int i = 0; i < TPC; ++i)
                    tma_load_4d(smem_u32(smem_a + (stage * TPC + i) * TC_A_BYTES), &tmap_a, full, cc * 64, x0s[i] * p.in_stride + kw + p.pad_off,
                                y0s[i] * p.in_stride + kh + p.pad_off, n0s[i]);
                tma_load_3d(smem_u32(smem_b + stage * b_bytes), &tmap_b, full, cc * 64, 0, tap);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer (whole warp, elected lane issues: see tc_common.cuh) ==============================
        {
            const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a)), b_lo0 = umma_desc_lo(smem_u32(smem_b));
            const uint32_t b_step = (uint32_t)b_bytes >> 4, idesc = p.idesc;
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_fast(smem_u32(&bars[stage]), phase);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t b_lo = b_lo0 + (uint32_t)stage * b_step;
#pragma unroll
                    for (int i = 0; i < TPC; ++i) {
                        const uint32_t a_lo = a_lo0 + (uint32_t)(stage * TPC + i) * (TC_A_BYTES >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)                 // 4 x (K = 16) per 64-channel block: +32 B inside the swizzle atom
                            umma_bf16_lo(tmem_base + (uint32_t)(i * p.Cout), a_lo + (uint32_t)(k * 2), b_lo + (uint32_t)(k * 2), idesc, (kb | k) != 0);
                    }
                    umma_commit(smem_u32(&bars[TC_STAGES + stage]));    // frees the smem stage when these MMAs retire
                }
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(smem_u32(&bars[2 * TC_STAGES]));            // accumulator complete
        }
    } else {
        // ============================== epilogue (warps 2..5) ==============================
        const int q = warp & 3;                                     // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;                                // accumulator row = pixel within the tile
        const int wi = m % p.bw;
        const int hi = (m / p.bw) % p.bh;
        const int ni = m / (p.bw * p.bh);
        mbar_wait(smem_u32(&bars[2 * TC_STAGES]), 0);
        tcgen05_fence_after();
#pragma unroll
      for (int ti = 0; ti < TPC; ++ti) {
        const int n = n0s[ti] + ni, oy = y0s[ti] + hi, ox = x0s[ti] + wi;
        const bool valid = n < p.N;
        float nz = 0.f;
        if (valid && p.noise) nz = p.noise[(long long)n * p.noise_sn + (long long)oy * p.OW + ox] * p.noise_gain;
        const float pos_gain = p.gain, neg_gain = p.gain * p.alpha;
        for (int c0 = 0; c0 < p.Cout; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ti * p.Cout + c0), v);
            if (valid) {
                __nv_bfloat16* yp = p.y + ((long long)n * p.y_img_pitch + (long long)oy * p.y_row_pitch + ox) * p.y_cs + c0;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (c0 + g * 8 >= p.Cout) break;                 // Cout is a multiple of 16, chunks are 32 wide
                    int4 out;
                    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float r[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int o = c0 + g * 8 + e * 2 + h;
                            float a = __uint_as_float(v[g * 8 + e * 2 + h]);
                            if (p.dcoef) a *= __ldg(p.dcoef + (long long)n * p.Cout + o);
                            a += nz;
                            if (p.bias) a += __ldg(p.bias + o);
                            a *= (a > 0.f) ? pos_gain : neg_gain;
                            if (p.clamp >= 0.f) a = fminf(fmaxf(a, -p.clamp), p.clamp);
                            if (p.next_scale) a *= __ldg(p.next_scale + (long long)n * p.Cout + o);
                            r[h] = a;
                        }
                        o2[e] = __floats2bfloat162_rn(r[0], r[1]);
                    }
                    *reinterpret_cast<int4*>(yp + g * 8) = out;
                }
            }
        }
      }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Row-resident persistent variant for the 128-wide layers (OW % 128 == 0, Cin_pad == 128, Cout == 128, 3x3, stride 1):
// the per-tap kernel above is bound by L2->SM bandwidth (every 128x128x64 MMA block streams 16 KiB of A and 16 KiB of
// B: ~128 B/clk/SM wanted, ~40 B/clk/SM available chip-wide), so this kernel cuts the traffic per FLOP ~3x:
//  * one CTA per SM, persistent; a work item is TWO output rows (2 x 128 pixels = two TMEM accumulators) of one image;
//  * for each 64-channel chunk the FOUR input rows those outputs touch are loaded ONCE (130 pixels each, halo included)
//    and kept in shared memory; the three horizontal taps are the same smem rows read through UMMA descriptors whose
//    start address is shifted by kw pixels (kw * 128 B -- the 128-byte swizzle is a function of the absolute smem
//    address, so a row-shifted start reads exactly what TMA wrote);
//  * every weight tile (tap, chunk) is streamed once per work item through an 8-deep ring and feeds 8 MMAs
//    (2 output rows x 4 K-steps) instead of 4;
//  * the kernel runs on CTA PAIRS (cta_group::2, M = 256): each CTA of the pair owns its own work item and accumulators
//    but holds only HALF of every weight tile (64 of the 128 output channels), which halves the weight traffic per SM;
//  * accumulators are double-buffered in TMEM (2 x 2 x 128 columns = all 512), so the epilogue of item i overlaps
//    the MMAs of item i+1; per-image epilogue vectors (demod, bias, next-layer styles) are staged in smem.
constexpr int R_THREADS = 320;                                      // TMA warp, MMA warp, 8 epilogue warps (4 per output row of the item)
constexpr int R_ABUF = 17 * 1024;                                  // 130 px x 128 B = 16640 B, padded to a 1 KiB multiple
constexpr int R_AROW_BYTES = 130 * 128;
constexpr int R_BSTAGES = 8;
constexpr int R_BBYTES = 64 * 128;                                  // this CTA's half (64 output channels) of a weight tile

struct ConvRowParams {
    __nv_bfloat16* y;
    int N, OH, OW, y_cs;
    int pad_off;                                                    // -1: zero padding 1 ('same'); 0: input already has the halo ('valid')
    long long y_row_pitch, y_img_pitch;
    int items_per_row, items_per_img, total_items;                  // item = (n, row pair, 128-px x segment)
    const float* dcoef; const float* noise; long long noise_sn; float noise_gain;
    const float* bias; float alpha, gain, clamp; const float* next_scale;
    uint32_t idesc;
    // optional fused ToRGB + triad colour mix (ToRGBColorTriadLayer, networks.py:451-485) on the activated output
    const float* rgb_w; const float* rgb_styles; const float* rgb_bias; const float* rgb_colors; float rgb_clamp;
    float* img; float* uvs; int write_y;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(R_THREADS, 1)
conv_tc_row128_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const ConvRowParams p) {
    extern __shared__ uint8_t smem_raw[];
    pdl_trigger();
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                         // [2 chunks][4 rows][R_ABUF]
    uint8_t* smem_b = smem + 8 * R_ABUF;                            // [R_BSTAGES][R_BBYTES]
    float* s_vec = reinterpret_cast<float*>(smem_b + R_BSTAGES * R_BBYTES);   // [6][128]: dcoef, bias, next_scale, 3 x modulated ToRGB weights of the current image
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_vec + 6 * 128);
    uint64_t* a_full = bars;            // [2]   (the leader's is the live one: TMA of both CTAs completes on it)
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* b_full = bars + 4;        // [R_BSTAGES]
    uint64_t* b_empty = bars + 4 + R_BSTAGES;
    uint64_t* acc_full = bars + 4 + 2 * R_BSTAGES;      // [2]
    uint64_t* acc_empty = acc_full + 2;                 // [2]   (the leader's: both CTAs' epilogue warps arrive on it)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int rank = (int)uniform_u32(cluster_ctarank());
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), 1);
                                      mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 16); }
        for (int i = 0; i < R_BSTAGES; ++i) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();                                             // barriers of BOTH CTAs are initialised before any remote arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    pdl_wait();
    // pair jp of the cluster: CTA r takes item 2 jp + r (the last pair may hold a dummy item: loaded, computed, never stored)
    const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int pairs = (p.total_items + 1) >> 1;

    if (warp == 0) {
        // ============================== TMA producer (both CTAs) ==============================
        if (lane == 0) {
            uint32_t a_phase = 0;
            int bs = 0; uint32_t b_phase = 0;
            for (int jp = cid; jp < pairs; jp += n_clusters) {
                const int item = min(2 * jp + rank, p.total_items - 1);
                const int n = item / p.items_per_img;
                const int rem = item - n * p.items_per_img;
                const int yp = rem / p.items_per_row, xs = rem - yp * p.items_per_row;
                const int y0 = yp * 2, x0 = xs * 128;
                for (int c = 0; c < 2; ++c) {
                    mbar_wait_fast(smem_u32(&a_empty[c]), a_phase ^ 1);
                    const uint32_t full = smem_u32(&a_full[c]);
                    if (leader) mbar_expect_tx(full, 2 * 4 * R_AROW_BYTES);
                    for (int j = 0; j < 4; ++j)
                        tma_load_4d_2sm(smem_u32(smem_a + (c * 4 + j) * R_ABUF), &tmap_a, full, c * 64, x0 + p.pad_off, y0 + j + p.pad_off, n);
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait_fast(smem_u32(&b_empty[bs]), b_phase ^ 1);
                        const uint32_t bf = smem_u32(&b_full[bs]);
                        if (leader) mbar_expect_tx(bf, 2 * R_BBYTES);
                        tma_load_3d_2sm(smem_u32(smem_b + bs * R_BBYTES), &tmap_b, bf, c * 64, rank * 64, tap);
                        if (++bs == R_BSTAGES) { bs = 0; b_phase ^= 1; }
                    }
                }
                a_phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer (leader CTA; whole warp, elected lane issues: see tc_common.cuh) ==============================
        if (leader) {
            const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a)), b_lo0 = umma_desc_lo(smem_u32(smem_b));
            const uint32_t idesc = p.idesc;
            uint32_t a_par = 0, acc_par0 = 0, acc_par1 = 0;
            int bs = 0; uint32_t b_phase = 0;
            int it = 0;
            for (int jp = cid; jp < pairs; jp += n_clusters, ++it) {
                const int ab = it & 1;
                // the epilogues of both CTAs have drained this accumulator pair
                if (ab) { mbar_wait_fast(smem_u32(&acc_empty[1]), acc_par1 ^ 1); acc_par1 ^= 1; }
                else    { mbar_wait_fast(smem_u32(&acc_empty[0]), acc_par0 ^ 1); acc_par0 ^= 1; }
                tcgen05_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(ab * 256);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    mbar_wait_fast(smem_u32(&a_full[c]), a_par);
                    tcgen05_fence_after();
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int kh = tap / 3, kw = tap - kh * 3;
                        mbar_wait_fast(smem_u32(&b_full[bs]), b_phase);
                        tcgen05_fence_after();
                        if (elect_one()) {
                            const uint32_t b_lo = b_lo0 + (uint32_t)bs * (R_BBYTES >> 4);
#pragma unroll
                            for (int r = 0; r < 2; ++r) {
                                // output row r takes input row r + kh; the horizontal tap is a kw-pixel (kw * 128 B) shift of the start address
                                const uint32_t a_lo = a_lo0 + (uint32_t)(((c * 4 + r + kh) * R_ABUF + kw * 128) >> 4);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_lo_2sm(d0 + (uint32_t)(r * 128), a_lo + (uint32_t)(k * 2), b_lo + (uint32_t)(k * 2), idesc, (c | tap | k) != 0);
                            }
                            umma_commit_2sm(smem_u32(&b_empty[bs]));
                        }
                        if (++bs == R_BSTAGES) { bs = 0; b_phase ^= 1; }
                    }
                    if (elect_one()) umma_commit_2sm(smem_u32(&a_empty[c]));         // the four input rows of this chunk may be overwritten
                }
                a_par ^= 1;
                if (elect_one()) umma_commit_2sm(smem_u32(&acc_full[ab]));
            }
        }
    } else {
        // ============================== epilogue (warps 2..9) ==============================
        // One thread per output pixel and row: warps 2..5 take output row 0 of the item, warps 6..9 row 1 (TMEM lane quarter
        // q = warp % 4, hardware rule).  With one epilogue warp per SM sub-partition (4 warps walking both rows in turn) the
        // epilogue itself -- ~14 dependent instructions per channel value behind LDS / TMEM-load latencies nothing else covered --
        // took longer per item than the MMAs and set the pace of the kernel; two warps per sub-partition hide each other's latencies.
        const int q = warp & 3;
        const int er = (warp - 2) >> 2;                             // output row of the item this warp owns
        const int m = q * 32 + lane;                                // pixel x within the 128-px segment
        const int et = threadIdx.x - 64;                            // 0..255
        uint32_t acc_phase[2] = {0, 0};
        int it = 0, cur_n = -1;
        const float pos_gain = p.gain, neg_gain = p.gain * p.alpha;
        const uint32_t s_vec_u32 = smem_u32(s_vec);
        for (int jp = cid; jp < pairs; jp += n_clusters, ++it) {
            const int ab = it & 1;
            const bool dummy = 2 * jp + rank >= p.total_items;
            const int item = min(2 * jp + rank, p.total_items - 1);
            const int n = item / p.items_per_img;
            const int rem = item - n * p.items_per_img;
            const int yp = rem / p.items_per_row, xs = rem - yp * p.items_per_row;
            const int y0 = yp * 2, ox = xs * 128 + m;
            if (n != cur_n) {                                       // per-image epilogue vectors -> smem (epilogue warps only)
                asm volatile("bar.sync 1, 256;" ::: "memory");      // everyone is done reading the previous image's vectors
                if (et < 128) {
                    s_vec[et] = p.dcoef ? p.dcoef[(long long)n * 128 + et] : 1.f;
                    s_vec[128 + et] = p.bias ? p.bias[et] : 0.f;
                    s_vec[256 + et] = p.next_scale ? p.next_scale[(long long)n * 128 + et] : 1.f;
                } else if (p.rgb_w) {
                    const int ec = et - 128;
                    const float st = p.rgb_styles[(long long)n * 128 + ec];
                    s_vec[384 + ec] = p.rgb_w[ec] * st; s_vec[512 + ec] = p.rgb_w[128 + ec] * st; s_vec[640 + ec] = p.rgb_w[256 + ec] * st;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                cur_n = n;
            }
            mbar_wait(smem_u32(&acc_full[ab]), acc_phase[ab]);
            acc_phase[ab] ^= 1;
            tcgen05_fence_after();
            if (!dummy) {
                const int r = er;
                const int oy = y0 + r;
                float nz = 0.f;
                if (p.noise) nz = p.noise[(long long)n * p.noise_sn + (long long)oy * p.OW + ox] * p.noise_gain;
                __nv_bfloat16* yrow = p.y + ((long long)n * p.y_img_pitch + (long long)oy * p.y_row_pitch + ox) * p.y_cs;
                float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 256 + r * 128 + c0), v);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        int4 out;
                        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&out);
                        // the per-channel vectors are the same for all 32 lanes (broadcast loads): fetched 16 bytes at a time, they cost
                        // 1.5 shared-memory wavefronts per channel instead of 6 -- the data pipe they ride on is the one the tensor
                        // core reads its operands through (ncu: LSU 41 % + UMMA 45 % of the pipe with scalar loads)
                        float dc[8], bs[8], ns[8], w0[8], w1[8], w2[8];
                        const int ob = c0 + g * 8;
#pragma unroll
                        for (int h4 = 0; h4 < 2; ++h4) {
                            lds_f4(s_vec_u32 + (uint32_t)((ob + 4 * h4) * 4), dc + 4 * h4);
                            lds_f4(s_vec_u32 + (uint32_t)((128 + ob + 4 * h4) * 4), bs + 4 * h4);
                            if (p.write_y) lds_f4(s_vec_u32 + (uint32_t)((256 + ob + 4 * h4) * 4), ns + 4 * h4);   // (the fused ToRGB launch does not store y)
                            if (p.rgb_w) {
                                lds_f4(s_vec_u32 + (uint32_t)((384 + ob + 4 * h4) * 4), w0 + 4 * h4);
                                lds_f4(s_vec_u32 + (uint32_t)((512 + ob + 4 * h4) * 4), w1 + 4 * h4);
                                lds_f4(s_vec_u32 + (uint32_t)((640 + ob + 4 * h4) * 4), w2 + 4 * h4);
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float rr[2];
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int k = e * 2 + h;
                                float a = __uint_as_float(v[g * 8 + k]) * dc[k] + nz + bs[k];
                                a *= (a > 0.f) ? pos_gain : neg_gain;
                                if (p.clamp >= 0.f) a = fminf(fmaxf(a, -p.clamp), p.clamp);
                                if (p.rgb_w) {
                                    t0 = fmaf(a, w0[k], t0); t1 = fmaf(a, w1[k], t1); t2 = fmaf(a, w2[k], t2);
                                }
                                rr[h] = p.write_y ? a * ns[k] : 0.f;
                            }
                            o2[e] = __floats2bfloat162_rn(rr[0], rr[1]);
                        }
                        if (p.write_y) *reinterpret_cast<int4*>(yrow + c0 + g * 8) = out;
                    }
                }
                if (p.rgb_w) {
                    t0 += p.rgb_bias[0]; t1 += p.rgb_bias[1]; t2 += p.rgb_bias[2];
                    if (p.rgb_clamp >= 0.f) {
                        t0 = fminf(fmaxf(t0, -p.rgb_clamp), p.rgb_clamp); t1 = fminf(fmaxf(t1, -p.rgb_clamp), p.rgb_clamp);
                        t2 = fminf(fmaxf(t2, -p.rgb_clamp), p.rgb_clamp);
                    }
                    const float mx = fmaxf(t0, fmaxf(t1, t2));
                    const float e0 = expf(t0 - mx), e1 = expf(t1 - mx), e2 = expf(t2 - mx);
                    const float inv = 1.f / (e0 + e1 + e2);
                    const float u0 = e0 * inv, u1 = e1 * inv, u2 = e2 * inv;
                    const long long plane = (long long)p.OH * p.OW, pix = (long long)oy * p.OW + ox;
                    const float* col = p.rgb_colors + (long long)n * 9;
                    if (p.uvs) {
                        float* up = p.uvs + (long long)n * 3 * plane + pix;
                        up[0] = u0; up[plane] = u1; up[2 * plane] = u2;
                    }
                    if (p.img) {
                        float* ip = p.img + (long long)n * 3 * plane + pix;
                        ip[0] = u0 * col[0] + u1 * col[1] + u2 * col[2];
                        ip[plane] = u0 * col[3] + u1 * col[4] + u2 * col[5];
                        ip[2 * plane] = u0 * col[6] + u1 * col[7] + u2 * col[8];
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(smem_u32(&acc_empty[ab]));
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();                                             // the peer's smem / TMEM stay alive until the leader's last MMA is done
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Encoder first layer on the tensor pipe: conv7x7(1 -> Cout <= 64... here N = Cout, multiple of 16) with reflect
// padding 3 as ONE K = 64 GEMM per 128-pixel tile (49 taps + 15 zero columns).  The im2col A tile does not exist in
// memory: each thread writes the 49 taps of ITS pixel (bf16) straight into the 128-byte-swizzled K-major smem layout
// the UMMA descriptor expects (16-byte chunk c of row r lives at chunk c ^ (r & 7)), a fence.proxy.async makes the
// generic-proxy writes visible to the tensor core, one thread issues the four K = 16 MMAs, and the usual TMEM
// epilogue applies the folded-BatchNorm bias + LeakyReLU and writes the interior of the reflect-padded NHWC buffer.
constexpr int E7T_TW = 16, E7T_TH = 8;                             // 16 x 8 = 128 output pixels per tile

__global__ void __launch_bounds__(128, 8)
enc_conv7x7_tc_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ wq, const float* __restrict__ bias,
                      __nv_bfloat16* __restrict__ y, int N, int H, int W, int Cout, int y_cs, float neg_slope, int preproc,
                      int tiles_x, int tiles_y, uint32_t idesc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                         // 128 rows x 128 B
    uint8_t* smem_b = smem + 128 * 128;                             // Cout rows x 128 B
    float* s_in = reinterpret_cast<float*>(smem_b + 64 * 128);      // [E7T_TH + 6][E7T_TW + 6]
    float* s_bias = s_in + (E7T_TH + 6) * (E7T_TW + 6);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_bias + 64);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // weights -> swizzled smem B tile (row o = out channel, 64 K values = 8 chunks of 16 B), once per CTA
    for (int i = threadIdx.x; i < Cout * 8; i += 128) {
        const int o = i >> 3, c = i & 7;
        const int4 v = *reinterpret_cast<const int4*>(wq + o * 64 + c * 8);
        *reinterpret_cast<int4*>(smem_b + o * 128 + ((c ^ (o & 7)) << 4)) = v;
    }
    if (threadIdx.x < Cout) s_bias[threadIdx.x] = bias[threadIdx.x];
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int m = threadIdx.x;                                      // pixel / accumulator row of this thread
    const int ly = m / E7T_TW, lx = m % E7T_TW;
    const int total = tiles_x * tiles_y * N;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int t = tile;
        const int tx = t % tiles_x; t /= tiles_x;
        const int ty = t % tiles_y; t /= tiles_y;
        const int n = t;
        const int ty0 = ty * E7T_TH, tx0 = tx * E7T_TW;
        for (int i = threadIdx.x; i < (E7T_TH + 6) * (E7T_TW + 6); i += 128) {
            const int r = i / (E7T_TW + 6), c = i - r * (E7T_TW + 6);
            int iy = ty0 + r - 3, ix = tx0 + c - 3;
            iy = iy < 0 ? -iy : iy; iy = iy >= H ? 2 * H - 2 - iy : iy;     // reflect
            ix = ix < 0 ? -ix : ix; ix = ix >= W ? 2 * W - 2 - ix : ix;
            iy = min(max(iy, 0), H - 1); ix = min(max(ix, 0), W - 1);       // tiles hanging over the border: any in-range value
            float v = x[((long long)n * H + iy) * W + ix];
            if (preproc == 1) v = 1.f - v;
            else if (preproc == 2) v = (1.f - v) * 2.f - 1.f;
            s_in[i] = v;
        }
        __syncthreads();
        // this thread's im2col row: taps 0..48 (kh * 7 + kw), zero padded to 64, as 8 chunks of 8 bf16
        {
            float taps[56];
#pragma unroll
            for (int kh = 0; kh < 7; ++kh)
#pragma unroll
                for (int kw = 0; kw < 7; ++kw) taps[kh * 7 + kw] = s_in[(ly + kh) * (E7T_TW + 6) + lx + kw];
#pragma unroll
            for (int k = 49; k < 56; ++k) taps[k] = 0.f;
            uint8_t* row = smem_a + m * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                int4 v = make_int4(0, 0, 0, 0);
                if (c < 7) {
                    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
                    for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(taps[c * 8 + 2 * e], taps[c * 8 + 2 * e + 1]);
                }
                *reinterpret_cast<int4*>(row + ((c ^ (m & 7)) << 4)) = v;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the async (tensor) proxy
        tcgen05_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tcgen05_fence_after();
            const uint64_t a_desc = umma_smem_desc(smem_u32(smem_a));
            const uint64_t b_desc = umma_smem_desc(smem_u32(smem_b));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, k != 0);
            umma_commit(smem_u32(bar));
        }
        mbar_wait(smem_u32(bar), phase);
        phase ^= 1;
        tcgen05_fence_after();
        const int oy = ty0 + ly, ox = tx0 + lx;
        const bool ok = oy < H && ox < W;
        // A pixel's channels are contiguous bytes of the output, but a pixel is owned by one lane: stored straight from the
        // owning lanes every instruction touches 32 lines 16 bytes at a time.  The warp's 32 pixels x 64 channels go through
        // its own quarter of the (now idle) A tile instead and leave as whole 128-byte lines.
        const long long pix = ok ? ((long long)n * (H + 2) + oy + 1) * (W + 2) + ox + 1 : -1;
        uint8_t* stg = smem_a + warp * 32 * 128;
        for (int c0 = 0; c0 < Cout; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (c0 + g * 8 >= Cout) break;
                int4 out;
                __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float a = __uint_as_float(v[g * 8 + 2 * e]) + s_bias[c0 + g * 8 + 2 * e];
                    float b = __uint_as_float(v[g * 8 + 2 * e + 1]) + s_bias[c0 + g * 8 + 2 * e + 1];
                    a = a > 0.f ? a : a * neg_slope;
                    b = b > 0.f ? b : b * neg_slope;
                    o2[e] = __floats2bfloat162_rn(a, b);
                }
                const int j = (c0 >> 3) + g;                         // 16-byte chunk of the pixel's 128-byte row
                *reinterpret_cast<int4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = out;
            }
        }
        __syncwarp();
        {
            const int chunks = Cout >> 3;                           // 16-byte chunks per pixel (<= 8)
            const int rd_ch = lane & 7, rd_row0 = lane >> 3;
#pragma unroll
            for (int r8 = 0; r8 < 8; ++r8) {
                const int row = r8 * 4 + rd_row0;
                const long long rp = __shfl_sync(0xffffffffu, pix, row);
                const int4 val = *reinterpret_cast<const int4*>(stg + row * 128 + ((rd_ch ^ (row & 7)) << 4));
                if (rp >= 0 && rd_ch < chunks) *reinterpret_cast<int4*>(y + rp * y_cs + rd_ch * 8) = val;
            }
        }
        tcgen05_fence_before();
        __syncthreads();                                            // A tile / s_in / TMEM accumulator are reused by the next tile
        tcgen05_fence_after();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(64u) : "memory");
}

// [Cout][Cin][K][K] f32 -> [KK][Cout][Cin_pad] bf16 (flip = true convolution)
__global__ void prepare_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wq, int Cout, int Cin, int Cin_pad, int KK, int flip) {
    const int64_t total = (int64_t)KK * Cout * Cin_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cin_pad);
        const int o = (int)((i / Cin_pad) % Cout);
        const int t = (int)(i / ((int64_t)Cin_pad * Cout));
        float v = 0.f;
        if (c < Cin) v = w[((int64_t)o * Cin + c) * KK + (flip ? KK - 1 - t : t)];
        wq[i] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

int make_tmap(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box, const char* what, int spatial_stride, int swizzle128, int f32) {
    const cuuint32_t estr[5] = {1, (cuuint32_t)spatial_stride, (cuuint32_t)spatial_stride, 1, 1};
    return make_tmap_strided(map, base, rank, dims, strides_bytes, box, estr, what, swizzle128, f32);
}

int make_tmap_strided(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, const cuuint32_t* estr, const char* what, int swizzle128, int f32) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(NBE_ECUDA, "conv_tc: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NBE_ECUDA, "conv_tc: cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return NBE_OK;
}

}  // namespace nbe

using namespace nbe;

extern "C" int nbe_prepare_weights_bf16(const float* w, void* wq, int Cout, int Cin, int K, int flip, nbe_stream_t stream) {
    NBE_REQUIRE(w && wq && Cout >= 1 && Cin >= 1 && K >= 1, "prepare_weights: bad arguments");
    const int Cin_pad = (Cin + 63) / 64 * 64;
    const int64_t total = (int64_t)K * K * Cout * Cin_pad;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    prepare_weights_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wq, Cout, Cin, Cin_pad, K * K, flip);
    return launched("prepare_weights_kernel");
}

extern "C" int nbe_conv_tc_bf16(const void* x, const void* wq, void* y,
                                int N, int OH, int OW, int Cin, int x_cs, int Cout, int y_cs, int K, int valid,
                                const float* dcoef, const float* noise, int64_t noise_sn, float noise_gain,
                                const float* bias, float alpha, float gain, float clamp, const float* next_scale,
                                nbe_stream_t stream) {
    return nbe_conv_tc_bf16_ex(x, wq, y, N, OH, OW, Cin, x_cs, Cout, y_cs, K, valid, 1, 0, 0, OW, (int64_t)OH * OW,
                               dcoef, noise, noise_sn, noise_gain, bias, alpha, gain, clamp, next_scale, stream);
}

struct TorgbArgs { const float* w; const float* styles; const float* bias; const float* colors; float clamp; float* img; float* uvs; int write_y; };

static int conv_tc_impl(const void* x, const void* wq, void* y,
                        int N, int OH, int OW, int Cin, int x_cs, int Cout, int y_cs, int K, int valid,
                        int in_stride, int in_h, int in_w, int64_t y_row_pitch, int64_t y_img_pitch,
                        const float* dcoef, const float* noise, int64_t noise_sn, float noise_gain,
                        const float* bias, float alpha, float gain, float clamp, const float* next_scale,
                        const TorgbArgs* rgb, nbe_stream_t stream) {
    NBE_REQUIRE(in_stride == 1 || in_stride == 2, "conv_tc: input stride must be 1 or 2");
    NBE_REQUIRE(in_stride == 1 || valid, "conv_tc: strided convolution needs a pre-padded input (valid = 1)");
    NBE_REQUIRE(y_row_pitch >= OW && y_img_pitch >= y_row_pitch * OH, "conv_tc: bad output pitches");
    NBE_REQUIRE(x && wq && y, "conv_tc: null tensor");
    NBE_REQUIRE(N >= 0 && OH >= 1 && OW >= 1 && Cin >= 1, "conv_tc: bad shape");
    NBE_REQUIRE(K == 1 || K == 3, "conv_tc: kernel size %d not supported (1 or 3)", K);
    NBE_REQUIRE(Cout >= 16 && Cout <= 256 && Cout % 16 == 0, "conv_tc: Cout must be a multiple of 16 in [16, 256]");
    NBE_REQUIRE(x_cs % 8 == 0 && x_cs >= Cin && y_cs % 8 == 0 && y_cs >= Cout, "conv_tc: channel strides must be multiples of 8");
    NBE_REQUIRE((((uintptr_t)x | (uintptr_t)wq | (uintptr_t)y) & 15) == 0, "conv_tc: tensors must be 16-byte aligned");
    NBE_REQUIRE((OW & (OW - 1)) == 0 && (OH & (OH - 1)) == 0, "conv_tc: output size must be a power of two");
    if (N == 0) return NBE_OK;

    ConvTcParams p;
    p.y = (__nv_bfloat16*)y; p.N = N; p.OH = OH; p.OW = OW; p.Cout = Cout; p.y_cs = y_cs;
    p.bw = OW < 128 ? OW : 128;
    p.bh = (128 / p.bw) < OH ? (128 / p.bw) : OH;
    p.bn = 128 / (p.bw * p.bh);
    p.tiles_x = OW / p.bw; p.tiles_y = OH / p.bh;
    p.K = K; p.KK = K * K;
    const int pad = K / 2;
    p.pad_off = valid ? 0 : -pad;
    const int need_h = valid ? (OH - 1) * in_stride + K : OH, need_w = valid ? (OW - 1) * in_stride + K : OW;
    const int IH = in_h > 0 ? in_h : need_h, IW = in_w > 0 ? in_w : need_w;
    NBE_REQUIRE(IH >= need_h && IW >= need_w, "conv_tc: input %dx%d too small for output %dx%d", IH, IW, OH, OW);
    p.in_stride = in_stride; p.y_row_pitch = y_row_pitch; p.y_img_pitch = y_img_pitch;
    const int Cin_pad = (Cin + 63) / 64 * 64;
    p.k_chunks = Cin_pad / 64;
    p.dcoef = dcoef; p.noise = noise; p.noise_sn = noise_sn; p.noise_gain = noise_gain;
    p.bias = bias; p.alpha = alpha; p.gain = gain; p.clamp = clamp; p.next_scale = next_scale;
    // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), K-major A and B, N >> 3 at 17, M >> 4 at 24
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.tmem_cols = Cout <= 32 ? 32 : Cout <= 64 ? 64 : Cout <= 128 ? 128 : 256;

    // ---- 128-wide layers: row-resident persistent kernel (see conv_tc_row128_kernel) ----
    static const bool force_v1 = getenv("NBE_CONV_TC_V1") != nullptr;
    if (!force_v1 && K == 3 && in_stride == 1 && Cout == 128 && Cin_pad == 128 && OW % 128 == 0 && OH % 2 == 0) {
        ConvRowParams r;
        r.y = (__nv_bfloat16*)y; r.N = N; r.OH = OH; r.OW = OW; r.y_cs = y_cs; r.pad_off = p.pad_off;
        r.y_row_pitch = y_row_pitch; r.y_img_pitch = y_img_pitch;
        r.items_per_row = OW / 128; r.items_per_img = (OH / 2) * r.items_per_row;
        const int64_t total = (int64_t)N * r.items_per_img;
        NBE_REQUIRE(total <= INT32_MAX, "conv_tc: too many work items");
        r.total_items = (int)total;
        r.dcoef = dcoef; r.noise = noise; r.noise_sn = noise_sn; r.noise_gain = noise_gain;
        r.bias = bias; r.alpha = alpha; r.gain = gain; r.clamp = clamp; r.next_scale = next_scale;
        // kind::f16, f32 accumulate, bf16 x bf16, K-major, N = 128, M = 256 (CTA pair)
        r.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        r.rgb_w = nullptr; r.rgb_styles = nullptr; r.rgb_bias = nullptr; r.rgb_colors = nullptr; r.rgb_clamp = -1.f;
        r.img = nullptr; r.uvs = nullptr; r.write_y = 1;
        if (rgb) {
            r.rgb_w = rgb->w; r.rgb_styles = rgb->styles; r.rgb_bias = rgb->bias; r.rgb_colors = rgb->colors; r.rgb_clamp = rgb->clamp;
            r.img = rgb->img; r.uvs = rgb->uvs; r.write_y = rgb->write_y;
        }
        CUtensorMap ta, tb;
        {
            cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)N};
            cuuint64_t strides[3] = {(cuuint64_t)x_cs * 2, (cuuint64_t)IW * x_cs * 2, (cuuint64_t)IH * IW * x_cs * 2};
            cuuint32_t box[4] = {64, 130, 1, 1};
            int st = make_tmap(&ta, x, 4, dims, strides, box, "activation rows");
            if (st) return st;
        }
        {
            cuuint64_t dims[3] = {(cuuint64_t)Cin_pad, (cuuint64_t)Cout, 9};
            cuuint64_t strides[2] = {(cuuint64_t)Cin_pad * 2, (cuuint64_t)Cout * Cin_pad * 2};
            cuuint32_t box[3] = {64, 64, 1};                        // half of the output channels per CTA of the pair
            int st = make_tmap(&tb, wq, 3, dims, strides, box, "weights");
            if (st) return st;
        }
        const size_t rsmem = 1024 + 8 * R_ABUF + R_BSTAGES * R_BBYTES + 6 * 128 * sizeof(float) + 256;
        static std::once_flag row_once;
        static cudaError_t row_err = cudaSuccess;
        std::call_once(row_once, [] {
            row_err = cudaFuncSetAttribute(conv_tc_row128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        });
        if (row_err != cudaSuccess) return fail(NBE_ECUDA, "conv_tc: cudaFuncSetAttribute(row128): %s", cudaGetErrorString(row_err));
        const int grid = (int)std::min<int64_t>(kNumSMs / 2, (total + 1) / 2) * 2;      // whole CTA pairs
        launch_pdl(conv_tc_row128_kernel, dim3(grid), dim3(R_THREADS), rsmem, (cudaStream_t)stream, ta, tb, r);
        return launched("conv_tc_row128_kernel");
    }

    if (rgb) return fail(NBE_EUNSUPPORTED, "conv_tc: the fused ToRGB epilogue needs a 128-wide layer (OW %% 128 == 0, Cin <= 128, Cout == 128)");

    CUtensorMap tmap_a, tmap_b;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)x_cs * 2, (cuuint64_t)IW * x_cs * 2, (cuuint64_t)IH * IW * x_cs * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(p.bw * in_stride), (cuuint32_t)(p.bh * in_stride), (cuuint32_t)p.bn};
        int st = make_tmap(&tmap_a, x, 4, dims, strides, box, "activations", in_stride);
        if (st) return st;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin_pad, (cuuint64_t)Cout, (cuuint64_t)p.KK};
        cuuint64_t strides[2] = {(cuuint64_t)Cin_pad * 2, (cuuint64_t)Cout * Cin_pad * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)Cout, 1};
        int st = make_tmap(&tmap_b, wq, 3, dims, strides, box, "weights");
        if (st) return st;
    }
    const int64_t tiles = (int64_t)p.tiles_x * p.tiles_y * ((N + p.bn - 1) / p.bn);
    NBE_REQUIRE(tiles <= INT32_MAX, "conv_tc: too many tiles");
    // two tiles per CTA when the weight tile is at least as large as an activation tile and both accumulators fit in TMEM
    static const bool one_tile = getenv("NBE_CONV_TC_ONE_TILE") != nullptr;          // A/B switch
    const int tpc = (!one_tile && Cout >= 128 && 2 * Cout <= 512 && tiles >= 2 * kNumSMs) ? 2 : 1;
    if (tpc == 2) { uint32_t c = 32; while (c < (uint32_t)(2 * Cout)) c <<= 1; p.tmem_cols = c; }
    const size_t smem = 1024 + (size_t)TC_STAGES * (tpc * TC_A_BYTES + Cout * 128) + 128;
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] {
        attr_err = cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess) return fail(NBE_ECUDA, "conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
    if (tpc == 2) launch_pdl(conv_tc_kernel<2>, dim3((int)((tiles + 1) / 2)), dim3(TC_THREADS), smem, (cudaStream_t)stream, tmap_a, tmap_b, p);
    else          launch_pdl(conv_tc_kernel<1>, dim3((int)tiles), dim3(TC_THREADS), smem, (cudaStream_t)stream, tmap_a, tmap_b, p);
    return launched("conv_tc_kernel");
}

extern "C" int nbe_conv_tc_bf16_ex(const void* x, const void* wq, void* y,
                                   int N, int OH, int OW, int Cin, int x_cs, int Cout, int y_cs, int K, int valid,
                                   int in_stride, int in_h, int in_w, int64_t y_row_pitch, int64_t y_img_pitch,
                                   const float* dcoef, const float* noise, int64_t noise_sn, float noise_gain,
                                   const float* bias, float alpha, float gain, float clamp, const float* next_scale,
                                   nbe_stream_t stream) {
    return conv_tc_impl(x, wq, y, N, OH, OW, Cin, x_cs, Cout, y_cs, K, valid, in_stride, in_h, in_w, y_row_pitch, y_img_pitch,
                        dcoef, noise, noise_sn, noise_gain, bias, alpha, gain, clamp, next_scale, nullptr, stream);
}

extern "C" int nbe_conv_tc_bf16_torgb(const void* x, const void* wq, void* y,
                                      int N, int OH, int OW, int Cin, int x_cs, int Cout, int y_cs, int valid,
                                      const float* dcoef, const float* noise, int64_t noise_sn, float noise_gain,
                                      const float* bias, float alpha, float gain, float clamp,
                                      const float* rgb_w, const float* rgb_styles, const float* rgb_bias, const float* rgb_colors,
                                      float rgb_clamp, float* img, float* uvs, int write_y, nbe_stream_t stream) {
    NBE_REQUIRE(rgb_w && rgb_styles && rgb_bias && rgb_colors && (img || uvs), "conv_tc_torgb: null ToRGB tensor");
    NBE_REQUIRE(y || !write_y, "conv_tc_torgb: feature output requested without a buffer");
    TorgbArgs a{rgb_w, rgb_styles, rgb_bias, rgb_colors, rgb_clamp, img, uvs, write_y};
    return conv_tc_impl(x, wq, y ? y : (void*)x, N, OH, OW, Cin, x_cs, Cout, y_cs, 3, valid, 1, 0, 0, OW, (int64_t)OH * OW,
                        dcoef, noise, noise_sn, noise_gain, bias, alpha, gain, clamp, nullptr, &a, stream);
}

extern "C" int nbe_enc_conv7x7_tc_bf16(const float* x, const void* wq, const float* bias, void* y, int N, int H, int W, int Cout,
                                       int y_cs, float neg_slope, int preproc, nbe_stream_t stream) {
    NBE_REQUIRE(x && wq && bias && y && N >= 0 && H >= 4 && W >= 4, "enc_conv7x7_tc: bad arguments");
    NBE_REQUIRE(Cout % 16 == 0 && Cout >= 16 && Cout <= 64 && y_cs % 8 == 0 && y_cs >= Cout, "enc_conv7x7_tc: Cout must be 16..64 in steps of 16");
    NBE_REQUIRE(preproc >= 0 && preproc <= 2, "enc_conv7x7_tc: unknown preprocessing %d", preproc);
    NBE_REQUIRE((((uintptr_t)wq | (uintptr_t)y) & 15) == 0, "enc_conv7x7_tc: tensors must be 16-byte aligned");
    if (N == 0) return NBE_OK;
    const int tiles_x = (W + E7T_TW - 1) / E7T_TW, tiles_y = (H + E7T_TH - 1) / E7T_TH;
    const int64_t total = (int64_t)tiles_x * tiles_y * N;
    NBE_REQUIRE(total <= INT32_MAX, "enc_conv7x7_tc: too many tiles");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const size_t smem = 1024 + 128 * 128 + 64 * 128 + (E7T_TH + 6) * (E7T_TW + 6) * 4 + 64 * 4 + 64;
    int grid = kNumSMs * 8;                                         // 8 CTAs per SM: each tile is a chain of barriers, the CTAs hide one another's latency
    if (total < grid) grid = (int)total;
    enc_conv7x7_tc_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(x, (const __nv_bfloat16*)wq, bias, (__nv_bfloat16*)y, N, H, W, Cout,
                                                                   y_cs, neg_slope, preproc, tiles_x, tiles_y, idesc);
    return launched("enc_conv7x7_tc_kernel");
}
