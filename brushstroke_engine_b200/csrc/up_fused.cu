// One kernel per up-sampling SynthesisLayer (SG2/training/networks.py:362-391 with up = 2; the convolution is
// conv2d_resample's transposed-conv path, SG2/torch_utils/ops/conv2d_resample.py:124-142):
//
//   T = conv_transpose2d(x * styles, W, stride 2)            (2H+1) x (2W+1), tcgen05, 9 taps per INPUT pixel
//   y = clamp(lrelu(4 * FIR4x4(pad1(T)) * dcoef + noise + bias) * gain) * next_styles          2H x 2W
//
// nbe_convT3x3s2_flat_bf16 + nbe_fir_act_nhwc_bf16 do the same in two launches and hand T over through HBM (at 128^2 and
// batch 256: 1.04 GB written, 1.09 GB read back).  Here T never leaves the chip's L2:
//
//   * a persistent CTA PAIR walks a contiguous run of 128-position items of the flat (zero-gapped) input, top to bottom
//     through its images; warp 0 = TMA, warp 1 = tcgen05 issuer (cta_group::2, all 9 x K-chunk weight half-tiles RESIDENT
//     in shared memory), exactly the machinery of conv_flat.cu;
//   * the four output-parity classes of an item are computed as two phases by ROW parity -- {(0,0),(0,1)} then {(1,0),(1,1)} --
//     so that one phase's two accumulators are the two column parities of ONE T row: warps 2-5 ("writers") read them from
//     TMEM, round to bf16 and store them into a per-CTA ring of the last R T rows in global memory (R = 16 at 64 -> 128), laid
//     out [row % R][channel pair][column parity][X + 1][2 channels] so that a warp's 32 positions are 128 contiguous bytes;
//   * warps 6-15 ("FIR") follow one item behind: as soon as the T rows oy-1 .. oy+2 of an output row are complete they
//     read them back (L2 hits: the ring is 0.5 MB per CTA and is rewritten every R rows) -- ALL rows of a work unit are
//     requested before the first one is used, so the L2 latency is paid once per unit --, filter them separably (horizontal
//     neighbours come from warp shuffles, vertical ones from a register window), apply the SynthesisLayer epilogue and
//     write y with whole 32-byte sectors through a small shared-memory transposition.
//
// The arithmetic (bf16 rounding of T, order of the FIR's fused multiply-adds, epilogue) is that of the two-kernel path, so
// the two agree BIT FOR BIT -- which is how tests/test_up_fused_gpu.py checks this kernel.
#include "tc_common.cuh"
#include <mutex>
#include <algorithm>
#include <cstdlib>

namespace nbe {

// 16 warps = 4 per SM sub-partition, whose 16 K registers then allow 128 per thread (a 17th warp would cap every thread at 96)
constexpr int U_THREADS = 512;                                       // warps: 0 TMA, 1 MMA, 2..5 T writers, 6..15 FIR
constexpr int U_FIR_WARPS = 10;
constexpr int U_MAX_ENT = 18;                                        // 9 taps x <= 2 K chunks (Cin <= 128)
constexpr int U_BHALF = 64 * 128;                                    // this CTA's half of a [128 Cout x 64 Cin] weight tile
constexpr int U_STAGE = 4 * 16 * 32;                                 // per FIR warp: 4 output rows x 16 pixels x 16 channels (32 bytes)

struct UpParams {
    __nv_bfloat16* y; int y_cs; long long y_row_pitch, y_img_pitch;
    __nv_bfloat16* scratch; long long scratch_cta;                    // ring base, elements per CTA
    int ring;                                                         // T rows kept per CTA (power of two)
    int N, H, W, P, positions, ipi, OH, OW, warm;
    int k_chunks, n_ent, ph_e0[2], ph_e1[2];
    uint32_t ent_w[U_MAX_ENT + 1];
    unsigned char ent_btile[U_MAX_ENT], ent_bk[U_MAX_ENT];
    int min_shift, box_rows, a_bytes;
    const float* f; float fgain;
    const float* scale; const float* noise; long long noise_sn; float noise_gain; const float* bias;
    float alpha, gain, clamp; const float* next_scale;
    uint32_t idesc, smem_need;
};

// output rows that are final once items 0..k of an image have been written: every T row below 2 * (complete grid rows)
__device__ __forceinline__ int rows_ready(int k, const UpParams& p) {
    if (k < 0) return 0;
    const int yc = min(((k + 1) * 128) / p.P, p.H + 1);               // complete grid rows
    return max(0, min(p.OH, 2 * yc - 2));
}

__device__ __forceinline__ uint32_t ld_g4(const __nv_bfloat16* p) {
    uint32_t v;
    asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(U_THREADS, 1)
up_layer_fused_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const UpParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    {
        uint32_t dyn;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if ((uint32_t)(smem - smem_raw) + p.smem_need > dyn) __trap();
    }
    uint8_t* smem_a = smem;                                           // [2][window of one 64-channel chunk]
    uint8_t* smem_b = smem + 2 * p.a_bytes;                           // [n_ent][8 KiB] resident weights
    uint8_t* smem_stage = smem_b + p.n_ent * U_BHALF;                 // [U_FIR_WARPS][U_STAGE]
    float* s_vec = reinterpret_cast<float*>(smem_stage + U_FIR_WARPS * U_STAGE);   // [3][128]: scale, bias, next_scale of this CTA's image
    float* s_f = s_vec + 3 * 128;                                     // [16] flipped filter * fgain
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_f + 16);
    uint64_t* a_full = bars;                 // [2] (the leader's is the live one)
    uint64_t* a_empty = bars + 2;            // [2]
    uint64_t* acc_full = bars + 4;           // [2]
    uint64_t* acc_empty = bars + 6;          // [2] (the leader's: both CTAs' writer warps arrive on it)
    uint64_t* res_full = bars + 8;           // resident weights of both CTAs landed
    uint64_t* t_ready = bars + 9;            // [2] writers -> FIR (local)
    uint64_t* fir_done = bars + 11;          // [2] FIR -> writers (local)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = (int)uniform_u32(threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int rank = (int)uniform_u32(cluster_ctarank());
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), 1);
            mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 8);
            mbar_init(smem_u32(&t_ready[i]), 4); mbar_init(smem_u32(&fir_done[i]), U_FIR_WARPS);
        }
        mbar_init(smem_u32(res_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_b) : "memory");
    }
    if (threadIdx.x < 16) {
        const int a = threadIdx.x >> 2, b = threadIdx.x & 3;
        s_f[threadIdx.x] = p.f[(3 - a) * 4 + (3 - b)] * p.fgain;        // flip_filter = False
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    // this pair's run of (image pair, item) work units; a run that starts inside an image first recomputes `warm` items whose
    // T rows its first output rows need (their own output rows belong to the previous pair)
    const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const long long G = (long long)((p.N + 1) >> 1) * p.ipi;
    const int g_own = (int)(G * cid / n_clusters), g_end = (int)(G * (cid + 1) / n_clusters);
    int g_begin = g_own;
    { const int k = g_own % p.ipi; g_begin = g_own - min(k, p.warm); }
    const int n_items = g_end - g_begin;

    if (warp == 0) {
        // ============================== TMA producer (both CTAs) ==============================
        if (lane == 0 && n_items > 0) {
            const uint32_t rf = smem_u32(res_full);
            if (leader) mbar_expect_tx(rf, 2u * (uint32_t)p.n_ent * U_BHALF);
            for (int e = 0; e < p.n_ent; ++e)
                tma_load_3d_2sm(smem_u32(smem_b + e * U_BHALF), &tmap_b, rf, p.ent_bk[e] * 64, rank * 64, p.ent_btile[e]);
            uint32_t acnt = 0;
            for (int g = g_begin; g < g_end; ++g) {
                const int ip = g / p.ipi, k = g - ip * p.ipi;
                const int n = min(2 * ip + rank, p.N - 1);
                const int q0 = k * 128;
                for (int ph = 0; ph < 2; ++ph)
                    for (int c = 0; c < p.k_chunks; ++c) {
                        const int slot = (int)(acnt & 1u);
                        const uint32_t par = (acnt >> 1) & 1u;
                        ++acnt;
                        mbar_wait_fast(smem_u32(&a_empty[slot]), par ^ 1);
                        const uint32_t full = smem_u32(&a_full[slot]);
                        if (leader) mbar_expect_tx(full, 2u * (uint32_t)(p.box_rows * 128));
                        tma_load_3d_2sm(smem_u32(smem_a + slot * p.a_bytes), &tmap_a, full, c * 64, q0 + p.min_shift, n);
                    }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer (leader CTA; whole warp, elected lane issues) ==============================
        if (leader && n_items > 0) {
            const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a)), b_lo0 = umma_desc_lo(smem_u32(smem_b));
            const uint32_t a_step = (uint32_t)p.a_bytes >> 4;
            const uint32_t idesc = p.idesc;
            const int k_chunks = p.k_chunks;
            mbar_wait(smem_u32(res_full), 0);
            tcgen05_fence_after();
            uint32_t acc_par[2] = {0, 0};
            uint32_t acnt = 0;
            for (int s = 0; s < 2 * n_items; ++s) {
                const int ph = s & 1, ab = s & 1;                       // phase ph of every item uses accumulator set ph
                const int e0 = p.ph_e0[ph], e1 = p.ph_e1[ph];
                mbar_wait_fast(smem_u32(&acc_empty[ab]), acc_par[ab] ^ 1);
                acc_par[ab] ^= 1;
                tcgen05_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(ab * 256);
                int e = e0;
                uint32_t w = p.ent_w[e0];
                for (int c = 0; c < k_chunks; ++c) {
                    const int slot = (int)(acnt & 1u);
                    const uint32_t a_par = (acnt >> 1) & 1u;
                    ++acnt;
                    mbar_wait_fast(smem_u32(&a_full[slot]), a_par);
                    tcgen05_fence_after();
                    const uint32_t a_lo = a_lo0 + (uint32_t)slot * a_step;
                    while (e < e1 && (int)(w >> 27) == c) {
                        const uint32_t wn = p.ent_w[e + 1];
                        const uint32_t b_lo = b_lo0 + (uint32_t)e * (U_BHALF >> 4);
                        const uint32_t al = a_lo + (w & 0xFFFFu);
                        const uint32_t d = d0 + ((w >> 16) & 0x3FFu);
                        const uint32_t acc0 = ((w >> 26) & 1u) ^ 1u;
                        if (elect_one()) {
                            umma_bf16_lo_2sm(d, al, b_lo, idesc, acc0);
                            umma_bf16_lo_2sm(d, al + 2, b_lo + 2, idesc, 1u);
                            umma_bf16_lo_2sm(d, al + 4, b_lo + 4, idesc, 1u);
                            umma_bf16_lo_2sm(d, al + 6, b_lo + 6, idesc, 1u);
                        }
                        ++e; w = wn;
                    }
                    if (elect_one()) umma_commit_2sm(smem_u32(&a_empty[slot]));
                }
                if (elect_one()) umma_commit_2sm(smem_u32(&acc_full[ab]));
            }
        }
    } else if (warp < 6) {
        // ============================== T writers (warps 2..5, both CTAs) ==============================
        // TMEM lane quarter qd = warp % 4 (hardware rule) -> positions qd*32 + lane; each warp stores all 128 channels of its positions.
        // Phase ph of an item holds T row 2Y + ph of every position: accumulator 0 = even columns (2X), 1 = odd columns (2X + 1).
        const int qd = warp & 3;
        const int m = qd * 32 + lane;
        __nv_bfloat16* ring = p.scratch + (long long)blockIdx.x * p.scratch_cta;
        const int plane = (p.W + 2) * 2;                               // elements of one [X + 1][2 ch] plane
        const int rmask = p.ring - 1;
        uint32_t acc_phase[2] = {0, 0};
        for (int i = 0; i < n_items; ++i) {
            const int g = g_begin + i;
            const int ip = g / p.ipi, k = g - ip * p.ipi;
            const int q = k * 128 + m;
            const int Y = q / p.P, X = q - Y * p.P;
            // ring rows are numbered continuously over the images of this run (2H + 2 per image), so that the first rows of an
            // image never land on the slots the FIR warps are still reading for the last rows of the previous one
            const int tb = (ip - g_begin / p.ipi) * (2 * p.H + 2);
            if (i >= 2) mbar_wait(smem_u32(&fir_done[i & 1]), (uint32_t)(((i - 2) >> 1) & 1));   // FIR step i-2 no longer reads the ring slots this item overwrites
#pragma unroll 1
            for (int ph = 0; ph < 2; ++ph) {
                const int ab = ph;
                mbar_wait(smem_u32(&acc_full[ab]), acc_phase[ab]);
                acc_phase[ab] ^= 1;
                tcgen05_fence_after();
                const bool row_ok = q < p.positions && Y <= p.H - ph;
                const bool ok0 = row_ok && X <= p.W, ok1 = row_ok && X < p.W;
                const int t = 2 * Y + ph;
                __nv_bfloat16* rowp = ring + ((long long)((tb + t) & rmask) * 64 * 2) * plane + (X + 1) * 2;
#pragma unroll 1
                for (int c32 = 0; c32 < 4; ++c32) {
                    uint32_t v0[32], v1[32];
                    const uint32_t ta = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(ab * 256 + c32 * 32);
                    tmem_ld32_nowait(ta, v0);
                    tmem_ld32_nowait(ta + 128, v1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int ch2 = c32 * 16 + j;
                        __nv_bfloat16* dst = rowp + (long long)ch2 * 2 * plane;
                        if (ok0) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
                        if (ok1) *reinterpret_cast<__nv_bfloat162*>(dst + plane) = __floats2bfloat162_rn(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(smem_u32(&acc_empty[ab]));
            }
            // both T rows of this item are in the ring: hand the item to the FIR warps (release at CTA scope)
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&t_ready[i & 1])) : "memory");
        }
    } else {
        // ============================== FIR + epilogue (warps 6..15, both CTAs) ==============================
        // lane = (channel pair cq of 4, position xl of 8): a warp filters 8 input columns (16 output pixels) x 8 channels at a time;
        // two such passes fill the 16 channels (32 bytes) of its 16 pixels in the staging buffer, which is then written out as
        // whole sectors.  Work units (segment of 8 columns, group of 16 channels) are dealt to the FIR warps round robin; a unit is
        // processed in groups of <= 4 output rows whose <= 7 T rows are ALL requested before the first one is used.
        const int wf = warp - 6;
        const int xl = lane & 7, cq = lane >> 3;
        const __nv_bfloat16* ring = p.scratch + (long long)blockIdx.x * p.scratch_cta;
        const int plane = (p.W + 2) * 2;
        const int rmask = p.ring - 1;
        uint8_t* stg = smem_stage + wf * U_STAGE;
        {   // the window arithmetic below is the rank-1 form f = fy (x) fx; anything else must take the two-kernel path
            bool sep = s_f[0] != 0.f;
            for (int a = 1; a < 4; ++a)
                for (int b = 1; b < 4; ++b)
                    sep = sep && fabsf(s_f[a * 4 + b] * s_f[0] - s_f[a * 4] * s_f[b]) <= 1e-6f * fabsf(s_f[a * 4 + b] * s_f[0]) + 1e-30f;
            if (!sep) {
                if (threadIdx.x == 192 && blockIdx.x == 0) printf("nbe up_layer_fused: the resample filter is not separable\n");
                __trap();
            }
        }
        float2 fx2[4], fy2[4];
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) { fx2[i2] = make_float2(s_f[i2], s_f[i2]); const float v = s_f[i2 * 4] / s_f[0]; fy2[i2] = make_float2(v, v); }
        const float g_pre = p.gain;                                     // lrelu(a) * gain == max(a*gain, a*gain*alpha): folded into scale / bias / noise
        const float clamp_hi = p.clamp >= 0.f ? p.clamp : INFINITY;
        const float ngain = p.noise_gain * g_pre;
        const float2 alpha2 = {p.alpha, p.alpha};
        const int n_units = (p.W >> 3) * 8;                              // (segments of 8 columns) x (8 groups of 16 channels)
        int cur_n = -1;
        for (int i = 0; i < n_items; ++i) {
            const int g = g_begin + i;
            const int ip = g / p.ipi, k = g - ip * p.ipi;
            const int n = 2 * ip + rank;
            const int tb = (ip - g_begin / p.ipi) * (2 * p.H + 2);      // running ring row of this image's T row 0 (see the writers)
            mbar_wait(smem_u32(&t_ready[i & 1]), (uint32_t)((i >> 1) & 1));
            if (g >= g_own && n < p.N) {
                const int r_lo = rows_ready(k - 1, p), r_hi = rows_ready(k, p);
                if (n != cur_n) {
                    asm volatile("bar.sync 2, 320;" ::: "memory");       // every FIR warp is done with the previous image's vectors
                    const int et = threadIdx.x - 192;
                    if (et < 128) {
                        s_vec[et] = (p.scale ? p.scale[(long long)n * 128 + et] : 1.f) * g_pre;
                        s_vec[128 + et] = (p.bias ? p.bias[et] : 0.f) * g_pre;
                        s_vec[256 + et] = p.next_scale ? p.next_scale[(long long)n * 128 + et] : 1.f;
                    }
                    asm volatile("bar.sync 2, 320;" ::: "memory");
                    cur_n = n;
                }
#pragma unroll 1
                for (int u = wf; u < n_units && r_hi > r_lo; u += U_FIR_WARPS) {
                    const int seg = u >> 3, cp = u & 7;
                    const int X = seg * 8 + xl;
#pragma unroll 1
                    for (int r0 = r_lo; r0 < r_hi; r0 += 4) {
                        const int nr = min(4, r_hi - r0);
#pragma unroll 1
                        for (int half = 0; half < 2; ++half) {
                            const int ch2 = (cp * 2 + half) * 4 + cq;    // channel pair of this lane
                            const __nv_bfloat16* colp = ring + (long long)ch2 * 2 * plane + (X + 1) * 2;   // + slot * 128 * plane ; odd columns at + plane
                            // ---- every T row this group needs, requested at once: rows r0-1 .. r0+nr+1
                            uint32_t v0[7], v1[7], ea[7], ed[7], ee[7];
#pragma unroll
                            for (int j = 0; j < 7; ++j) {
                                const int t = r0 - 1 + j;
                                v0[j] = v1[j] = ea[j] = ed[j] = ee[j] = 0u;
                                if (j < nr + 3 && t >= 0 && t <= 2 * p.H) {
                                    const __nv_bfloat16* rp = colp + (long long)((tb + t) & rmask) * 128 * plane;
                                    v0[j] = ld_g4(rp);
                                    v1[j] = ld_g4(rp + plane);
                                    if (xl == 0) ea[j] = ld_g4(rp + plane - 2);                       // odd column of X - 1 (zero guard at X = -1)
                                    if (xl == 7) { ed[j] = ld_g4(rp + 2); ee[j] = ld_g4(rp + plane + 2); }   // columns of X + 1 (zero guard at X = W)
                                }
                            }
                            const float2 sc2 = *reinterpret_cast<const float2*>(s_vec + ch2 * 2);
                            const float2 bs2 = *reinterpret_cast<const float2*>(s_vec + 128 + ch2 * 2);
                            const float2 ns2 = *reinterpret_cast<const float2*>(s_vec + 256 + ch2 * 2);
                            const float* nzp = p.noise ? p.noise + (long long)n * p.noise_sn + (long long)r0 * p.OW + 2 * X : nullptr;
                            float2 h[7][2];                              // horizontally filtered rows [T row][pixel of the pair], channel pair packed
#pragma unroll
                            for (int j = 0; j < 7; ++j) {
                                if (j < nr + 3) {
                                    uint32_t a = __shfl_up_sync(0xffffffffu, v1[j], 1, 8);
                                    uint32_t d = __shfl_down_sync(0xffffffffu, v0[j], 1, 8);
                                    uint32_t e = __shfl_down_sync(0xffffffffu, v1[j], 1, 8);
                                    if (xl == 0) a = ea[j];
                                    if (xl == 7) { d = ed[j]; e = ee[j]; }
                                    // T columns 2X-1 .. 2X+3 = a, v0, v1, d, e ; outputs 2X (a v0 v1 d) and 2X+1 (v0 v1 d e)
                                    const float2 c_[5] = {bf16x2_to_f2(a), bf16x2_to_f2(v0[j]), bf16x2_to_f2(v1[j]), bf16x2_to_f2(d), bf16x2_to_f2(e)};
#pragma unroll
                                    for (int px = 0; px < 2; ++px)
                                        h[j][px] = fma2(fx2[3], c_[px + 3], fma2(fx2[2], c_[px + 2], fma2(fx2[1], c_[px + 1], mul2(fx2[0], c_[px]))));
                                }
                                if (j >= 3 && j < nr + 3) {
                                    const int r = j - 3;                     // output row r0 + r  (T rows r0+r-1 .. r0+r+2 = h[j-3 .. j])
                                    float2 nz2 = {0.f, 0.f};
                                    if (nzp) nz2 = *reinterpret_cast<const float2*>(nzp + (long long)r * p.OW);
#pragma unroll
                                    for (int px = 0; px < 2; ++px) {
                                        const float nzg = (px ? nz2.y : nz2.x) * ngain;
                                        const float2 nzv = {nzg, nzg};
                                        const float2 acc = fma2(fy2[3], h[j][px], fma2(fy2[2], h[j - 1][px], fma2(fy2[1], h[j - 2][px], mul2(fy2[0], h[j - 3][px]))));
                                        float2 a2 = fma2(acc, sc2, add2(nzv, bs2));
                                        const float2 m2 = mul2(a2, alpha2);
                                        a2.x = fmaxf(a2.x, m2.x); a2.y = fmaxf(a2.y, m2.y);
                                        a2.x = fminf(fmaxf(a2.x, -clamp_hi), clamp_hi); a2.y = fminf(fmaxf(a2.y, -clamp_hi), clamp_hi);
                                        const float2 o = mul2(a2, ns2);
                                        // staging: [row][pixel 2 xl + px of the segment][16 channels]: this lane's pair at channel (half*4 + cq) * 2
                                        *reinterpret_cast<__nv_bfloat162*>(stg + r * 512 + (2 * xl + px) * 32 + (half * 4 + cq) * 4) = __floats2bfloat162_rn(o.x, o.y);
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        {   // 16 pixels x 32 bytes per row -> whole sectors: lane = (pixel, half sector)
                            const int px = lane >> 1, hf = lane & 1;
                            __nv_bfloat16* yp = p.y + (((long long)n * p.y_img_pitch + (long long)r0 * p.y_row_pitch + seg * 16 + px) * p.y_cs + cp * 16 + hf * 8);
                            for (int r = 0; r < nr; ++r) {
                                const uint4 val = *reinterpret_cast<const uint4*>(stg + r * 512 + px * 32 + hf * 16);
                                __stcs(reinterpret_cast<uint4*>(yp + (long long)r * p.y_row_pitch * p.y_cs), val);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&fir_done[i & 1])) : "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                               // the peer's smem / TMEM stay alive until the leader's last MMA is done
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace nbe

using namespace nbe;

// T rows a CTA's ring must hold: the writers of item k+1 run while the FIR warps still read for item k and must not reach
// the slots of item k-1's rows: 4 * ceil(128 / P) + 5 rows (an item spans up to ceil(128 / P) + 1 grid rows), rounded to 2^n
static int ring_rows(int W) {
    const int P = W + 1, need = 4 * ((128 + P - 1) / P) + 8;
    int r = 16;
    while (r < need) r *= 2;
    return r;
}

extern "C" int64_t nbe_up_layer_fused_scratch_bytes(int W) {
    if (W < 8) return -1;
    return (int64_t)kNumSMs * ring_rows(W) * 64 * 2 * (W + 2) * 2 * 2;
}

extern "C" int nbe_up_layer_fused_bf16(const void* x, const void* wq, const float* f, void* y, void* scratch, int64_t scratch_bytes,
                                       int N, int H, int W, int Cin, int x_cs, int x_pitch, int Cout,
                                       int y_cs, int64_t y_row_pitch, int64_t y_img_pitch, float fgain,
                                       const float* dcoef, const float* noise, int64_t noise_sn, float noise_gain,
                                       const float* bias, float alpha, float gain, float clamp, const float* next_scale,
                                       nbe_stream_t stream) {
    NBE_REQUIRE(x && wq && f && y && scratch && N >= 0 && H >= 1 && W >= 1 && Cin >= 1, "up_layer_fused: bad arguments");
    if (Cout != 128 || Cin > 128 || W % 8 != 0 || W > 120)
        return fail(NBE_EUNSUPPORTED, "up_layer_fused: needs Cout == 128, Cin <= 128, W a multiple of 8 and <= 120 (got Cout %d, Cin %d, W %d)", Cout, Cin, W);
    if (!(gain > 0.f && alpha >= 0.f && alpha <= 1.f))
        return fail(NBE_EUNSUPPORTED, "up_layer_fused: needs gain > 0 and 0 <= alpha <= 1");
    NBE_REQUIRE(x_pitch == W + 1, "up_layer_fused: input pitch must be W + 1 (one zero gap column)");
    NBE_REQUIRE(x_cs % 8 == 0 && x_cs >= Cin && y_cs % 8 == 0 && y_cs >= Cout, "up_layer_fused: channel strides must be multiples of 8");
    NBE_REQUIRE((((uintptr_t)x | (uintptr_t)wq | (uintptr_t)y | (uintptr_t)scratch) & 15) == 0, "up_layer_fused: tensors must be 16-byte aligned");
    NBE_REQUIRE(y_row_pitch >= 2 * W && y_img_pitch >= y_row_pitch * 2 * H, "up_layer_fused: bad output pitches");
    NBE_REQUIRE(scratch_bytes >= nbe_up_layer_fused_scratch_bytes(W), "up_layer_fused: scratch of %lld bytes required (zero-initialised once)",
                (long long)nbe_up_layer_fused_scratch_bytes(W));
    NBE_REQUIRE(!noise || (((uintptr_t)noise & 7) == 0 && noise_sn % 2 == 0), "up_layer_fused: noise must be 8-byte aligned");
    if (N == 0) return NBE_OK;
    // the FIR warps use the rank-1 form of the filter (fx = first row, fy = first column / corner), like fir_act_tiled_kernel
    UpParams p{};
    p.y = (__nv_bfloat16*)y; p.y_cs = y_cs; p.y_row_pitch = y_row_pitch; p.y_img_pitch = y_img_pitch;
    p.ring = ring_rows(W);
    p.scratch = (__nv_bfloat16*)scratch; p.scratch_cta = (long long)p.ring * 64 * 2 * (W + 2) * 2;
    p.N = N; p.H = H; p.W = W; p.P = x_pitch; p.positions = (H + 1) * x_pitch; p.OH = 2 * H; p.OW = 2 * W;
    p.ipi = (p.positions + 127) / 128;
    p.warm = (3 * x_pitch + 127) / 128;
    const int Cin_pad = (Cin + 63) / 64 * 64;
    p.k_chunks = Cin_pad / 64;
    // tap program: phase = output row parity py; accumulator gl = output column parity px
    struct Tap { int shift, acc, btile; };
    Tap taps[2][6]; int ntaps[2] = {0, 0};
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px)
            for (int kh = py; kh < 3; kh += 2)
                for (int kw = px; kw < 3; kw += 2) taps[py][ntaps[py]++] = {-((kh - py) / 2) * x_pitch - (kw - px) / 2, px, kh * 3 + kw};
    p.min_shift = -x_pitch - 1;
    int e = 0;
    short ent_shift[U_MAX_ENT]; unsigned char ent_acc[U_MAX_ENT], ent_first[U_MAX_ENT], ent_c[U_MAX_ENT];
    for (int ph = 0; ph < 2; ++ph) {
        p.ph_e0[ph] = e;
        bool seen[2] = {false, false};
        for (int c = 0; c < p.k_chunks; ++c)
            for (int t = 0; t < ntaps[ph]; ++t) {
                ent_c[e] = (unsigned char)c; ent_shift[e] = (short)taps[ph][t].shift; ent_acc[e] = (unsigned char)taps[ph][t].acc;
                p.ent_btile[e] = (unsigned char)taps[ph][t].btile; p.ent_bk[e] = (unsigned char)c;
                ent_first[e] = seen[taps[ph][t].acc] ? 0 : 1; seen[taps[ph][t].acc] = true;
                ++e;
            }
        p.ph_e1[ph] = e;
    }
    p.n_ent = e;
    for (int i = 0; i < e; ++i) {
        const int a_off = (ent_shift[i] - p.min_shift) * 8;
        p.ent_w[i] = (uint32_t)a_off | ((uint32_t)ent_acc[i] * 128u) << 16 | (uint32_t)ent_first[i] << 26 | (uint32_t)ent_c[i] << 27;
    }
    p.ent_w[e] = 0xFFFFFFFFu;
    const int win_rows = 128 + x_pitch + 1;
    p.box_rows = (win_rows + 7) / 8 * 8;
    NBE_REQUIRE(p.box_rows <= 256, "up_layer_fused: window too large");
    p.a_bytes = (p.box_rows * 128 + 1023) & ~1023;
    p.f = f; p.fgain = fgain; p.scale = dcoef; p.noise = noise; p.noise_sn = noise_sn; p.noise_gain = noise_gain; p.bias = bias;
    p.alpha = alpha; p.gain = gain; p.clamp = clamp; p.next_scale = next_scale;
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    p.smem_need = (uint32_t)(2 * p.a_bytes + p.n_ent * U_BHALF + U_FIR_WARPS * U_STAGE + (3 * 128 + 16) * sizeof(float) + 256);
    const size_t limit = 227 * 1024;
    if (p.smem_need + 1024 > limit) return fail(NBE_EUNSUPPORTED, "up_layer_fused: %u bytes of shared memory needed", p.smem_need);
    CUtensorMap ta, tb;
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)(H * x_pitch), (cuuint64_t)N};
        cuuint64_t strides[2] = {(cuuint64_t)x_cs * 2, (cuuint64_t)H * x_pitch * x_cs * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)p.box_rows, 1};
        int st = make_tmap(&ta, x, 3, dims, strides, box, "flat activations");
        if (st) return st;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin_pad, 128, 9};
        cuuint64_t strides[2] = {(cuuint64_t)Cin_pad * 2, (cuuint64_t)128 * Cin_pad * 2};
        cuuint32_t box[3] = {64, 64, 1};
        int st = make_tmap(&tb, wq, 3, dims, strides, box, "weights");
        if (st) return st;
    }
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    static int max_threads = 0, num_regs = 0;
    std::call_once(once, [] {
        err = cudaFuncSetAttribute(up_layer_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncAttributes at;
        if (err == cudaSuccess) err = cudaFuncGetAttributes(&at, up_layer_fused_kernel);
        if (err == cudaSuccess) { max_threads = at.maxThreadsPerBlock; num_regs = at.numRegs; }
    });
    if (err != cudaSuccess) return fail(NBE_ECUDA, "up_layer_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    if (max_threads < U_THREADS)
        return fail(NBE_EUNSUPPORTED, "up_layer_fused: the kernel (%d registers / thread) can run %d threads per CTA, %d needed", num_regs, max_threads, U_THREADS);
    const int64_t G = (int64_t)((N + 1) / 2) * p.ipi;
    NBE_REQUIRE(G <= INT32_MAX / 2, "up_layer_fused: too many work items");
    const int grid = (int)std::min<int64_t>(kNumSMs / 2, G) * 2;
    const size_t smem = std::min(limit, (size_t)p.smem_need + 1024);
    up_layer_fused_kernel<<<grid, U_THREADS, smem, (cudaStream_t)stream>>>(ta, tb, p);
    return launched("up_layer_fused_kernel");
}
