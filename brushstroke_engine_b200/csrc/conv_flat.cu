// "Flat" shifted-window implicit GEMM on tcgen05 CTA pairs: 3x3 stride-1 convolutions AND stride-2 transposed
// convolutions of NHWC bf16 activations whose rows are stored with a pitch P >= W + 1 and zero gap columns.
//
// With such a layout a pixel is one row of a [positions x channels] matrix (position q = y * P + x), the zero gap
// columns / TMA out-of-bounds rows supply the convolution's padding, and EVERY filter tap is the same matrix shifted by a
// constant number of rows (kh * P + kw).  A CTA therefore loads a window of positions ONCE per 64-channel chunk
// (128 * T + halo rows, 128-byte swizzled) and feeds all taps from it through UMMA descriptors whose start address is
// shifted by whole rows -- the 128-byte swizzle is a function of the absolute smem address, so a row-shifted start reads
// exactly what TMA wrote.
//
// A launch runs a small "tap program": taps = (row shift, accumulator, weight tile); accumulators = "classes" with
// their own output mapping; phases = groups of classes whose accumulators fit in half of TMEM, so that the epilogue of one
// phase overlaps the MMAs of the next.  Two programs are built on the host:
//   * conv3x3 ('same' over a zero-gapped input, or 'valid' over a haloed one): 9 taps -> 1 class, T = 2 tiles per item;
//   * transposed conv 3x3 stride 2 (the up-sampling layers at their ALGORITHMIC cost: 9 taps per *input* pixel instead
//     of 9 per output pixel): output parity class (py, px) takes the taps with kh = py (mod 2), kw = px (mod 2):
//     4 + 2 + 2 + 1 taps -> 4 classes in 2 phases, written interleaved into T[2Y+py, 2X+px]; the 4x4 FIR +
//     bias/activation follow in nbe_fir_act_nhwc_bf16.
//
// The kernel runs on CTA PAIRS (cta_group::2): the two SMs of a TPC execute ONE M = 256 MMA; CTA r of the pair owns item
// 2j + r (its own position window and its own 128 accumulator lanes) and holds only HALF of every weight tile (64 of the
// 128 output channels) -- the tensor core reads both halves.  Per SM this halves the L2 -> SM weight traffic and the
// weight footprint in shared memory; when all weights of a phase fit (Cin <= 128) they stay RESIDENT while the pair
// sweeps its items phase by phase, and only position windows stream from L2 through a deep ring.
//
// Stride-2 3x3 convolutions (the geometry encoder's down-sampling layers) run on the same kernel at their algorithmic
// cost: the reflect-padded NHWC input is read as its four parity planes xp[2Y'+a, 2X'+b] (a 5-D tensor map: the pixel
// pair (b, c) is the contiguous dimension), each plane is a flat [positions x channels] matrix with pitch Wp/2, and tap
// (kh, kw) is plane (kh & 1, kw & 1) shifted by (kh >> 1) * P + (kw >> 1) rows.  A launch is therefore described by a list
// of ENTRIES (K chunk, row shift, accumulator, weight tile): for ordinary convolutions every tap meets every chunk, for
// the strided ones a chunk (= 64 channels of one plane) meets only the taps of its plane.
//
// With 2.25 taps per class tile the transposed conv leaves the single MMA-issuing thread ~64 cycles per instruction, so
// its loop is kept to a handful of instructions per tap: tap programs live in registers, descriptors are formed by 32-bit
// adds on a precomputed low word, waits have a fast path.
#include "tc_common.cuh"
#include <mutex>
#include <algorithm>
#include <cstdlib>

namespace nbe {

constexpr int F_MAX_ENT = 128;                                     // 9 taps x up to 14 K chunks (Cin <= 896)
constexpr int F_MAX_CLASSES = 4;
constexpr int F_STAGE_BYTES = 32 * 64;                             // per epilogue warp: 32 positions x 32 channels
constexpr int F_MAX_ABUF = 6;
constexpr int F_BSTAGES = 6;
constexpr int F_BHALF = 64 * 128;                                  // this CTA's half of a [128 Cout x 64 Cin] weight tile

struct FlatParams {
    __nv_bfloat16* y;
    int N, P, positions, tiles_per_img, T, items_per_img, total_items;
    // entries, sorted by A chunk within each phase: MMA group (A chunk c shifted by `shift` rows) x (weight tile btile, K block bk)
    int n_ent;
    short ent_c[F_MAX_ENT], ent_shift[F_MAX_ENT];
    unsigned char ent_acc[F_MAX_ENT], ent_first[F_MAX_ENT], ent_btile[F_MAX_ENT], ent_bk[F_MAX_ENT];
    unsigned char ent_co[F_MAX_ENT];                                // output-channel block (of 128) of the entry's weight tile: classes may be Cout slices
    // the same, packed for the MMA issuer: A start offset in 16-byte units [0,16) | accumulator column offset [16,26) | first [26] | chunk [27,32)
    uint32_t ent_w[F_MAX_ENT + 1];
    int cls_sy[F_MAX_CLASSES], cls_sx[F_MAX_CLASSES], cls_oy[F_MAX_CLASSES], cls_ox[F_MAX_CLASSES], cls_vy[F_MAX_CLASSES], cls_vx[F_MAX_CLASSES];
    int cls_co[F_MAX_CLASSES];                                      // first output channel of the class (0, or 128 for the second Cout slice)
    int n_vec;                                                      // per-channel epilogue vector length: 128, or 256 with two Cout slices
    int min_shift, n_boxes, box_rows, k_chunks;
    // phases: groups of classes whose accumulators fit in half of TMEM; each phase owns the entries [ph_e0, ph_e1)
    int n_phases, ph_e0[2], ph_e1[2], ph_G[2], ph_cls[2][F_MAX_CLASSES], Gmax;
    // resident: the weights of ONE phase stay in shared memory while the pair sweeps all of its items (phase-major order),
    // so that only the position windows stream from L2; otherwise weights stream through a F_BSTAGES ring (item-major)
    int resident, b_tiles, n_abuf;
    int debug;                                                      // NBE_FLAT_DEBUG: 1 = epilogue skipped, 2 = epilogue without its global stores (WRONG results)
    int contiguous;                                                 // item pairs are dealt in contiguous runs (1) or round robin (0)
    int nbuf;                                                       // accumulator sets in TMEM: 2 (epilogue overlaps the next MMAs) or 1
    // planes: A chunks are 64 channels of one parity plane of a padded NHWC image (stride-2 convs); cpp = chunks per plane
    int planes, cpp, plane_C, plane_rows;
    int cout_off;                                                   // first output channel of this launch within the weight tiles
    uint32_t smem_need;
    int y_cs; long long y_row_pitch, y_img_pitch; int noise_w;
    int vec_stride;                                                 // per-sample stride of dcoef / next_scale
    const float* dcoef; const float* noise; long long noise_sn; float noise_gain;
    const float* bias; int act; float alpha, gain, clamp; const float* next_scale;
    uint32_t idesc;
};

// Compile-time tap programs of the two launches that dominate (PROG 1: 3x3 convolution, PROG 2: transposed 3x3 stride-2
// convolution).  For them the MMA warp does not walk the entry table: its per-chunk loop is fully unrolled, every tap's row
// shift is (ay * P + ax) rows with ay, ax known at compile time, accumulator column and the "first MMA of an accumulator"
// flag are immediates -- the SASS is UTCHMMAs with a few uniform adds in between, as in conv_tc_row128_kernel.  The host
// checks the entry table it built against these functions before it selects PROG != 0 (launch_flat).
__host__ __device__ constexpr int prog_phases(int prog) { return prog == 2 ? 2 : 1; }
__host__ __device__ constexpr int prog_ntaps(int prog, int ph) { return prog == 1 ? 9 : (ph == 0 ? 6 : 3); }
__host__ __device__ constexpr int prog_G(int prog) { return prog == 2 ? 2 : 1; }
// transposed conv: class (py, px) sums the taps kh = py, kw = px (mod 2) over x[Y - (kh - py) / 2, X - (kw - px) / 2]; the
// window starts P + 1 rows before the tile, so a tap reads rows shifted by ((1 - dy) * P + (1 - dx))
__host__ __device__ constexpr int prog_ay(int prog, int ph, int t) {
    if (prog == 1) return t / 3;
    if (ph == 0) return (t == 0 || t == 1 || t == 4) ? 1 : 0;       // (0,0): kh = 0,0,2,2 ; (0,1): kh = 0,2
    return 1;                                                       // (1,0): kh = 1,1 ; (1,1): kh = 1
}
__host__ __device__ constexpr int prog_ax(int prog, int ph, int t) {
    if (prog == 1) return t % 3;
    if (ph == 0) return (t == 0 || t == 2 || t >= 4) ? 1 : 0;       // (0,0): kw = 0,2,0,2 ; (0,1): kw = 1,1
    return (t == 1) ? 0 : 1;                                        // (1,0): kw = 0,2 ; (1,1): kw = 1
}
__host__ __device__ constexpr int prog_acc(int prog, int ph, int t) { return prog == 1 ? 0 : (ph == 0 ? (t >= 4 ? 1 : 0) : (t >= 2 ? 1 : 0)); }
__host__ __device__ constexpr int prog_first(int prog, int ph, int t) { return prog == 1 ? (t == 0) : (ph == 0 ? (t == 0 || t == 4) : (t == 0 || t == 2)); }

// all taps of one 64-channel chunk of phase PH, issued by the elected lane of the MMA warp
template <int PROG, int PH, bool RES, int TT>
__device__ __forceinline__ void issue_chunk_prog(uint32_t a_lo, uint32_t row1, uint32_t row2, uint32_t b_lo0, uint32_t b_res, uint32_t d0, uint32_t idesc,
                                                 bool later_chunk, int& bs, uint32_t& b_par, int b_ring, uint64_t* b_full, uint64_t* b_empty) {
    constexpr int NT = prog_ntaps(PROG, PH), G = prog_G(PROG);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int ay = prog_ay(PROG, PH, t), ax = prog_ax(PROG, PH, t), acc = prog_acc(PROG, PH, t), first = prog_first(PROG, PH, t);
        uint32_t b_lo;
        if (RES) b_lo = b_res + (uint32_t)(t * (F_BHALF >> 4));
        else {
            mbar_wait_fast(smem_u32(&b_full[bs]), b_par);
            tcgen05_fence_after();
            b_lo = b_lo0 + (uint32_t)bs * (F_BHALF >> 4);
        }
        const uint32_t al0 = a_lo + (ay == 0 ? 0u : (ay == 1 ? row1 : row2)) + (uint32_t)(ax * 8);
        const uint32_t acc0 = first ? (later_chunk ? 1u : 0u) : 1u;
        if (elect_one()) {
#pragma unroll
            for (int i = 0; i < TT; ++i) {
                const uint32_t al = al0 + (uint32_t)(i * 1024);
                const uint32_t d = d0 + (uint32_t)((i * G + acc) * 128);
                umma_bf16_lo_2sm(d, al, b_lo, idesc, acc0);
                umma_bf16_lo_2sm(d, al + 2, b_lo + 2, idesc, 1u);
                umma_bf16_lo_2sm(d, al + 4, b_lo + 4, idesc, 1u);
                umma_bf16_lo_2sm(d, al + 6, b_lo + 6, idesc, 1u);
            }
            if (!RES) umma_commit_2sm(smem_u32(&b_empty[bs]));
        }
        if (!RES) { if (++bs == b_ring) { bs = 0; b_par ^= 1; } }
    }
}

// EPI 0: accumulators are stored as they are (bf16); 1: full SynthesisLayer epilogue.  EW: epilogue warps.  RES: the phase
// weights are resident in shared memory; TT: position tiles per item -- both compile-time because the single MMA-issuing
// warp sets the pace of this kernel (ncu: ~600 cycles of issue loop per 4-MMA entry against 256 cycles of tensor work, no
// barrier ever blocking it), so every runtime branch inside its loop costs throughput.
template <int EPI, int EW, bool RES, int TT, int PROG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + EW * 32, 1)
conv_tc_flat_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const FlatParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    pdl_trigger();
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    {
        uint32_t dyn;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if ((uint32_t)(smem - smem_raw) + p.smem_need > dyn) __trap();   // the host sized the buffer for an aligned base
    }
    const int a_bytes = (p.n_boxes * p.box_rows * 128 + 1023) & ~1023;
    uint8_t* smem_a = smem;                                         // [n_abuf][window of one 64-channel chunk]
    uint8_t* smem_b = smem + p.n_abuf * a_bytes;                    // [b_tiles][8 KiB]: ring, or the resident phase weights
    uint8_t* smem_stage = smem_b + p.b_tiles * F_BHALF;             // [EW][F_STAGE_BYTES] epilogue transposition buffers
    float* s_vec = reinterpret_cast<float*>(smem_stage + EW * F_STAGE_BYTES);   // [3][256] (EPI == 1)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_vec + (EPI == 1 ? 3 * 256 : 0));
    uint64_t* a_full = bars;                            // [F_MAX_ABUF]  (the leader's is the live one)
    uint64_t* a_empty = a_full + F_MAX_ABUF;            // [F_MAX_ABUF]
    uint64_t* b_full = a_empty + F_MAX_ABUF;            // [F_BSTAGES]
    uint64_t* b_empty = b_full + F_BSTAGES;             // [F_BSTAGES]
    uint64_t* acc_full = b_empty + F_BSTAGES;           // [2]
    uint64_t* acc_empty = acc_full + 2;                 // [2]  (the leader's: both CTAs' epilogue warps arrive on it)
    uint64_t* res_full = acc_empty + 2;                 // resident weights of both CTAs landed
    uint64_t* res_free = res_full + 1;                  // all MMAs of the phase have drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_free + 1);

    const int warp = (int)uniform_u32(threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int rank = (int)uniform_u32(cluster_ctarank());
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < F_MAX_ABUF; ++i) { mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), 1); }
        for (int i = 0; i < F_BSTAGES; ++i) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 2 * EW); }
        mbar_init(smem_u32(res_full), 1); mbar_init(smem_u32(res_free), 1);
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
    const int set_cols = TT * p.Gmax * 128;                           // TMEM columns of one accumulator set
    // schedule of this pair: n_local item pairs x n_phases, item-major (streamed weights) or phase-major (resident weights).
    // Pair j = (image pair j / items_per_img, tile j % items_per_img): CTA r takes that tile of image 2 * (j / items_per_img) + r,
    // so both CTAs share q0 and with it every descriptor offset (the last image pair may hold a dummy).
    const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int pairs = ((p.N + 1) >> 1) * p.items_per_img;
    // every pair takes one CONTIGUOUS run of item pairs [j_first, j_first + n_local): the image -- and with it the per-image epilogue
    // vectors, two named barriers and an exposed global-memory round trip -- changes once per items_per_img items instead of at
    // every item (round-robin dealing put ~n_clusters / items_per_img images between two consecutive items of a pair)
    const int j_first = p.contiguous ? (int)((long long)pairs * cid / n_clusters) : cid;
    const int j_step = p.contiguous ? 1 : n_clusters;
    const int n_local = p.contiguous ? (int)((long long)pairs * (cid + 1) / n_clusters) - j_first : (pairs - cid + n_clusters - 1) / n_clusters;
    const int n_steps = n_local * p.n_phases;

    if (warp == 0) {
        // ============================== TMA producer (both CTAs) ==============================
        if (lane == 0) {
            uint32_t acnt = 0, bcnt = 0;
            int k = 0, ph = 0;
            auto load_resident = [&](int ph_) {
                const int e0 = p.ph_e0[ph_], e1 = p.ph_e1[ph_];
                const uint32_t rf = smem_u32(res_full);
                if (leader) mbar_expect_tx(rf, 2u * (uint32_t)(e1 - e0) * F_BHALF);
                for (int e = e0; e < e1; ++e)
                    tma_load_3d_2sm(smem_u32(smem_b + (e - e0) * F_BHALF), &tmap_b, rf, p.ent_bk[e] * 64, p.cout_off + p.ent_co[e] * 128 + rank * 64, p.ent_btile[e]);
            };
            // everything this warp reads may be another kernel's output -- the activations, and the weights too when a caller
            // prepares them right before the launch (nbe_modulated_conv2d does): nothing is fetched ahead of the wait
            pdl_wait();
            if (RES && n_steps > 0) load_resident(0);
            for (int s = 0; s < n_steps; ++s) {
                const int j = j_first + k * j_step;
                const int np = j / p.items_per_img;
                const int n = min(2 * np + rank, p.N - 1);                       // dummy: loads stay in range, nothing is stored
                const int q0 = (j - np * p.items_per_img) * TT * 128;
                const int e0 = p.ph_e0[ph], e1 = p.ph_e1[ph];
                if (RES && k == 0 && ph > 0) {
                    mbar_wait(smem_u32(res_free), (uint32_t)((ph - 1) & 1));   // previous phase's MMAs are done with smem_b
                    load_resident(ph);
                }
                int e = e0;
                const int wrow0 = p.planes ? q0 / p.P : 0;                        // first plane row of the window
                for (int c = 0; c < p.k_chunks; ++c) {
                    const int slot = (int)(acnt % (uint32_t)p.n_abuf);
                    const uint32_t par = (acnt / (uint32_t)p.n_abuf) & 1u;
                    ++acnt;
                    mbar_wait_fast(smem_u32(&a_empty[slot]), par ^ 1);
                    const uint32_t full = smem_u32(&a_full[slot]);
                    if (leader) mbar_expect_tx(full, 2u * (uint32_t)(p.n_boxes * p.box_rows * 128));
                    if (p.planes) {
                        const int pl = c / p.cpp, cl = c - pl * p.cpp;
                        tma_load_5d_2sm(smem_u32(smem_a + slot * a_bytes), &tmap_a, full, (pl & 1) * p.plane_C + cl * 64, 0, pl >> 1, wrow0, n);
                    } else {
                        for (int b = 0; b < p.n_boxes; ++b)
                            tma_load_3d_2sm(smem_u32(smem_a + slot * a_bytes + b * p.box_rows * 128), &tmap_a, full, c * 64,
                                            q0 + p.min_shift + b * p.box_rows, n);
                    }
                    if (!RES)
                        for (; e < e1 && p.ent_c[e] == c; ++e) {
                            const int bs = (int)(bcnt % (uint32_t)p.b_tiles);
                            const uint32_t bpar = (bcnt / (uint32_t)p.b_tiles) & 1u;
                            ++bcnt;
                            mbar_wait_fast(smem_u32(&b_empty[bs]), bpar ^ 1);
                            const uint32_t bf = smem_u32(&b_full[bs]);
                            if (leader) mbar_expect_tx(bf, 2u * F_BHALF);
                            tma_load_3d_2sm(smem_u32(smem_b + bs * F_BHALF), &tmap_b, bf, p.ent_bk[e] * 64, p.cout_off + p.ent_co[e] * 128 + rank * 64, p.ent_btile[e]);
                        }
                }
                if (RES) { if (++k == n_local) { k = 0; ++ph; } }
                else            { if (++ph == p.n_phases) { ph = 0; ++k; } }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer (leader CTA only; whole warp, elected lane issues) ==============================
        if (leader) {
            const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a)), b_lo0 = umma_desc_lo(smem_u32(smem_b));
            const uint32_t a_step = (uint32_t)a_bytes >> 4;
            const uint32_t idesc = p.idesc;
            constexpr int T = TT;
            const int k_chunks = p.k_chunks, n_abuf = p.n_abuf;
            constexpr bool resident = RES;
            const int b_ring = p.b_tiles;
            uint32_t acc_par0 = 0, acc_par1 = 0;
            int slot = 0; uint32_t a_par = 0;
            int bs = 0; uint32_t b_par = 0;
            int k = 0, ph = 0;
            for (int s = 0; s < n_steps; ++s) {
                const int e0 = p.ph_e0[ph], e1 = p.ph_e1[ph], G = p.ph_G[ph];
                // planes: the window starts at a plane-row boundary, q0 sits (q0 mod P) rows into it
                uint32_t row0 = 0;
                if (p.planes) {
                    const int j = j_first + k * j_step;
                    const int q0 = (j % p.items_per_img) * T * 128;
                    row0 = (uint32_t)(q0 % p.P);
                }
                const int ab = (p.nbuf == 2) ? (s & 1) : 0;
                if (resident && k == 0) { mbar_wait(smem_u32(res_full), (uint32_t)(ph & 1)); tcgen05_fence_after(); }
                if (ab) { mbar_wait_fast(smem_u32(&acc_empty[1]), acc_par1 ^ 1); acc_par1 ^= 1; }
                else    { mbar_wait_fast(smem_u32(&acc_empty[0]), acc_par0 ^ 1); acc_par0 ^= 1; }
                tcgen05_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(ab * set_cols);
                const uint32_t d_tile = (uint32_t)(G * 128);
                int e = e0;
                uint32_t w = p.ent_w[e0];
                if (PROG != 0) {
                    // compile-time tap program: row shifts are (ay * P + ax) rows, ay, ax in {0, 1, 2}
                    const uint32_t row1 = (uint32_t)p.P * 8u, row2 = (uint32_t)p.P * 16u;
                    uint32_t b_res = b_lo0;
                    for (int c = 0; c < k_chunks; ++c) {
                        mbar_wait_fast(smem_u32(&a_full[slot]), a_par);
                        tcgen05_fence_after();
                        const uint32_t a_lo = a_lo0 + (uint32_t)slot * a_step;
                        if (prog_phases(PROG) == 1 || ph == 0) {
                            issue_chunk_prog<PROG, 0, RES, TT>(a_lo, row1, row2, b_lo0, b_res, d0, idesc, c != 0, bs, b_par, b_ring, b_full, b_empty);
                            b_res += (uint32_t)(prog_ntaps(PROG, 0) * (F_BHALF >> 4));
                        } else {
                            issue_chunk_prog<PROG, 1, RES, TT>(a_lo, row1, row2, b_lo0, b_res, d0, idesc, c != 0, bs, b_par, b_ring, b_full, b_empty);
                            b_res += (uint32_t)(prog_ntaps(PROG, 1) * (F_BHALF >> 4));
                        }
                        if (elect_one()) umma_commit_2sm(smem_u32(&a_empty[slot]));
                        if (++slot == n_abuf) { slot = 0; a_par ^= 1; }
                    }
                } else
                for (int c = 0; c < k_chunks; ++c) {
                    mbar_wait_fast(smem_u32(&a_full[slot]), a_par);
                    tcgen05_fence_after();
                    const uint32_t a_lo = a_lo0 + (uint32_t)slot * a_step + row0 * 8u;
                    while (e < e1 && (int)(w >> 27) == c) {
                        const uint32_t wn = p.ent_w[e + 1];          // next entry (the table has a sentinel), fetched ahead of the issue
                        uint32_t b_lo;
                        if (resident) b_lo = b_lo0 + (uint32_t)(e - e0) * (F_BHALF >> 4);
                        else {
                            mbar_wait_fast(smem_u32(&b_full[bs]), b_par);
                            tcgen05_fence_after();
                            b_lo = b_lo0 + (uint32_t)bs * (F_BHALF >> 4);
                        }
                        // position tile i: rows [i*128 + shift - min_shift, +128) of the window
                        const uint32_t al0 = a_lo + (w & 0xFFFFu);
                        const uint32_t dd = d0 + ((w >> 16) & 0x3FFu);
                        const uint32_t acc0 = ((w >> 26) & 1u) ^ 1u;
                        if (elect_one()) {
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                if (i < T) {
                                    const uint32_t al = al0 + (uint32_t)(i * 1024);
                                    const uint32_t d = dd + (uint32_t)i * d_tile;
                                    umma_bf16_lo_2sm(d, al, b_lo, idesc, acc0);
                                    umma_bf16_lo_2sm(d, al + 2, b_lo + 2, idesc, 1u);
                                    umma_bf16_lo_2sm(d, al + 4, b_lo + 4, idesc, 1u);
                                    umma_bf16_lo_2sm(d, al + 6, b_lo + 6, idesc, 1u);
                                }
                            }
                            if (!resident) umma_commit_2sm(smem_u32(&b_empty[bs]));
                        }
                        if (!resident) { if (++bs == b_ring) { bs = 0; b_par ^= 1; } }
                        ++e; w = wn;
                    }
                    if (elect_one()) umma_commit_2sm(smem_u32(&a_empty[slot]));
                    if (++slot == n_abuf) { slot = 0; a_par ^= 1; }
                }
                if (elect_one()) umma_commit_2sm(smem_u32(&acc_full[ab]));
                if (resident) { if (++k == n_local) { if (elect_one()) umma_commit_2sm(smem_u32(res_free)); k = 0; ++ph; } }
                else          { if (++ph == p.n_phases) { ph = 0; ++k; } }
            }
        }
    } else {
        // ============================== epilogue (warps 2..9, both CTAs) ==============================
        // The epilogue has 4x less MMA time to hide behind than in an ordinary 3x3 conv (2.25 taps per class tile), and even
        // for ordinary convs it is what sets the pace when each SM sub-partition holds few epilogue warps (~14 dependent
        // instructions per value behind TMEM-load / LDS latencies): it runs on EW = 8 or 16 warps: TMEM lane quarter
        // qd = warp % 4 (hardware rule), channel slice hsel = (warp - 2) / 4 of 128 / (EW / 4) channels.
        // A TMEM lane (= position) is owned by one thread, but a position's channels are contiguous bytes of the output:
        // storing them straight from the owning lanes costs one wavefront per lane and instruction.  Each warp therefore
        // transposes 32 positions x 32 channels through a swizzled shared-memory buffer and writes 64-byte runs
        // (4 lanes per position, 8 positions per instruction).
        constexpr int NC32 = 128 / (EW / 4) / 32;                       // 32-column TMEM chunks per warp and class tile: 2 (EW = 8) or 1 (EW = 16)
        pdl_wait();
        const int qd = warp & 3, hsel = (warp - 2) >> 2;
        const int m = qd * 32 + lane;
        const int et = threadIdx.x - 64;
        uint8_t* stg = smem_stage + (warp - 2) * F_STAGE_BYTES;
        const uint32_t stg_wr = smem_u32(stg) + lane * 64;
        const uint32_t stg_rd = smem_u32(stg);
        const int wr_sw = (lane >> 1) & 3;
        const int rd_row0 = lane >> 2, rd_ch = lane & 3;
        uint32_t acc_phase[2] = {0, 0};
        int cur_n = -1;
        const float pos_gain = p.gain, neg_gain = p.gain * p.alpha;
        const uint32_t s_vec_u32 = smem_u32(s_vec);
        int k = 0, ph = 0;
        for (int s = 0; s < n_steps; ++s) {
            const int j = j_first + k * j_step;
            const int np = j / p.items_per_img;
            const bool dummy = 2 * np + rank >= p.N;
            const int n = min(2 * np + rank, p.N - 1);
            const int q0 = (j - np * p.items_per_img) * TT * 128;
            const int G = p.ph_G[ph];
            const int ab = (p.nbuf == 2) ? (s & 1) : 0;
            if (EPI == 1 && n != cur_n) {
                asm volatile("bar.sync 1, %0;" :: "n"(EW * 32) : "memory");
                if (et < p.n_vec) {
                    s_vec[et] = p.dcoef ? p.dcoef[(long long)n * p.vec_stride + et] : 1.f;
                    s_vec[256 + et] = p.bias ? p.bias[et] : 0.f;
                    s_vec[512 + et] = p.next_scale ? p.next_scale[(long long)n * p.vec_stride + et] : 1.f;
                }
                asm volatile("bar.sync 1, %0;" :: "n"(EW * 32) : "memory");
                cur_n = n;
            }
            mbar_wait(smem_u32(&acc_full[ab]), acc_phase[ab]);
            acc_phase[ab] ^= 1;
            tcgen05_fence_after();
            if (!dummy && p.debug != 1) {          // (p.debug: timing experiments only, see launch_flat)
#pragma unroll 1
                for (int i = 0; i < TT; ++i) {
                    const int q = q0 + i * 128 + m;
                    const int gy = q / p.P, gx = q - gy * p.P;
#pragma unroll 1
                    for (int gl = 0; gl < G; ++gl) {
                        const int g = p.ph_cls[ph][gl];              // global class (output mapping) of local accumulator gl
                        const int cco = p.cls_co[g];
                        const bool valid = q < p.positions && gy < p.cls_vy[g] && gx < p.cls_vx[g];
                        const int oy = gy * p.cls_sy[g] + p.cls_oy[g], ox = gx * p.cls_sx[g] + p.cls_ox[g];
                        float nz = 0.f;
                        if (EPI == 1 && valid && p.noise) nz = p.noise[(long long)n * p.noise_sn + (long long)oy * p.noise_w + ox] * p.noise_gain;
                        // output pixel index of this lane's position (-1: not stored), handed to the lanes that write it
                        const long long pix = valid ? (long long)n * p.y_img_pitch + (long long)oy * p.y_row_pitch + ox : -1;
                        long long rp[4];
#pragma unroll
                        for (int r8 = 0; r8 < 4; ++r8) rp[r8] = __shfl_sync(0xffffffffu, pix, r8 * 8 + rd_row0);
                        // both 32-column halves of this warp's 64 channels are requested before the one wait
                        uint32_t vv[NC32][32];
                        {
                            const uint32_t ta = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(ab * set_cols + (i * G + gl) * 128 + hsel * (NC32 * 32));
                            tmem_ld32_nowait(ta, vv[0]);
                            if (NC32 == 2) tmem_ld32_nowait(ta + 32, vv[NC32 - 1]);
                            tmem_ld_wait();
                        }
#pragma unroll
                        for (int c32 = 0; c32 < NC32; ++c32) {
                            const int c0 = hsel * (NC32 * 32) + c32 * 32;
                            uint32_t (&v)[32] = vv[c32];
#pragma unroll
                            for (int gg = 0; gg < 4; ++gg) {
                                uint32_t o[4];
                                // per-channel vectors: broadcast loads, 16 bytes at a time (the shared-memory data pipe also feeds the UMMAs)
                                float dc[8], bs[8], ns[8];
                                if (EPI == 1) {
                                    const int ob = c0 + gg * 8;
#pragma unroll
                                    for (int h4 = 0; h4 < 2; ++h4) {
                                        lds_f4(s_vec_u32 + (uint32_t)((cco + ob + 4 * h4) * 4), dc + 4 * h4);
                                        lds_f4(s_vec_u32 + (uint32_t)((256 + cco + ob + 4 * h4) * 4), bs + 4 * h4);
                                        lds_f4(s_vec_u32 + (uint32_t)((512 + cco + ob + 4 * h4) * 4), ns + 4 * h4);
                                    }
                                }
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    float r0 = __uint_as_float(v[gg * 8 + e * 2]), r1 = __uint_as_float(v[gg * 8 + e * 2 + 1]);
                                    if (EPI == 1) {
                                        float rr[2] = {r0, r1};
#pragma unroll
                                        for (int h = 0; h < 2; ++h) {
                                            const int k = e * 2 + h;
                                            float a = rr[h] * dc[k] + nz + bs[k];
                                            if (p.act) {
                                                a *= (a > 0.f) ? pos_gain : neg_gain;
                                                if (p.clamp >= 0.f) a = fminf(fmaxf(a, -p.clamp), p.clamp);
                                            }
                                            rr[h] = a * ns[k];
                                        }
                                        r0 = rr[0]; r1 = rr[1];
                                    }
                                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(r0, r1);
                                    o[e] = *reinterpret_cast<const uint32_t*>(&b2);
                                }
                                // 16-byte chunk gg of this lane's 64-byte row, XOR-swizzled so that 8 lanes cover 8 bank groups
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                             :: "r"(stg_wr + (uint32_t)((gg ^ wr_sw) << 4)), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                            }
                            __syncwarp();
#pragma unroll
                            for (int r8 = 0; r8 < 4; ++r8) {
                                const int row = r8 * 8 + rd_row0;
                                int4 val;
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                                             : "r"(stg_rd + (uint32_t)(row * 64 + ((rd_ch ^ ((row >> 1) & 3)) << 4))));
                                long long pixel = rp[r8];
                                if (p.debug == 3 && pixel >= 0) pixel &= 0xFFFF;              // timing experiment: all stores land in 16 MB (L2-resident)
                                if (pixel >= 0 && p.debug != 2) *reinterpret_cast<int4*>(p.y + pixel * p.y_cs + cco + c0 + rd_ch * 8) = val;
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(smem_u32(&acc_empty[ab]));
            if (RES) { if (++k == n_local) { k = 0; ++ph; } }
            else            { if (++ph == p.n_phases) { ph = 0; ++k; } }
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();                                             // the peer's smem / TMEM stay alive until the leader's last MMA is done
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host ----------------------------------------------------------------------------------------
// One tap of a launch's program, before expansion into entries
struct FlatTap { int shift, acc, btile, plane, co; };               // plane: parity plane the tap reads (-1: all chunks); co: Cout block of 128

struct FlatInput {
    const void* x; int N, Cin, x_cs;
    int in_positions;                                               // flat mode: positions per image
    int planes, Hp, Wp;                                             // plane mode: padded image extent (even)
};

static int launch_flat(const FlatInput& in, const void* wq, int n_wtiles, int w_cout, FlatParams& p, const FlatTap* taps, const int* phase_ntaps,
                       cudaStream_t stream, int prog = 0) {
    const int Cin_pad = (in.Cin + 63) / 64 * 64;
    const int cpp = Cin_pad / 64;
    if (p.n_vec == 0) p.n_vec = 128;
    p.planes = in.planes; p.cpp = cpp; p.plane_C = in.Cin;
    p.k_chunks = in.planes ? 4 * cpp : cpp;
    // expand the tap program into entries sorted by chunk within each phase
    int ntaps = 0;
    for (int ph = 0; ph < p.n_phases; ++ph) ntaps += phase_ntaps[ph];
    int max_shift = taps[0].shift;
    p.min_shift = taps[0].shift;
    for (int t = 1; t < ntaps; ++t) { p.min_shift = std::min(p.min_shift, taps[t].shift); max_shift = std::max(max_shift, taps[t].shift); }
    int e = 0, t0 = 0, max_phase_tiles = 0;
    for (int ph = 0; ph < p.n_phases; ++ph) {
        p.ph_e0[ph] = e;
        bool seen[F_MAX_CLASSES] = {false, false, false, false};
        for (int c = 0; c < p.k_chunks; ++c)
            for (int t = t0; t < t0 + phase_ntaps[ph]; ++t) {
                if (in.planes && taps[t].plane != c / cpp) continue;
                if (e >= F_MAX_ENT) return fail(NBE_EUNSUPPORTED, "conv_flat: tap program too long (%d chunks x %d taps)", p.k_chunks, ntaps);
                p.ent_c[e] = (short)c; p.ent_shift[e] = (short)taps[t].shift; p.ent_acc[e] = (unsigned char)taps[t].acc;
                p.ent_btile[e] = (unsigned char)taps[t].btile; p.ent_bk[e] = (unsigned char)(in.planes ? c % cpp : c);
                p.ent_co[e] = (unsigned char)taps[t].co;
                p.ent_first[e] = seen[taps[t].acc] ? 0 : 1; seen[taps[t].acc] = true;
                ++e;
            }
        p.ph_e1[ph] = e;
        max_phase_tiles = std::max(max_phase_tiles, e - p.ph_e0[ph]);
        t0 += phase_ntaps[ph];
    }
    p.n_ent = e;
    for (int i = 0; i < e; ++i) {
        const int a_off = (p.ent_shift[i] - p.min_shift) * 8;
        if (a_off < 0 || a_off > 0xFFFF || p.ent_c[i] > 31) return fail(NBE_EUNSUPPORTED, "conv_flat: tap program out of range");
        p.ent_w[i] = (uint32_t)a_off | ((uint32_t)p.ent_acc[i] * 128u) << 16 | (uint32_t)p.ent_first[i] << 26 | (uint32_t)p.ent_c[i] << 27;
    }
    p.ent_w[e] = 0xFFFFFFFFu;                                       // sentinel: chunk 31 never matches
    if (max_shift > 32767 || p.min_shift < -32768) return fail(NBE_EUNSUPPORTED, "conv_flat: row pitch too large");
    int win_rows;
    if (in.planes) {
        // whole plane rows: the window starts at the row that holds q0 and covers 128 T + max_shift more positions
        const int rows = (p.P - 1 + 128 * p.T + max_shift + p.P - 1) / p.P;       // ceil((worst start offset + tiles + halo) / P)
        p.n_boxes = 1; p.box_rows = rows * p.P; p.plane_rows = rows;
        win_rows = p.box_rows;
        if (rows > 256 || p.P > 256) return fail(NBE_EUNSUPPORTED, "conv_flat: plane window too large");
    } else {
        win_rows = 128 * p.T + (max_shift - p.min_shift);
        p.n_boxes = (win_rows + 255) / 256;                         // TMA boxes of <= 256 rows, whole 8-row swizzle atoms
        p.box_rows = ((win_rows + p.n_boxes - 1) / p.n_boxes + 7) / 8 * 8;
        p.plane_rows = 0;
    }
    p.tiles_per_img = (p.positions + 127) / 128;
    p.items_per_img = (p.tiles_per_img + p.T - 1) / p.T;
    const int64_t total = (int64_t)in.N * p.items_per_img;
    if (total * p.n_phases > INT32_MAX) return fail(NBE_EINVAL, "conv_flat: too many work items");
    p.total_items = (int)total;
    p.N = in.N;
    p.nbuf = (2 * p.T * p.Gmax * 128 <= 512) ? 2 : 1;
    if (p.T * p.Gmax * 128 > 512) return fail(NBE_EUNSUPPORTED, "conv_flat: accumulators do not fit in TMEM");
    // kind::f16: D = f32, A = B = bf16, K-major; N = 128, M = 256 (the pair)
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const bool raw = !p.dcoef && !p.noise && !p.bias && !p.act && !p.next_scale;
    const size_t limit = 227 * 1024;
    static const bool no_resident = getenv("NBE_FLAT_NO_RESIDENT") != nullptr;
    const size_t a_bytes = ((size_t)p.n_boxes * p.box_rows * 128 + 1023) & ~(size_t)1023;
    constexpr int epi_warps = 8;                                    // (16 measured 3-7 % slower: the windows lose a buffer)
    const size_t epi_bytes = (size_t)epi_warps * F_STAGE_BYTES + (raw ? 0 : 3 * 256 * sizeof(float)) + 256;
    CUtensorMap ta, tb;
    if (in.planes) {
        // (pixel pair x channels, X', row parity, Y', image) over the padded NHWC image
        cuuint64_t dims[5] = {(cuuint64_t)2 * in.Cin, (cuuint64_t)in.Wp / 2, 2, (cuuint64_t)in.Hp / 2, (cuuint64_t)in.N};
        cuuint64_t strides[4] = {(cuuint64_t)2 * in.x_cs * 2, (cuuint64_t)in.Wp * in.x_cs * 2, (cuuint64_t)2 * in.Wp * in.x_cs * 2,
                                 (cuuint64_t)in.Hp * in.Wp * in.x_cs * 2};
        cuuint32_t box[5] = {64, (cuuint32_t)p.P, 1, (cuuint32_t)p.plane_rows, 1};
        int st = make_tmap(&ta, in.x, 5, dims, strides, box, "parity planes");
        if (st) return st;
    } else {
        cuuint64_t dims[3] = {(cuuint64_t)in.Cin, (cuuint64_t)in.in_positions, (cuuint64_t)in.N};
        cuuint64_t strides[2] = {(cuuint64_t)in.x_cs * 2, (cuuint64_t)in.in_positions * in.x_cs * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)p.box_rows, 1};
        int st = make_tmap(&ta, in.x, 3, dims, strides, box, "flat activations");
        if (st) return st;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin_pad, (cuuint64_t)w_cout, (cuuint64_t)n_wtiles};
        cuuint64_t strides[2] = {(cuuint64_t)Cin_pad * 2, (cuuint64_t)w_cout * Cin_pad * 2};
        cuuint32_t box[3] = {64, 64, 1};                            // half of the 128 output channels per CTA of the pair
        int st = make_tmap(&tb, wq, 3, dims, strides, box, "weights");
        if (st) return st;
    }
    using KernelFn = void (*)(CUtensorMap, CUtensorMap, FlatParams);
    static KernelFn generic[2][2][2] = {       // [EPI][RES][T - 1], entry-table MMA loop
        {{conv_tc_flat_kernel<0, 8, false, 1, 0>, conv_tc_flat_kernel<0, 8, false, 2, 0>}, {conv_tc_flat_kernel<0, 8, true, 1, 0>, conv_tc_flat_kernel<0, 8, true, 2, 0>}},
        {{conv_tc_flat_kernel<1, 8, false, 1, 0>, conv_tc_flat_kernel<1, 8, false, 2, 0>}, {conv_tc_flat_kernel<1, 8, true, 1, 0>, conv_tc_flat_kernel<1, 8, true, 2, 0>}}};
    static KernelFn conv_prog[2] = {conv_tc_flat_kernel<1, 8, false, 2, 1>, conv_tc_flat_kernel<1, 8, true, 2, 1>};          // [RES]: 3x3 conv, full epilogue, T = 2
    static KernelFn convt_prog[2][2] = {{conv_tc_flat_kernel<0, 8, false, 1, 2>, conv_tc_flat_kernel<0, 8, false, 2, 2>},     // [RES][T - 1]: transposed conv, raw epilogue
                                        {conv_tc_flat_kernel<0, 8, true, 1, 2>, conv_tc_flat_kernel<0, 8, true, 2, 2>}};
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    std::call_once(once, [] {
        KernelFn all[14];
        int n = 0;
        for (int i = 0; i < 8; ++i) all[n++] = generic[i >> 2][(i >> 1) & 1][i & 1];
        all[n++] = conv_prog[0]; all[n++] = conv_prog[1];
        for (int i = 0; i < 4; ++i) all[n++] = convt_prog[i >> 1][i & 1];
        for (int i = 0; i < n && err == cudaSuccess; ++i)
            err = cudaFuncSetAttribute((const void*)all[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (err != cudaSuccess) return fail(NBE_ECUDA, "conv_flat: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    const int64_t pairs = (int64_t)((in.N + 1) / 2) * p.items_per_img;
    static const int grid_pairs_env = getenv("NBE_FLAT_GRID_PAIRS") ? atoi(getenv("NBE_FLAT_GRID_PAIRS")) : 0;   // timing experiment: fewer CTA pairs
    const int grid = (int)std::min<int64_t>(grid_pairs_env > 0 ? std::min(grid_pairs_env, kNumSMs / 2) : kNumSMs / 2, pairs) * 2;
    // resident weights pay off when a pair sweeps several items per phase and the largest phase fits next to >= 3 windows
    static const bool round_robin = getenv("NBE_FLAT_ROUND_ROBIN") != nullptr;      // A/B switch: the former item order
    p.contiguous = round_robin ? 0 : 1;
    static const int debug_mode = getenv("NBE_FLAT_DEBUG") ? atoi(getenv("NBE_FLAT_DEBUG")) : 0;
    p.debug = debug_mode;
    static const int res_min_abuf = getenv("NBE_FLAT_RES_MIN_ABUF") ? atoi(getenv("NBE_FLAT_RES_MIN_ABUF")) : 3;   // experiment: resident weights next to fewer windows
    p.resident = !no_resident && pairs >= 2 * (int64_t)(grid / 2) && epi_bytes + (size_t)max_phase_tiles * F_BHALF + (size_t)res_min_abuf * a_bytes <= limit;
    p.b_tiles = p.resident ? max_phase_tiles : F_BSTAGES;
    // streamed weights: a ring of 4 tiles instead of 6 when that buys one more window buffer (windows take longer to arrive)
    if (!p.resident && (limit - epi_bytes - 4 * F_BHALF) / a_bytes > (limit - epi_bytes - (size_t)F_BSTAGES * F_BHALF) / a_bytes
        && (limit - epi_bytes - (size_t)F_BSTAGES * F_BHALF) / a_bytes < F_MAX_ABUF) p.b_tiles = 4;
    const size_t rest = epi_bytes + (size_t)p.b_tiles * F_BHALF;
    if (rest + 2 * a_bytes > limit) return fail(NBE_EUNSUPPORTED, "conv_flat: window of %d rows does not fit in shared memory", win_rows);
    p.n_abuf = (int)std::min<size_t>(F_MAX_ABUF, (limit - rest) / a_bytes);
    p.smem_need = (uint32_t)(rest + (size_t)p.n_abuf * a_bytes);
    // the kernel aligns its base to 1 KiB; dynamic shared memory normally starts aligned, so the slack is only added when it fits
    const size_t smem = std::min(limit, (size_t)p.smem_need + 1024);
    if (p.T != 1 && p.T != 2) return fail(NBE_EUNSUPPORTED, "conv_flat: 1 or 2 tiles per item");
    // compile-time tap program, if the entry table is exactly what it encodes (and the epilogue / tile variant is instantiated)
    static const bool table_only = getenv("NBE_FLAT_TABLE") != nullptr;          // A/B switch: always walk the entry table
    int use_prog = table_only ? 0 : prog;
    if (use_prog == 1 && (raw || p.T != 2)) use_prog = 0;
    if (use_prog == 2 && !raw) use_prog = 0;
    if (use_prog && (in.planes || p.n_phases != prog_phases(use_prog))) use_prog = 0;
    for (int ph = 0; use_prog && ph < p.n_phases; ++ph) {
        const int nt = use_prog == 1 ? prog_ntaps(1, 0) : prog_ntaps(2, ph);
        if (p.ph_e1[ph] - p.ph_e0[ph] != nt * p.k_chunks || p.ph_G[ph] != prog_G(use_prog)) { use_prog = 0; break; }
        for (int c = 0; use_prog && c < p.k_chunks; ++c)
            for (int t = 0; t < nt; ++t) {
                const int e_ = p.ph_e0[ph] + c * nt + t;
                const int ay = prog_ay(use_prog, ph, t), ax = prog_ax(use_prog, ph, t);
                if (p.ent_c[e_] != c || p.ent_shift[e_] - p.min_shift != ay * p.P + ax || p.ent_acc[e_] != prog_acc(use_prog, ph, t) ||
                    p.ent_first[e_] != (c == 0 && prog_first(use_prog, ph, t) ? 1 : 0)) { use_prog = 0; break; }
            }
    }
    KernelFn fn = use_prog == 1 ? conv_prog[p.resident ? 1 : 0] : use_prog == 2 ? convt_prog[p.resident ? 1 : 0][p.T - 1]
                                : generic[raw ? 0 : 1][p.resident ? 1 : 0][p.T - 1];
    launch_pdl(fn, dim3(grid), dim3(64 + epi_warps * 32), smem, stream, ta, tb, p);
    return launched("conv_tc_flat_kernel");
}

}  // namespace nbe

using namespace nbe;

extern "C" int nbe_conv3x3_flat_bf16(const void* x, const void* wq, void* y,
                                     int N, int OH, int OW, int Cin, int x_cs, int x_pitch, int valid, int Cout, int y_cs,
                                     int64_t y_row_pitch, int64_t y_img_pitch,
                                     const float* dcoef, const float* noise, int64_t noise_sn, float noise_gain,
                                     const float* bias, float alpha, float gain, float clamp, const float* next_scale,
                                     nbe_stream_t stream) {
    NBE_REQUIRE(x && wq && y && N >= 0 && OH >= 1 && OW >= 1 && Cin >= 1, "conv3x3_flat: bad arguments");
    NBE_REQUIRE(Cout >= 128 && Cout % 128 == 0, "conv3x3_flat: Cout must be a multiple of 128 (one pass per 128 output channels)");
    NBE_REQUIRE(x_pitch >= OW + (valid ? 2 : 1), "conv3x3_flat: input pitch %d too small (needs a zero gap column / the halo)", x_pitch);
    NBE_REQUIRE(x_cs % 8 == 0 && x_cs >= Cin && y_cs % 8 == 0 && y_cs >= Cout, "conv3x3_flat: channel strides must be multiples of 8");
    NBE_REQUIRE((((uintptr_t)x | (uintptr_t)wq | (uintptr_t)y) & 15) == 0, "conv3x3_flat: tensors must be 16-byte aligned");
    NBE_REQUIRE(y_row_pitch >= OW && y_img_pitch >= y_row_pitch * OH, "conv3x3_flat: bad output pitches");
    if (N == 0) return NBE_OK;
    // 256 output channels per launch as TWO accumulator classes (Cout slices of 128) over the same position windows: the windows
    // are loaded once instead of once per 128-channel pass and one launch floor disappears; the accumulators then fill TMEM
    // (2 tiles x 2 slices x 128 columns), so the epilogue is not overlapped -- which costs what the saved window loads buy
    static const bool one_pass_256 = getenv("NBE_FLAT_ONE_PASS_256") != nullptr;     // A/B switch, off: measured equal (stride-2 128->256: 0.199 vs 0.201 ms) or slower (ScaleUp conv: 0.121 vs 0.099 ms)
    const int nh = (one_pass_256 && Cout % 256 == 0) ? 2 : 1;
    for (int co = 0; co < Cout; co += 128 * nh) {
        FlatParams p{};
        p.y = (__nv_bfloat16*)y + co; p.P = x_pitch; p.positions = OH * x_pitch; p.T = 2;
        FlatTap taps[18];
        const int off = valid ? 0 : -1;
        for (int half = 0; half < nh; ++half) {
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw) taps[half * 9 + kh * 3 + kw] = {(kh + off) * x_pitch + (kw + off), half, kh * 3 + kw, -1, half};
            p.cls_sy[half] = 1; p.cls_sx[half] = 1; p.cls_oy[half] = 0; p.cls_ox[half] = 0; p.cls_vy[half] = OH; p.cls_vx[half] = OW;
            p.cls_co[half] = half * 128; p.ph_cls[0][half] = half;
        }
        p.n_phases = 1; p.ph_G[0] = nh; p.Gmax = nh; p.n_vec = 128 * nh;
        const int phase_ntaps[2] = {9 * nh, 0};
        p.y_cs = y_cs; p.y_row_pitch = y_row_pitch; p.y_img_pitch = y_img_pitch; p.noise_w = OW; p.vec_stride = Cout; p.cout_off = co;
        p.dcoef = dcoef ? dcoef + co : nullptr; p.noise = noise; p.noise_sn = noise_sn; p.noise_gain = noise_gain;
        p.bias = bias ? bias + co : nullptr; p.act = 1; p.alpha = alpha; p.gain = gain; p.clamp = clamp;
        p.next_scale = next_scale ? next_scale + co : nullptr;
        const int in_rows = valid ? OH + 2 : OH;
        FlatInput in{x, N, Cin, x_cs, in_rows * x_pitch, 0, 0, 0};
        int st = launch_flat(in, wq, 9, Cout, p, taps, phase_ntaps, (cudaStream_t)stream, nh == 1 ? 1 : 0);
        if (st) return st;
    }
    return NBE_OK;
}

extern "C" int nbe_conv3x3s2_flat_bf16(const void* xp, const void* wq, void* y,
                                       int N, int H, int W, int Cin, int Cout, int y_cs, int64_t y_row_pitch, int64_t y_img_pitch,
                                       const float* bias, float alpha, float gain, float clamp, const float* next_scale,
                                       nbe_stream_t stream) {
    NBE_REQUIRE(xp && wq && y && N >= 0 && H >= 2 && W >= 2 && (H % 2) == 0 && (W % 2) == 0, "conv3x3s2_flat: bad arguments (H, W must be even)");
    NBE_REQUIRE(Cin >= 64 && Cin % 64 == 0, "conv3x3s2_flat: Cin must be a multiple of 64 (dense channels: x_cs == Cin)");
    NBE_REQUIRE(Cout >= 128 && Cout % 128 == 0, "conv3x3s2_flat: Cout must be a multiple of 128");
    NBE_REQUIRE(y_cs % 8 == 0 && y_cs >= Cout, "conv3x3s2_flat: channel stride must be a multiple of 8");
    NBE_REQUIRE((((uintptr_t)xp | (uintptr_t)wq | (uintptr_t)y) & 15) == 0, "conv3x3s2_flat: tensors must be 16-byte aligned");
    const int OH = H / 2, OW = W / 2;
    NBE_REQUIRE(y_row_pitch >= OW && y_img_pitch >= y_row_pitch * OH, "conv3x3s2_flat: bad output pitches");
    if (N == 0) return NBE_OK;
    // out[Y, X] = sum_{kh,kw} W[kh,kw] xp[2Y + kh, 2X + kw]  (xp = the input with its 1-pixel border, (H+2) x (W+2)):
    // tap (kh, kw) reads parity plane (kh & 1, kw & 1) at plane position (Y + (kh >> 1), X + (kw >> 1))
    const int P = (W + 2) / 2;
    static const bool one_pass_256 = getenv("NBE_FLAT_ONE_PASS_256") != nullptr;     // A/B switch, off (see nbe_conv3x3_flat_bf16)
    const int nh = (one_pass_256 && Cout % 256 == 0) ? 2 : 1;
    for (int co = 0; co < Cout; co += 128 * nh) {
        FlatParams p{};
        p.y = (__nv_bfloat16*)y + co; p.P = P; p.positions = OH * P;
        // two position tiles per item share every weight tile, unless the padding of the last item costs more than that saves
        const int tiles = (OH * P + 127) / 128;
        p.T = ((tiles + 1) / 2 * 2 * 100 > tiles * 115) ? 1 : 2;
        static const int force_t = getenv("NBE_S2_T") ? atoi(getenv("NBE_S2_T")) : 0;      // A/B switch
        if (force_t == 1 || force_t == 2) p.T = force_t;
        FlatTap taps[18];
        int t = 0;
        for (int half = 0; half < nh; ++half) {
            for (int pl = 0; pl < 4; ++pl)
                for (int kh = (pl >> 1); kh < 3; kh += 2)
                    for (int kw = (pl & 1); kw < 3; kw += 2) taps[t++] = {(kh >> 1) * P + (kw >> 1), half, kh * 3 + kw, pl, half};
            p.cls_sy[half] = 1; p.cls_sx[half] = 1; p.cls_oy[half] = 0; p.cls_ox[half] = 0; p.cls_vy[half] = OH; p.cls_vx[half] = OW;
            p.cls_co[half] = half * 128; p.ph_cls[0][half] = half;
        }
        p.n_phases = 1; p.ph_G[0] = nh; p.Gmax = nh; p.n_vec = 128 * nh;
        const int phase_ntaps[2] = {9 * nh, 0};
        p.y_cs = y_cs; p.y_row_pitch = y_row_pitch; p.y_img_pitch = y_img_pitch; p.noise_w = OW; p.vec_stride = Cout; p.cout_off = co;
        p.dcoef = nullptr; p.noise = nullptr; p.noise_sn = 0; p.noise_gain = 0.f;
        p.bias = bias ? bias + co : nullptr; p.act = 1; p.alpha = alpha; p.gain = gain; p.clamp = clamp;
        p.next_scale = next_scale ? next_scale + co : nullptr;
        FlatInput in{xp, N, Cin, Cin, 0, 1, H + 2, W + 2};
        int st = launch_flat(in, wq, 9, Cout, p, taps, phase_ntaps, (cudaStream_t)stream);
        if (st) return st;
    }
    return NBE_OK;
}

extern "C" int nbe_convT3x3s2_flat_bf16(const void* x, const void* wq, void* t_out,
                                        int N, int H, int W, int Cin, int x_cs, int x_pitch, int Cout, int t_cs,
                                        int64_t t_row_pitch, int64_t t_img_pitch, const float* dcoef, nbe_stream_t stream) {
    NBE_REQUIRE(x && wq && t_out && N >= 0 && H >= 1 && W >= 1 && Cin >= 1, "convT3x3s2_flat: bad arguments");
    NBE_REQUIRE(Cout >= 128 && Cout % 128 == 0, "convT3x3s2_flat: Cout must be a multiple of 128 (one pass per 128 output channels)");
    NBE_REQUIRE(x_pitch >= W + 1, "convT3x3s2_flat: input pitch %d too small (needs a zero gap column)", x_pitch);
    NBE_REQUIRE(x_cs % 8 == 0 && x_cs >= Cin && t_cs % 8 == 0 && t_cs >= Cout, "convT3x3s2_flat: channel strides must be multiples of 8");
    NBE_REQUIRE((((uintptr_t)x | (uintptr_t)wq | (uintptr_t)t_out) & 15) == 0, "convT3x3s2_flat: tensors must be 16-byte aligned");
    NBE_REQUIRE(t_row_pitch >= 2 * W + 1 && t_img_pitch >= t_row_pitch * (2 * H + 1), "convT3x3s2_flat: bad output pitches");
    if (N == 0) return NBE_OK;
  for (int co = 0; co < Cout; co += 128) {
    FlatParams p{};
    p.y = (__nv_bfloat16*)t_out + co; p.P = x_pitch; p.positions = (H + 1) * x_pitch;
    // Cin <= 128: the phase weights stay resident, one tile per item and double-buffered accumulators.  Wider inputs stream
    // their weights and are L2-bound on them: two tiles per item share every weight tile (all 512 TMEM columns, the epilogue
    // is then not overlapped -- less than the halved weight traffic buys)
    static const bool convt_t1 = getenv("NBE_CONVT_T1") != nullptr;
    p.T = (Cin > 128 && !convt_t1 && (H + 1) * x_pitch >= 1024) ? 2 : 1;            // small maps: padding of the last item costs more
    // T[2Y+kh, 2X+kw] += W[kh,kw] x[Y,X]  (F.conv_transpose2d, SG2/torch_utils/ops/conv2d_resample.py:124-138):
    // class (py,px) at grid (Y',X') sums the taps with kh = py, kw = px (mod 2) over x[Y' - (kh-py)/2, X' - (kw-px)/2].
    // Two phases of two classes each, split by ROW parity -- {(0,0): 4 taps, (0,1): 2 taps} and {(1,0): 2 taps, (1,1): 1 tap} --
    // so that a phase's two accumulators take 256 TMEM columns and the other 256 hold the previous phase while its epilogue
    // runs, and a phase writes whole T rows (its two classes are the even and the odd pixels of the same rows).  Measured
    // against the tap-balanced split {(0,0),(1,1)} / {(0,1),(1,0)}: no difference (0.32 ms at 64 -> 128, batch 256).  What
    // this launch waits for is its global stores: 0.18 ms with the stores removed, 0.17 ms without any epilogue
    // (NBE_FLAT_DEBUG=2 / 1, timing experiments with wrong results).
    // NBE_CONVT_ONE_PHASE (experiment): all four classes in ONE phase -- every item is finished in one visit (images complete in
    // order), all nine taps stay resident, the four accumulators fill TMEM (the epilogue is not overlapped)
    static const bool one_phase = getenv("NBE_CONVT_ONE_PHASE") != nullptr;
    const int n_ph = (one_phase && p.T == 1) ? 1 : 2, cls_per_ph = 4 / n_ph;
    FlatTap taps[9];
    int phase_ntaps[2] = {0, 0};
    int t = 0;
    for (int ph = 0; ph < n_ph; ++ph) {
        for (int gl = 0; gl < cls_per_ph; ++gl) {
            const int g = ph * cls_per_ph + gl;                     // class (py, px) = (g >> 1, g & 1): a phase holds whole row parities
            const int py = g >> 1, px = g & 1;
            p.ph_cls[ph][gl] = g;
            for (int kh = py; kh < 3; kh += 2)
                for (int kw = px; kw < 3; kw += 2) { taps[t++] = {-((kh - py) / 2) * x_pitch - (kw - px) / 2, gl, kh * 3 + kw, -1}; ++phase_ntaps[ph]; }
            p.cls_sy[g] = 2; p.cls_sx[g] = 2; p.cls_oy[g] = py; p.cls_ox[g] = px;
            p.cls_vy[g] = py ? H : H + 1; p.cls_vx[g] = px ? W : W + 1;
        }
        p.ph_G[ph] = cls_per_ph;
    }
    p.n_phases = n_ph; p.Gmax = cls_per_ph;
    p.y_cs = t_cs; p.y_row_pitch = t_row_pitch; p.y_img_pitch = t_img_pitch; p.noise_w = 0; p.vec_stride = Cout; p.cout_off = co;
    p.dcoef = dcoef ? dcoef + co : nullptr; p.noise = nullptr; p.noise_sn = 0; p.noise_gain = 0.f;
    p.bias = nullptr; p.act = 0; p.alpha = 1.f; p.gain = 1.f; p.clamp = -1.f; p.next_scale = nullptr;
    FlatInput in{x, N, Cin, x_cs, H * x_pitch, 0, 0, 0};
    int st = launch_flat(in, wq, 9, Cout, p, taps, phase_ntaps, (cudaStream_t)stream, n_ph == 2 ? 2 : 0);
    if (st) return st;
  }
    return NBE_OK;
}
