// "Flat" shifted-window implicit GEMM on tcgen05: 3x3 stride-1 convolutions AND stride-2 transposed convolutions of
// NHWC bf16 activations whose rows are stored with a pitch P >= W + 1 and zero gap columns.
//
// With such a layout a pixel is one row of a [positions x channels] matrix (position q = y * P + x), the zero gap
// columns / TMA out-of-bounds rows supply the convolution's padding, and EVERY filter tap is the same matrix shifted by a
// constant number of rows (kh * P + kw).  One CTA therefore loads a window of positions ONCE per 64-channel chunk
// (128 * T + halo rows, 128-byte swizzled) and feeds all taps from it through UMMA descriptors whose start address is
// shifted by whole rows -- the 128-byte swizzle is a function of the absolute smem address, so a row-shifted start reads
// exactly what TMA wrote.  Weight tiles stream through a 4-deep ring and are shared by the T position tiles of the item.
//
// A launch runs a small "tap program": taps = (row shift, accumulator, weight tile); accumulators = "classes" with
// their own output mapping.  Two programs are built on the host:
//   * conv3x3 ('same' over a zero-gapped input, or 'valid' over a haloed one): 9 taps -> 1 class, T = 2 tiles per item,
//     accumulators double-buffered in TMEM (2 x 2 x 128 columns) so the epilogue overlaps the next item's MMAs;
//   * transposed conv 3x3 stride 2 (the up-sampling layers at their ALGORITHMIC cost: 9 taps per *input* pixel instead
//     of 9 per output pixel): output parity class (py, px) takes the taps with kh = py (mod 2), kw = px (mod 2):
//     4 + 2 + 2 + 1 taps -> 4 classes = 4 accumulators (all 512 TMEM columns), written interleaved into
//     T[2Y+py, 2X+px]; the 4x4 FIR + bias/activation follow in nbe_fir_act_nhwc_bf16.
#include "tc_common.cuh"
#include <mutex>

namespace nbe {

constexpr int F_BOX_ROWS = 64;
constexpr int F_BOX_BYTES = F_BOX_ROWS * 128;
constexpr int F_BSTAGES = 4;
constexpr int F_BBYTES = 128 * 128;
constexpr int F_MAX_TAPS = 9;
constexpr int F_MAX_CLASSES = 4;
constexpr int F_THREADS = 192;

struct FlatParams {
    __nv_bfloat16* y;
    int N, P, positions, tiles_per_img, T, G, nbuf, items_per_img, total_items;
    int ntaps;
    int tap_shift[F_MAX_TAPS], tap_acc[F_MAX_TAPS], tap_btile[F_MAX_TAPS], tap_first[F_MAX_TAPS];
    int cls_sy[F_MAX_CLASSES], cls_sx[F_MAX_CLASSES], cls_oy[F_MAX_CLASSES], cls_ox[F_MAX_CLASSES], cls_vy[F_MAX_CLASSES], cls_vx[F_MAX_CLASSES];
    int min_shift, n_boxes, k_chunks;
    int y_cs; long long y_row_pitch, y_img_pitch; int noise_w;
    const float* dcoef; const float* noise; long long noise_sn; float noise_gain;
    const float* bias; int act; float alpha, gain, clamp; const float* next_scale;
    uint32_t idesc;
};

template <int EPI>   // 0: accumulators are stored as they are (bf16); 1: full SynthesisLayer epilogue
__global__ void __launch_bounds__(F_THREADS, 1)
conv_tc_flat_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const FlatParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = p.n_boxes * F_BOX_BYTES;
    uint8_t* smem_a = smem;                                         // [2][n_boxes * 8 KiB]
    uint8_t* smem_b = smem + 2 * a_bytes;                           // [F_BSTAGES][16 KiB]
    float* s_vec = reinterpret_cast<float*>(smem_b + F_BSTAGES * F_BBYTES);   // [3][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_vec + 3 * 128);
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* b_full = bars + 4;        // [F_BSTAGES]
    uint64_t* b_empty = bars + 4 + F_BSTAGES;
    uint64_t* acc_full = bars + 4 + 2 * F_BSTAGES;      // [2]
    uint64_t* acc_empty = acc_full + 2;                 // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), 1);
                                      mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 4); }
        for (int i = 0; i < F_BSTAGES; ++i) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int set_cols = p.T * p.G * 128;                           // TMEM columns of one accumulator set

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            int ai = 0; uint32_t a_phase[2] = {0, 0};
            int bs = 0; uint32_t b_phase = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int n = item / p.items_per_img;
                const int q0 = (item - n * p.items_per_img) * p.T * 128;
                for (int c = 0; c < p.k_chunks; ++c) {
                    mbar_wait(smem_u32(&a_empty[ai]), a_phase[ai] ^ 1);
                    const uint32_t full = smem_u32(&a_full[ai]);
                    mbar_expect_tx(full, (uint32_t)a_bytes);
                    for (int b = 0; b < p.n_boxes; ++b)
                        tma_load_3d(smem_u32(smem_a + ai * a_bytes + b * F_BOX_BYTES), &tmap_a, full, c * 64,
                                    q0 + p.min_shift + b * F_BOX_ROWS, n);
                    a_phase[ai] ^= 1; ai ^= 1;
                    for (int t = 0; t < p.ntaps; ++t) {
                        mbar_wait(smem_u32(&b_empty[bs]), b_phase ^ 1);
                        const uint32_t bf = smem_u32(&b_full[bs]);
                        mbar_expect_tx(bf, F_BBYTES);
                        tma_load_3d(smem_u32(smem_b + bs * F_BBYTES), &tmap_b, bf, c * 64, 0, p.tap_btile[t]);
                        if (++bs == F_BSTAGES) { bs = 0; b_phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        if (lane == 0) {
            int ai = 0; uint32_t a_phase[2] = {0, 0}, acc_phase[2] = {0, 0};
            int bs = 0; uint32_t b_phase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
                const int ab = (p.nbuf == 2) ? (it & 1) : 0;
                mbar_wait(smem_u32(&acc_empty[ab]), acc_phase[ab] ^ 1);
                tcgen05_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(ab * set_cols);
                for (int c = 0; c < p.k_chunks; ++c) {
                    mbar_wait(smem_u32(&a_full[ai]), a_phase[ai]);
                    tcgen05_fence_after();
                    const uint32_t a_base = smem_u32(smem_a + ai * a_bytes);
                    for (int t = 0; t < p.ntaps; ++t) {
                        mbar_wait(smem_u32(&b_full[bs]), b_phase);
                        tcgen05_fence_after();
                        const uint64_t b_desc = umma_smem_desc(smem_u32(smem_b + bs * F_BBYTES));
                        const uint32_t accum0 = (c != 0 || !p.tap_first[t]) ? 1u : 0u;
                        for (int i = 0; i < p.T; ++i) {
                            // position tile i, tap t: rows [i*128 + shift - min_shift, +128) of the window
                            const uint64_t a_desc = umma_smem_desc(a_base + (uint32_t)((i * 128 + p.tap_shift[t] - p.min_shift) * 128));
                            const uint32_t d = d0 + (uint32_t)((i * p.G + p.tap_acc[t]) * 128);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(d, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), p.idesc, accum0 | (uint32_t)(k != 0));
                        }
                        umma_commit(smem_u32(&b_empty[bs]));
                        if (++bs == F_BSTAGES) { bs = 0; b_phase ^= 1; }
                    }
                    umma_commit(smem_u32(&a_empty[ai]));
                    a_phase[ai] ^= 1; ai ^= 1;
                }
                umma_commit(smem_u32(&acc_full[ab]));
                acc_phase[ab] ^= 1;
            }
        }
    } else {
        // ============================== epilogue (warps 2..5) ==============================
        const int qd = warp & 3;
        const int m = qd * 32 + lane;
        const int et = threadIdx.x - 64;
        uint32_t acc_phase[2] = {0, 0};
        int it = 0, cur_n = -1;
        const float pos_gain = p.gain, neg_gain = p.gain * p.alpha;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
            const int ab = (p.nbuf == 2) ? (it & 1) : 0;
            const int n = item / p.items_per_img;
            const int q0 = (item - n * p.items_per_img) * p.T * 128;
            if (EPI == 1 && n != cur_n) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                s_vec[et] = p.dcoef ? p.dcoef[(long long)n * 128 + et] : 1.f;
                s_vec[128 + et] = p.bias ? p.bias[et] : 0.f;
                s_vec[256 + et] = p.next_scale ? p.next_scale[(long long)n * 128 + et] : 1.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                cur_n = n;
            }
            mbar_wait(smem_u32(&acc_full[ab]), acc_phase[ab]);
            acc_phase[ab] ^= 1;
            tcgen05_fence_after();
#pragma unroll 1
            for (int i = 0; i < p.T; ++i) {
                const int q = q0 + i * 128 + m;
                const int gy = q / p.P, gx = q - gy * p.P;
#pragma unroll 1
                for (int g = 0; g < p.G; ++g) {
                    const bool valid = q < p.positions && gy < p.cls_vy[g] && gx < p.cls_vx[g];
                    const int oy = gy * p.cls_sy[g] + p.cls_oy[g], ox = gx * p.cls_sx[g] + p.cls_ox[g];
                    float nz = 0.f;
                    if (valid && p.noise) nz = p.noise[(long long)n * p.noise_sn + (long long)oy * p.noise_w + ox] * p.noise_gain;
                    __nv_bfloat16* yrow = p.y + ((long long)n * p.y_img_pitch + (long long)oy * p.y_row_pitch + ox) * p.y_cs;
#pragma unroll 1
                    for (int c0 = 0; c0 < 128; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(ab * set_cols + (i * p.G + g) * 128 + c0), v);
                        if (!valid) continue;
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) {
                            int4 out;
                            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (EPI == 0) {
                                    o2[e] = __floats2bfloat162_rn(__uint_as_float(v[gg * 8 + e * 2]), __uint_as_float(v[gg * 8 + e * 2 + 1]));
                                } else {
                                    float rr[2];
#pragma unroll
                                    for (int h = 0; h < 2; ++h) {
                                        const int o = c0 + gg * 8 + e * 2 + h;
                                        float a = __uint_as_float(v[gg * 8 + e * 2 + h]) * s_vec[o] + nz + s_vec[128 + o];
                                        if (p.act) {
                                            a *= (a > 0.f) ? pos_gain : neg_gain;
                                            if (p.clamp >= 0.f) a = fminf(fmaxf(a, -p.clamp), p.clamp);
                                        }
                                        rr[h] = a * s_vec[256 + o];
                                    }
                                    o2[e] = __floats2bfloat162_rn(rr[0], rr[1]);
                                }
                            }
                            *reinterpret_cast<int4*>(yrow + c0 + gg * 8) = out;
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&acc_empty[ab])) : "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host ----------------------------------------------------------------------------------------
static int launch_flat(const void* x, const void* wq, FlatParams& p, int N, int in_positions, int Cin, int x_cs, int n_wtiles,
                       cudaStream_t stream) {
    const int Cin_pad = (Cin + 63) / 64 * 64;
    p.k_chunks = Cin_pad / 64;
    int max_shift = p.tap_shift[0];
    p.min_shift = p.tap_shift[0];
    for (int t = 1; t < p.ntaps; ++t) { if (p.tap_shift[t] < p.min_shift) p.min_shift = p.tap_shift[t]; if (p.tap_shift[t] > max_shift) max_shift = p.tap_shift[t]; }
    const int win_rows = 128 * p.T + (max_shift - p.min_shift);
    p.n_boxes = (win_rows + F_BOX_ROWS - 1) / F_BOX_ROWS;
    p.tiles_per_img = (p.positions + 127) / 128;
    p.items_per_img = (p.tiles_per_img + p.T - 1) / p.T;
    const int64_t total = (int64_t)N * p.items_per_img;
    if (total > INT32_MAX) return fail(NBE_EINVAL, "conv_flat: too many work items");
    p.total_items = (int)total;
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const size_t smem = 1024 + 2 * (size_t)p.n_boxes * F_BOX_BYTES + F_BSTAGES * F_BBYTES + 3 * 128 * sizeof(float) + 256;
    if (smem > 227 * 1024) return fail(NBE_EUNSUPPORTED, "conv_flat: window of %d rows does not fit in shared memory", win_rows);
    CUtensorMap ta, tb;
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)in_positions, (cuuint64_t)N};
        cuuint64_t strides[2] = {(cuuint64_t)x_cs * 2, (cuuint64_t)in_positions * x_cs * 2};
        cuuint32_t box[3] = {64, F_BOX_ROWS, 1};
        int st = make_tmap(&ta, x, 3, dims, strides, box, "flat activations");
        if (st) return st;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin_pad, 128, (cuuint64_t)n_wtiles};
        cuuint64_t strides[2] = {(cuuint64_t)Cin_pad * 2, (cuuint64_t)128 * Cin_pad * 2};
        cuuint32_t box[3] = {64, 128, 1};
        int st = make_tmap(&tb, wq, 3, dims, strides, box, "weights");
        if (st) return st;
    }
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    std::call_once(once, [] {
        err = cudaFuncSetAttribute(conv_tc_flat_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(conv_tc_flat_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (err != cudaSuccess) return fail(NBE_ECUDA, "conv_flat: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    const int grid = total < kNumSMs ? (int)total : kNumSMs;
    const bool raw = !p.dcoef && !p.noise && !p.bias && !p.act && !p.next_scale;
    if (raw) conv_tc_flat_kernel<0><<<grid, F_THREADS, smem, stream>>>(ta, tb, p);
    else     conv_tc_flat_kernel<1><<<grid, F_THREADS, smem, stream>>>(ta, tb, p);
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
    NBE_REQUIRE(Cout == 128, "conv3x3_flat: Cout must be 128");
    NBE_REQUIRE(x_pitch >= OW + (valid ? 2 : 1), "conv3x3_flat: input pitch %d too small (needs a zero gap column / the halo)", x_pitch);
    NBE_REQUIRE(x_cs % 8 == 0 && x_cs >= Cin && y_cs % 8 == 0 && y_cs >= Cout, "conv3x3_flat: channel strides must be multiples of 8");
    NBE_REQUIRE((((uintptr_t)x | (uintptr_t)wq | (uintptr_t)y) & 15) == 0, "conv3x3_flat: tensors must be 16-byte aligned");
    NBE_REQUIRE(y_row_pitch >= OW && y_img_pitch >= y_row_pitch * OH, "conv3x3_flat: bad output pitches");
    if (N == 0) return NBE_OK;
    FlatParams p;
    p.y = (__nv_bfloat16*)y; p.N = N; p.P = x_pitch; p.positions = OH * x_pitch; p.T = 2; p.G = 1; p.nbuf = 2;
    p.ntaps = 9;
    const int off = valid ? 0 : -1;
    for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
            const int t = kh * 3 + kw;
            p.tap_shift[t] = (kh + off) * x_pitch + (kw + off); p.tap_acc[t] = 0; p.tap_btile[t] = t; p.tap_first[t] = (t == 0);
        }
    p.cls_sy[0] = 1; p.cls_sx[0] = 1; p.cls_oy[0] = 0; p.cls_ox[0] = 0; p.cls_vy[0] = OH; p.cls_vx[0] = OW;
    p.y_cs = y_cs; p.y_row_pitch = y_row_pitch; p.y_img_pitch = y_img_pitch; p.noise_w = OW;
    p.dcoef = dcoef; p.noise = noise; p.noise_sn = noise_sn; p.noise_gain = noise_gain;
    p.bias = bias; p.act = 1; p.alpha = alpha; p.gain = gain; p.clamp = clamp; p.next_scale = next_scale;
    const int in_rows = valid ? OH + 2 : OH;
    return launch_flat(x, wq, p, N, in_rows * x_pitch, Cin, x_cs, 9, (cudaStream_t)stream);
}

extern "C" int nbe_convT3x3s2_flat_bf16(const void* x, const void* wq, void* t_out,
                                        int N, int H, int W, int Cin, int x_cs, int x_pitch, int Cout, int t_cs,
                                        int64_t t_row_pitch, int64_t t_img_pitch, const float* dcoef, nbe_stream_t stream) {
    NBE_REQUIRE(x && wq && t_out && N >= 0 && H >= 1 && W >= 1 && Cin >= 1, "convT3x3s2_flat: bad arguments");
    NBE_REQUIRE(Cout == 128, "convT3x3s2_flat: Cout must be 128");
    NBE_REQUIRE(x_pitch >= W + 1, "convT3x3s2_flat: input pitch %d too small (needs a zero gap column)", x_pitch);
    NBE_REQUIRE(x_cs % 8 == 0 && x_cs >= Cin && t_cs % 8 == 0 && t_cs >= Cout, "convT3x3s2_flat: channel strides must be multiples of 8");
    NBE_REQUIRE((((uintptr_t)x | (uintptr_t)wq | (uintptr_t)t_out) & 15) == 0, "convT3x3s2_flat: tensors must be 16-byte aligned");
    NBE_REQUIRE(t_row_pitch >= 2 * W + 1 && t_img_pitch >= t_row_pitch * (2 * H + 1), "convT3x3s2_flat: bad output pitches");
    if (N == 0) return NBE_OK;
    FlatParams p;
    p.y = (__nv_bfloat16*)t_out; p.N = N; p.P = x_pitch; p.positions = (H + 1) * x_pitch; p.T = 1; p.G = 4; p.nbuf = 1;
    // T[2Y+kh, 2X+kw] += W[kh,kw] x[Y,X]  (F.conv_transpose2d, SG2/torch_utils/ops/conv2d_resample.py:124-138):
    // class (py,px) at grid (Y',X') sums the taps with kh = py, kw = px (mod 2) over x[Y' - (kh-py)/2, X' - (kw-px)/2]
    int t = 0;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            const int g = py * 2 + px;
            bool first = true;
            for (int kh = py; kh < 3; kh += 2)
                for (int kw = px; kw < 3; kw += 2) {
                    p.tap_shift[t] = -((kh - py) / 2) * x_pitch - (kw - px) / 2;
                    p.tap_acc[t] = g; p.tap_btile[t] = kh * 3 + kw; p.tap_first[t] = first ? 1 : 0;
                    first = false; ++t;
                }
            p.cls_sy[g] = 2; p.cls_sx[g] = 2; p.cls_oy[g] = py; p.cls_ox[g] = px;
            p.cls_vy[g] = py ? H : H + 1; p.cls_vx[g] = px ? W : W + 1;
        }
    p.ntaps = t;
    p.y_cs = t_cs; p.y_row_pitch = t_row_pitch; p.y_img_pitch = t_img_pitch; p.noise_w = 0;
    p.dcoef = dcoef; p.noise = nullptr; p.noise_sn = 0; p.noise_gain = 0.f;
    p.bias = nullptr; p.act = 0; p.alpha = 1.f; p.gain = 1.f; p.clamp = -1.f; p.next_scale = nullptr;
    return launch_flat(x, wq, p, N, H * x_pitch, Cin, x_cs, 9, (cudaStream_t)stream);
}
