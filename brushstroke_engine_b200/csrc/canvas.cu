// Engine composite + canvas-side integer work of the patch scheduler (bit-exact index arithmetic).
#include "common.cuh"
#include <algorithm>

namespace nbe {

// rgba from uvs/colors (+ optional UVS remap), float and/or cropped uint8 HWC tile.
__global__ void __launch_bounds__(256)
triad_composite_kernel(const float* __restrict__ uvs, const float* __restrict__ colors01, const float* __restrict__ sfactor,
                       int mode, float* __restrict__ out_f32, uint8_t* __restrict__ out_u8, int N, int H, int W, int m) {
    pdl_trigger();
    pdl_wait();
    const int HW = H * W;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * HW) return;
    const int n = idx / HW, pix = idx - n * HW;
    const int y = pix / W, x = pix - y * W;
    float U = uvs[((int64_t)n * 3 + 0) * HW + pix];
    float V = uvs[((int64_t)n * 3 + 1) * HW + pix];
    float S = uvs[((int64_t)n * 3 + 2) * HW + pix];
    if (sfactor) {
        // StyleUVSMapper._map_style_s (mapper.py:53-72), same op order in fp32
        float Sp = sfactor[n] * S;
        if (Sp > 1.0f) Sp = 1.0f;
        const float delta = 1.f - Sp;
        const float uvf = (delta <= 0.000001f) ? 0.f : __fdiv_rn(delta, __fadd_rn(U, V));
        U = uvf * U; V = uvf * V; S = Sp;
    }
    const float* c = colors01 + n * 9;
    float rgba[4];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)                                // torch.sum over k of uvs_k * colors[ch,k]: ((u*c0 + v*c1) + s*c2)
        rgba[ch] = __fadd_rn(__fadd_rn(__fmul_rn(U, c[ch * 3 + 0]), __fmul_rn(V, c[ch * 3 + 1])), __fmul_rn(S, c[ch * 3 + 2]));
    rgba[3] = (mode == 0) ? __fadd_rn(U, V) : 1.f;
    if (out_f32) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) out_f32[((int64_t)n * 4 + ch) * HW + pix] = rgba[ch];
    }
    if (out_u8 && y >= m && y < H - m && x >= m && x < W - m) {
        const int T = W - 2 * m, TH = H - 2 * m;
        uchar4 px;
        uint8_t* o = reinterpret_cast<uint8_t*>(&px);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float v = __fmul_rn(rgba[ch], 255.f);
            v = fminf(fmaxf(v, 0.f), 255.f);
            o[ch] = (uint8_t)v;                                    // truncation, as tensor.to(torch.uint8)
        }
        reinterpret_cast<uchar4*>(out_u8)[((int64_t)n * TH + (y - m)) * T + (x - m)] = px;
    }
}

// CanvasPaintEngine._render_stroke_torch tail (brush.py:905-935): mode 0 'clear' (stroke colour, generated foreground alpha),
// 1 'stroke' (opaque stroke colour), 2 'canvas' (the generated canvas), 3 'full' (canvas under the stroke).
__global__ void __launch_bounds__(256)
canvas_composite_kernel(const float* __restrict__ uvs, const float* __restrict__ colors01, const float* __restrict__ alpha_fg,
                        int64_t alpha_sn, const float* __restrict__ gen_canvas, int mode, float* __restrict__ out_f32,
                        uint8_t* __restrict__ out_u8, int N, int H, int W, int m) {
    const int HW = H * W;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * HW) return;
    const int n = idx / HW, pix = idx - n * HW;
    const int y = pix / W, x = pix - y * W;
    const float U = uvs[((int64_t)n * 3 + 0) * HW + pix], V = uvs[((int64_t)n * 3 + 1) * HW + pix], S = uvs[((int64_t)n * 3 + 2) * HW + pix];
    const float a = alpha_fg[(int64_t)n * alpha_sn + pix];
    const float* c = colors01 + n * 9;
    float rgba[4];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float stroke = __fadd_rn(__fadd_rn(__fmul_rn(U, c[ch * 3 + 0]), __fmul_rn(V, c[ch * 3 + 1])), __fmul_rn(S, c[ch * 3 + 2]));
        const float cv = __fdiv_rn(__fadd_rn(gen_canvas[((int64_t)n * 3 + ch) * HW + pix], 1.0f), 2.0f);
        rgba[ch] = (mode <= 1) ? stroke : (mode == 2) ? cv : __fadd_rn(__fmul_rn(__fsub_rn(1.f, a), cv), __fmul_rn(a, stroke));
    }
    rgba[3] = (mode == 0) ? a : 1.f;
    if (out_f32) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) out_f32[((int64_t)n * 4 + ch) * HW + pix] = rgba[ch];
    }
    if (out_u8 && y >= m && y < H - m && x >= m && x < W - m) {
        const int T = W - 2 * m, TH = H - 2 * m;
        uchar4 px;
        uint8_t* o = reinterpret_cast<uint8_t*>(&px);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float v = __fmul_rn(rgba[ch], 255.f);
            v = fminf(fmaxf(v, 0.f), 255.f);
            o[ch] = (uint8_t)v;                                    // truncation, as tensor.to(torch.uint8)
        }
        reinterpret_cast<uchar4*>(out_u8)[((int64_t)n * TH + (y - m)) * T + (x - m)] = px;
    }
}

// counts[r * ncols + c] = number of stroke (== 0) pixels of the P x P crop at (r * stride, c * stride): the "> 10 stroke
// pixels" filter of generate_stitching_crops (style_transfer.py:45).  One warp per crop.
__global__ void __launch_bounds__(256)
count_stroke_kernel(const uint8_t* __restrict__ canvas, int canvas_h, int canvas_w, int P, int stride, int nrows, int ncols,
                    int32_t* __restrict__ counts) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nrows * ncols) return;
    const int y0 = (warp / ncols) * stride, x0 = (warp % ncols) * stride;
    int n = 0;
    for (int i = lane; i < P * P; i += 32) {
        const int y = y0 + i / P, x = x0 + i % P;
        if (y < canvas_h && x < canvas_w && canvas[(long long)y * canvas_w + x] == 0) ++n;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) n += __shfl_xor_sync(0xffffffffu, n, off);
    if (lane == 0) counts[warp] = n;
}

__global__ void __launch_bounds__(256)
gather_geom_kernel(const uint8_t* __restrict__ canvas, int canvas_h, int canvas_w, const int32_t* __restrict__ crops,
                   float* __restrict__ geom, int N, int P) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * P * P) return;
    const int n = idx / (P * P), r = idx - n * P * P;
    const int y = r / P, x = r - y * P;
    const int cy = crops[2 * n] + y, cx = crops[2 * n + 1] + x;
    uint8_t g = 255;
    if (cy >= 0 && cy < canvas_h && cx >= 0 && cx < canvas_w) g = canvas[(int64_t)cy * canvas_w + cx];
    const uint8_t inv = (uint8_t)(255 - g);                        // paint_image_main.py:167
    geom[idx] = __fsub_rn(1.f, __fdiv_rn((float)inv, 255.0f));     // brush.py:679
}

// owner[y,x] = max raster index of a tile covering the pixel, -1 if none (last writer wins in raster order).
// Scatter formulation: one thread per (tile, tile pixel) does an atomicMax of the tile index -- O(tiles * T^2)
// instead of O(pixels * tiles); `owner` is pre-filled with -1 by the host wrapper.
__global__ void __launch_bounds__(256)
tile_owner_kernel(const int32_t* __restrict__ tile_yx, int n_tiles, int T, int32_t* __restrict__ owner, int h, int w) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_tiles * T * T) return;
    const int t = (int)(idx / (T * T));
    const int r = (int)(idx - (int64_t)t * T * T);
    const int y = tile_yx[2 * t] + r / T, x = tile_yx[2 * t + 1] + r % T;
    if (y < 0 || y >= h || x < 0 || x >= w) return;
    atomicMax(&owner[(int64_t)y * w + x], t);
}

__global__ void __launch_bounds__(256)
place_tiles_kernel(const uint8_t* __restrict__ tiles, const int32_t* __restrict__ tile_yx, const int32_t* __restrict__ order,
                   int N, int T, const int32_t* __restrict__ owner, uint8_t* __restrict__ canvas, int h, int w) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * T * T) return;
    const int n = (int)(idx / (T * T));
    const int r = (int)(idx - (int64_t)n * T * T);
    const int ly = r / T, lx = r - ly * T;
    const int y = tile_yx[2 * n] + ly, x = tile_yx[2 * n + 1] + lx;
    if (y < 0 || y >= h || x < 0 || x >= w) return;
    if (owner && owner[(int64_t)y * w + x] != order[n]) return;
    reinterpret_cast<uchar4*>(canvas)[(int64_t)y * w + x] = reinterpret_cast<const uchar4*>(tiles)[idx];
}


// Feature blending of the painting helper (forger/ui/brush.py:190-242, forger/train/stitching.py:18-25) for a batch of
// patches whose feature windows are DISJOINT (one wavefront of the crop grid), in one pass over the block output x:
//   m      = the feature canvas already holds this pixel            alpha  = m ? base_alpha : 1
//   x_b    = bf16((1 - alpha) * saved + alpha * x)                  (BlendedFeatures.blend with weight 1 - alpha)
//   update = (base_alpha > 0.99 | (m & base_alpha > 0)) & inside the crop margin:  canvas <- x_b, mask <- 1
//   x      = x_b * scale   (the consuming layer's styles; the un-blended flat path fuses this into the conv epilogue)
// A pixel's C / 8 threads sit in one warp: all of them read the mask before the first one updates it.
__global__ void __launch_bounds__(256)
blend_window_kernel(__nv_bfloat16* __restrict__ x, int x_pitch, int x_cs, int R, int C, __nv_bfloat16* __restrict__ fcanvas,
                    uint8_t* __restrict__ fmask, int FW, const int32_t* __restrict__ fyx, const float* __restrict__ base_alpha, int cm,
                    const float* __restrict__ scale, int64_t total) {
    const int vpp = C >> 3;                                           // 16-byte vectors per pixel
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vpp);
        const int64_t pixn = i / vpp;
        const int pix = (int)(pixn % (R * R));
        const int n = (int)(pixn / (R * R));
        const int r = pix / R, c = pix - r * R;
        const int64_t fpix = (int64_t)(fyx[2 * n] + r) * FW + fyx[2 * n + 1] + c;
        const bool m = fmask[fpix] != 0;
        const float ab = base_alpha[pix];
        const bool inner = r >= cm && r < R - cm && c >= cm && c < R - cm;
        const bool update = (ab > 0.99f || (m && ab > 0.f)) && inner;
        const float alpha = m ? ab : 1.f;
        const float a = 1.f - alpha;                                  // weight of the saved features
        __nv_bfloat16* xp = x + (((int64_t)n * R + r) * x_pitch + c) * x_cs + v * 8;
        __nv_bfloat16* fp = fcanvas + fpix * C + v * 8;
        int4 xv = *reinterpret_cast<const int4*>(xp);
        const int4 sv = *reinterpret_cast<const int4*>(fp);
        __nv_bfloat16* xe = reinterpret_cast<__nv_bfloat16*>(&xv);
        const __nv_bfloat16* se = reinterpret_cast<const __nv_bfloat16*>(&sv);
#pragma unroll
        for (int e = 0; e < 8; ++e)                                   // two roundings like torch's `alpha * features + (1 - alpha) * other` (stitching.py:24-25): no FMA contraction
            xe[e] = __float2bfloat16_rn(__fadd_rn(__fmul_rn(a, __bfloat162float(se[e])), __fmul_rn(1.f - a, __bfloat162float(xe[e]))));
        __syncwarp();                                                 // every thread of the pixel has read the mask
        if (update) {
            *reinterpret_cast<int4*>(fp) = xv;
            if (v == 0) fmask[fpix] = 1;
        }
        if (scale) {
            const float* sc = scale + (int64_t)n * C + v * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) xe[e] = __float2bfloat16_rn(__bfloat162float(xe[e]) * sc[e]);
        }
        *reinterpret_cast<int4*>(xp) = xv;
    }
}

}  // namespace nbe

using namespace nbe;

extern "C" int nbe_triad_composite(const float* uvs, const float* colors01, const float* sfactor, int mode,
                                   float* out_f32, uint8_t* out_u8, int N, int H, int W, int crop_margin,
                                   nbe_stream_t stream) {
    NBE_REQUIRE(uvs && colors01 && (out_f32 || out_u8), "triad_composite: null tensor");
    NBE_REQUIRE(mode == 0 || mode == 1, "triad_composite: unknown render mode %d", mode);
    NBE_REQUIRE(N >= 0 && H >= 1 && W >= 1 && crop_margin >= 0 && 2 * crop_margin < H && 2 * crop_margin < W, "triad_composite: bad shape");
    if (N == 0) return NBE_OK;
    const int64_t total = (int64_t)N * H * W;
    NBE_REQUIRE(total <= INT32_MAX, "triad_composite: too large");
    launch_pdl(triad_composite_kernel, dim3((int)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
               uvs, colors01, sfactor, mode, out_f32, out_u8, N, H, W, crop_margin);
    return launched("triad_composite_kernel");
}

extern "C" int nbe_canvas_composite(const float* uvs, const float* colors01, const float* alpha_fg, int64_t alpha_sn,
                                    const float* gen_canvas, int mode, float* out_f32, uint8_t* out_u8,
                                    int N, int H, int W, int crop_margin, nbe_stream_t stream) {
    NBE_REQUIRE(uvs && colors01 && alpha_fg && gen_canvas && N >= 0 && H >= 1 && W >= 1, "canvas_composite: bad arguments");
    NBE_REQUIRE(mode >= 0 && mode <= 3, "canvas_composite: unknown render mode %d", mode);
    NBE_REQUIRE(crop_margin >= 0 && 2 * crop_margin < H && 2 * crop_margin < W, "canvas_composite: bad crop margin");
    NBE_REQUIRE(out_f32 || out_u8, "canvas_composite: no output");
    if (N == 0) return NBE_OK;
    const int64_t total = (int64_t)N * H * W;
    NBE_REQUIRE(total <= INT32_MAX, "canvas_composite: too large");
    canvas_composite_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        uvs, colors01, alpha_fg, alpha_sn, gen_canvas, mode, out_f32, out_u8, N, H, W, crop_margin);
    return launched("canvas_composite_kernel");
}

extern "C" int nbe_count_stroke_pixels(const uint8_t* canvas, int canvas_h, int canvas_w, int P, int stride, int nrows, int ncols,
                                       int32_t* counts, nbe_stream_t stream) {
    NBE_REQUIRE(canvas && counts && canvas_h >= 1 && canvas_w >= 1 && P >= 1 && stride >= 1 && nrows >= 0 && ncols >= 0,
                "count_stroke_pixels: bad arguments");
    const int64_t crops = (int64_t)nrows * ncols;
    if (crops == 0) return NBE_OK;
    NBE_REQUIRE(crops <= INT32_MAX / 32, "count_stroke_pixels: too many crops");
    count_stroke_kernel<<<(int)((crops * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(canvas, canvas_h, canvas_w, P, stride, nrows, ncols, counts);
    return launched("count_stroke_kernel");
}

extern "C" int nbe_gather_geom_patches(const uint8_t* canvas, int canvas_h, int canvas_w, const int32_t* crops, float* geom,
                                       int N, int P, nbe_stream_t stream) {
    NBE_REQUIRE(canvas && crops && geom && N >= 0 && P >= 1 && canvas_h >= 1 && canvas_w >= 1, "gather_geom_patches: bad arguments");
    if (N == 0) return NBE_OK;
    const int64_t total = (int64_t)N * P * P;
    NBE_REQUIRE(total <= INT32_MAX, "gather_geom_patches: too large");
    gather_geom_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(canvas, canvas_h, canvas_w, crops, geom, N, P);
    return launched("gather_geom_kernel");
}

extern "C" int nbe_tile_owner_map(const int32_t* tile_yx, int n_tiles, int T, int32_t* owner, int canvas_h, int canvas_w,
                                  nbe_stream_t stream) {
    NBE_REQUIRE(tile_yx && owner && n_tiles >= 0 && T >= 1 && canvas_h >= 1 && canvas_w >= 1, "tile_owner_map: bad arguments");
    cudaError_t e = cudaMemsetAsync(owner, 0xFF, (size_t)canvas_h * canvas_w * sizeof(int32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(NBE_ECUDA, "tile_owner_map: memset: %s", cudaGetErrorString(e));
    if (n_tiles == 0) return NBE_OK;
    const int64_t total = (int64_t)n_tiles * T * T;
    tile_owner_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(tile_yx, n_tiles, T, owner, canvas_h, canvas_w);
    return launched("tile_owner_kernel");
}

extern "C" int nbe_place_tiles(const uint8_t* tiles, const int32_t* tile_yx, const int32_t* order, int N, int T,
                               const int32_t* owner, uint8_t* canvas, int canvas_h, int canvas_w, nbe_stream_t stream) {
    NBE_REQUIRE(tiles && tile_yx && canvas && N >= 0 && T >= 1, "place_tiles: bad arguments");
    NBE_REQUIRE(!owner || order, "place_tiles: owner map given without tile order");
    if (N == 0) return NBE_OK;
    const int64_t total = (int64_t)N * T * T;
    place_tiles_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(tiles, tile_yx, order, N, T, owner, canvas, canvas_h, canvas_w);
    return launched("place_tiles_kernel");
}

extern "C" int nbe_blend_window_nhwc_bf16(void* x, int x_pitch, int x_cs, int R, int C, void* fcanvas, uint8_t* fmask, int FH, int FW,
                                          const int32_t* fyx, const float* base_alpha, int crop_margin, const float* scale, int N,
                                          nbe_stream_t stream) {
    NBE_REQUIRE(x && fcanvas && fmask && fyx && base_alpha && N >= 0 && R >= 1, "blend_window: bad arguments");
    NBE_REQUIRE(C >= 8 && C % 8 == 0 && 32 % (C / 8) == 0, "blend_window: C / 8 must divide 32 (got C = %d)", C);
    NBE_REQUIRE(x_pitch >= R && x_cs >= C && x_cs % 8 == 0 && FH >= R && FW >= R, "blend_window: bad pitches");
    NBE_REQUIRE(((int64_t)R * R * (C / 8)) % 32 == 0, "blend_window: a patch must fill whole warps");
    NBE_REQUIRE(crop_margin >= 0 && 2 * crop_margin < R, "blend_window: bad crop margin");
    NBE_REQUIRE((((uintptr_t)x | (uintptr_t)fcanvas) & 15) == 0, "blend_window: tensors must be 16-byte aligned");
    if (N == 0) return NBE_OK;
    const int64_t total = (int64_t)N * R * R * (C / 8);
    const int64_t blocks = std::min<int64_t>((total + 255) / 256, (int64_t)kNumSMs * 16);
    blend_window_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)x, x_pitch, x_cs, R, C, (__nv_bfloat16*)fcanvas, fmask, FW,
                                                                      fyx, base_alpha, crop_margin, scale, total);
    return launched("blend_window_kernel");
}
