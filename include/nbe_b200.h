/*
 * nbe_b200.h -- C ABI of libnbe_b200.so, the B200 (sm_100a) implementation of the
 * NeuBE generator-forward hot path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers / sizes / a
 * cudaStream_t (as void*), launches asynchronously on that stream, never
 * synchronises, never allocates device memory that outlives the call (the caller
 * -- the Python shim, via torch's allocator -- owns every buffer, matching the
 * reference where the plugin returns a torch-allocated tensor), and returns
 * 0 on success or a negative NBE_E* code; nbe_last_error() gives the message.
 * Nothing throws across the boundary.  There is no CPU fallback: without a
 * CUDA device the calls return NBE_ECUDA.
 *
 * "Replaces" lines cite the reference interface (paths relative to
 * /root/reference; SG2 = thirdparty/stylegan2_ada_pytorch).
 */
#ifndef NBE_B200_H_
#define NBE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBE_ABI_VERSION 1

typedef void* nbe_stream_t;                 /* cudaStream_t */

enum nbe_status { NBE_OK = 0, NBE_EINVAL = -1, NBE_ECUDA = -2, NBE_EUNSUPPORTED = -3 };
enum nbe_dtype  { NBE_F32 = 0, NBE_F16 = 1, NBE_BF16 = 2, NBE_F64 = 3 };

/* activation indices = the reference's `cuda_idx` (SG2/torch_utils/ops/bias_act.py:23-33) */
enum nbe_act { NBE_ACT_LINEAR = 1, NBE_ACT_RELU = 2, NBE_ACT_LRELU = 3, NBE_ACT_TANH = 4, NBE_ACT_SIGMOID = 5,
               NBE_ACT_ELU = 6, NBE_ACT_SELU = 7, NBE_ACT_SOFTPLUS = 8, NBE_ACT_SWISH = 9 };

int         nbe_abi_version(void);
const char* nbe_last_error(void);
/* Number of kernels this library has launched since load (bench.py's `gpu_launches`). */
int64_t     nbe_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * bias_act  --  y = clamp(act(x + b[(i / step_b) % size_b]) * gain, +-clamp), clamp < 0 = off.
 * Replaces: bias_act_plugin.bias_act(x, b, xref, yref, dy, grad=0, dim, act, alpha, gain, clamp)
 *           SG2/torch_utils/ops/bias_act.cpp:32-90, kernel bias_act.cu:23-147 (forward, grad == 0).
 * x, y: dense, size_x elements of `dtype`; b: size_b elements of `dtype` or NULL (size_b = 0).
 * step_b = stride (in elements) of the bias dimension, as bias_act.cpp:75. */
int nbe_bias_act(const void* x, const void* b, void* y, int64_t size_x, int64_t size_b, int64_t step_b,
                 int act, float alpha, float gain, float clamp, int dtype, nbe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * upfirdn2d  --  pad, zero-insert upsample, 2-D FIR, decimate (closed form in SURVEY.md appendix C.2).
 * Replaces: upfirdn2d_plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)
 *           SG2/torch_utils/ops/upfirdn2d.cpp:16-94, kernels upfirdn2d.cu:29-200.
 * x: [N,C,H,W] with element strides xs_*; y: [N,C,OH,OW] with element strides ys_*, where
 * OH = (H*upy + pady0 + pady1 - fh + downy) / downy (same for W); f: fh x fw float32, dense, on device. */
int nbe_upfirdn2d(const void* x, const float* f, void* y,
                  int N, int C, int H, int W, int64_t xs_n, int64_t xs_c, int64_t xs_h, int64_t xs_w,
                  int OH, int OW, int64_t ys_n, int64_t ys_c, int64_t ys_h, int64_t ys_w,
                  int fh, int fw, int upx, int upy, int downx, int downy,
                  int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                  int dtype, nbe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Direct convolution, true FP32 (no TF32), NCHW -- the FP32-mode core of modulated_conv2d.
 *   y[n,o,oy,ox] = post( dcoef[n,o] * sum_{i,kh,kw} w[o,i,kh',kw'] * (x[n,g*cin_g+i, oy*stride+kh-pad, ox*stride+kw-pad] * xscale[n,i]) )
 * with (kh',kw') = (kh,kw) when flip == 0 (correlation, = F.conv2d) or (K-1-kh, K-1-kw) when flip == 1 (convolution);
 * out-of-range taps read 0.  post(v) = v + noise[n? ,oy,ox] * noise_gain, then optionally the bias_act epilogue.
 * Replaces: F.conv2d call sites SG2/torch_utils/ops/conv2d_gradfix.py:38 as used by conv2d_resample.py:29-54,144-147
 *           and the demodulation / noise glue of SG2/training/networks.py:66-76 (un-fused order).
 * xscale [N,Cin] / dcoef [N,Cout] / noise / bias may be NULL.  noise element (n,oy,ox) is at noise[n*noise_sn + oy*OW + ox]
 * (noise_sn = 0 broadcasts one map).  act == 0 skips the bias_act epilogue. */
int nbe_conv2d_f32(const float* x, const float* w, float* y,
                   int N, int Cin, int H, int W, int Cout, int K, int pad, int stride, int groups, int flip,
                   const float* xscale, const float* dcoef,
                   const float* noise, int64_t noise_sn, float noise_gain,
                   const float* bias, int act, float alpha, float gain, float clamp,
                   nbe_stream_t stream);

/* wsq[o,i] = sum_k w[o,i,k]^2   (plan-time constant for demodulation). */
int nbe_weight_sqsum_f32(const float* w, float* wsq, int Cout, int Cin, int KK, nbe_stream_t stream);

/* d[n,o] = rsqrt(sum_i styles[n,i]^2 * wsq[o,i] + 1e-8)      -- SG2/training/networks.py:59-64 */
int nbe_demod_coefs_f32(const float* styles, const float* wsq, float* d, int N, int Cin, int Cout, nbe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fully connected layer: y[n,o] = act( sum_i x[n,i] * w[o,i] * wgain + b[o] * bgain ) * act_gain
 * Replaces: FullyConnectedLayer.forward SG2/training/networks.py:109-122 (addmm / matmul + bias_act).
 * x may be float64 (x_is_f64 != 0: z arrives as float64, forger/ui/brush.py:669); output float32.
 * normalize != 0 first applies normalize_2nd_moment (networks.py:24-26) to each row of x. */
int nbe_fc_f32(const void* x, int x_is_f64, const float* w, const float* b, float* y, int N, int In, int Out,
               int64_t xs_n, int64_t ys_n, float wgain, float bgain, int act, float alpha, float act_gain,
               int normalize, nbe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Per-patch shifted constant noise.
 *   out[n,i,j] = bilinear(noise_const, row = frac(lin[j] + p1[n]) * (R-1), col = frac(lin[i] + p0[n]) * (R-1))
 *   with p = float32(positions[n] % mod) / float32(mod - 1), positions int64 [N,2] = (y, x).
 * Replaces: the grid_sample branch of SynthesisLayer.forward SG2/training/networks.py:371-382 together with
 *           Generator.forward_pre_mapped's normalisation networks_modified.py:351-353.  lin = torch.linspace(0,1,R). */
int nbe_shifted_noise_f32(const float* noise_const, const float* lin, const int64_t* positions, float* out,
                          int N, int R, int mod, nbe_stream_t stream);

/* All style affines and demodulation coefficients of one forward pass in one launch (the arrays are HOST arrays of
 * n_layers device pointers / ints):  styles_l[n,:] = affine_l(ws[n, w_index_l, :]) (FullyConnectedLayer, lr 1),
 * channels >= post_from_l multiplied by post_scale_l (ToRGB: 1/sqrt(C) after its 9 colour outputs, networks.py:455-462);
 * dcoef_l[n,o] = rsqrt(sum_i styles_l[n,i]^2 wsq_l[o,i] + 1e-8) from the UN-scaled styles when wsq_l / dcoef_l are non-NULL.
 * Replaces 12 x nbe_fc_f32 + 11 x nbe_demod_coefs_f32. */
int nbe_styles_demod_f32(const float* ws, int N, int num_ws, int w_dim, int n_layers,
                         const void* const* affine_w, const void* const* affine_b, const void* const* wsq,
                         void* const* styles, void* const* dcoef, const int* cin, const int* cout, const int* w_index,
                         const float* post_scale, const int* post_from, nbe_stream_t stream);

/* nbe_styles_demod_f32 that also writes the modulated constant input of the first synthesis block in the same launch
 * (in_layer >= 0): in_out[n, y, x, c] = bf16(in_const[y, x, c] * styles_{in_layer}[n, c]) for y < in_h, x < in_w, with in_const
 * float32 [in_h, in_w, cin] and in_out bf16 [N, in_h, in_pitch, cin] (columns >= in_w -- the zero gap of the flat layout -- are not
 * touched).  Replaces: `x = self.const ... repeat` (SG2/training/networks.py:642-643) followed by the `x * styles` of the
 * first modulated_conv2d (SG2/training/networks.py:68).  in_layer = -1: identical to nbe_styles_demod_f32. */
int nbe_styles_demod_input_f32(const float* ws, int N, int num_ws, int w_dim, int n_layers,
                               const void* const* affine_w, const void* const* affine_b, const void* const* wsq,
                               void* const* styles, void* const* dcoef, const int* cin, const int* cout, const int* w_index,
                               const float* post_scale, const int* post_from,
                               int in_layer, const float* in_const, void* in_out, int in_h, int in_w, int in_pitch,
                               nbe_stream_t stream);

/* nbe_shifted_noise_f32 for n_layers noise buffers at once (HOST arrays of device pointers; out_l is [N, res_l, res_l]). */
int nbe_shifted_noise_all_f32(const int64_t* positions, int N, int mod, int n_layers,
                              const void* const* noise_const, const void* const* lin, void* const* out, const int* res,
                              nbe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core (BF16, tcgen05 + TMA + TMEM) path.  Activations are NHWC bf16 with a channel stride
 * (so geometry features can live in the same buffer: SG2/training/networks_modified.py:219 torch.cat).
 */

/* NCHW float32 -> NHWC bf16 (dst[n,h,w,c_off + c], pixel stride dst_cs elements), optional per-(n,c) scale. */
int nbe_pack_nhwc_bf16(const float* src, void* dst, int N, int C, int H, int W, int dst_cs, int c_off,
                       const float* scale, nbe_stream_t stream);
/* NHWC bf16 (channel stride src_cs, first C channels) -> NCHW float32 */
int nbe_unpack_nchw_f32(const void* src, float* dst, int N, int C, int H, int W, int src_cs, nbe_stream_t stream);

/* U[n, 0..2H+1, 0..2W+1, c] = 4 * FIR4x4( zero-stuff( x[n,:,:,c] * scale[n,c] ) ) with padding (3,2,3,2):
 * the up=2 branch of conv2d_resample written FIR-first (= the reference's own generic fallback,
 * SG2/torch_utils/ops/conv2d_resample.py:149-154), so that the transposed conv becomes a *valid* 3x3
 * convolution over U.  x: NHWC bf16 [N,H,W,C] (pixel stride xs_c), U: NHWC bf16 [N,2H+2,2W+2,C] dense.
 * f: 4 separable taps are NOT assumed; f is the 4x4 float32 filter (setup_filter([1,3,3,1])).  scale may be NULL. */
int nbe_upsample2x_nhwc_bf16(const void* x, const float* f, const float* scale, void* u,
                             int N, int H, int W, int C, int xs_c, nbe_stream_t stream);


/* Same with an input row pitch in pixels (x is [N, H, x_pitch >= W, xs_c]; e.g. the zero-gapped buffers of the flat path). */
int nbe_upsample2x_nhwc_bf16_ex(const void* x, const float* f, const float* scale, void* u,
                                int N, int H, int W, int C, int xs_c, int x_pitch, nbe_stream_t stream);

/* wq: bf16 weights re-laid out as [K*K][Cout][Cin_pad] (tap-major, Cin_pad = Cin rounded up to 64, zero padded,
 * already flipped if the layer is a true convolution) -- produced by nbe_prepare_weights_bf16. */
int nbe_prepare_weights_bf16(const float* w, void* wq, int Cout, int Cin, int K, int flip, nbe_stream_t stream);

/* 3x3 (or 1x1) stride-1 convolution as an implicit GEMM on tcgen05 tensor cores:
 *   M = N*OH*OW pixels (tiles of 128), N = Cout (128), K = KK * Cin_pad; A tiles come straight from the NHWC
 *   activation tensor through a 4-D TMA box per filter tap (out-of-bounds = zero padding), B tiles from wq,
 *   accumulators live in TMEM, and the epilogue fuses demodulation, noise, bias, leaky-ReLU, gain, clamp and the
 *   NEXT layer's modulation:   y = clamp(lrelu(acc * dcoef[n,o] + noise*noise_gain + bias[o]) * gain, clamp) * next_scale[n,o]
 * x: NHWC bf16 [N, IH, IW, x_cs>=Cin]; "valid" selects IH = OH+2 (no padding, used after nbe_upsample2x) versus
 * IH = OH with zero padding 1.  y: NHWC bf16 [N,OH,OW,y_cs] written at channel offset 0.
 * Replaces: modulated_conv2d SG2/training/networks.py:31-88 + bias_act of SynthesisLayer.forward :386-390
 *           (cuDNN grouped conv, K5/K6 in SURVEY.md section 2.3). */
int nbe_conv_tc_bf16(const void* x, const void* wq, void* y,
                     int N, int OH, int OW, int Cin, int x_cs, int Cout, int y_cs, int K, int valid,
                     const float* dcoef, const float* noise, int64_t noise_sn, float noise_gain,
                     const float* bias, float alpha, float gain, float clamp, const float* next_scale,
                     nbe_stream_t stream);

/* Same kernel with (a) an input pixel stride of 1 or 2 (TMA traversal stride; needs valid = 1, the input then has
 * at least (O-1)*in_stride + K rows/cols; in_h/in_w give its real size, 0 = exactly that) and (b) explicit output pitches in pixels, so that the epilogue can write into the
 * interior of a larger (e.g. reflect-padded) NHWC buffer: element (n,oy,ox,c) goes to
 * y[(n*y_img_pitch + oy*y_row_pitch + ox)*y_cs + c].  Used by the geometry encoder
 * (forger/experimental/autoenc/simple_autoencoder.py:95-109,155-199: conv3x3 stride 1/2 + folded BN + LeakyReLU). */
int nbe_conv_tc_bf16_ex(const void* x, const void* wq, void* y,
                        int N, int OH, int OW, int Cin, int x_cs, int Cout, int y_cs, int K, int valid,
                        int in_stride, int in_h, int in_w, int64_t y_row_pitch, int64_t y_img_pitch,
                        const float* dcoef, const float* noise, int64_t noise_sn, float noise_gain,
                        const float* bias, float alpha, float gain, float clamp, const float* next_scale,
                        nbe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * "Flat" tensor-core convolutions: activations are NHWC bf16 whose rows are stored with a pitch x_pitch >= W + 1 and ZERO
 * gap columns, so that every filter tap is a constant row shift of the [positions x channels] matrix (see csrc/conv_flat.cu).
 * wq as produced by nbe_prepare_weights_bf16; the kernels work on 128 output channels at a time.
 */

/* 3x3 stride-1 modulated convolution + the fused epilogue of nbe_conv_tc_bf16.
 * valid = 0: x is [N, OH, x_pitch >= OW+1, x_cs] (zero gap columns supply the padding);
 * valid = 1: x is [N, OH+2, x_pitch >= OW+2, x_cs] and already contains the 1-pixel halo (e.g. the FIR-upsampled U).
 * y element (n,oy,ox,c) -> y[(n*y_img_pitch + oy*y_row_pitch + ox)*y_cs + c]; gap columns of y are never written.
 * Cout: a multiple of 128 (one pass per 128 output channels; dcoef / next_scale are [N, Cout]). */
int nbe_conv3x3_flat_bf16(const void* x, const void* wq, void* y,
                          int N, int OH, int OW, int Cin, int x_cs, int x_pitch, int valid, int Cout, int y_cs,
                          int64_t y_row_pitch, int64_t y_img_pitch,
                          const float* dcoef, const float* noise, int64_t noise_sn, float noise_gain,
                          const float* bias, float alpha, float gain, float clamp, const float* next_scale,
                          nbe_stream_t stream);

/* 3x3 stride-2 convolution (+ bias, leaky ReLU, gain, clamp, next_scale) at its algorithmic cost: the geometry encoder's
 * down-sampling layers (simple_autoencoder.py:46-70, Conv2d(k=3, stride=2, padding=1, padding_mode='reflect') + BN folded
 * into wq / bias + LeakyReLU).  xp: [N, H+2, W+2, Cin] NHWC bf16, dense channels, ALREADY carrying its 1-pixel border
 * (nbe_reflect_border_nhwc_bf16); H, W even; Cin % 64 == 0; Cout % 128 == 0 (one pass per 128 output channels).
 *   y[n, Y, X, o] = post( sum_{kh,kw,i} wq[kh*3+kw][o][i] * xp[n, 2Y+kh, 2X+kw, i] + bias[o] ),   Y < H/2, X < W/2
 * The input is read as its four parity planes xp[2Y'+a, 2X'+b] through one 5-D tensor map; tap (kh, kw) is plane
 * (kh & 1, kw & 1) shifted by (kh >> 1, kw >> 1) -- see csrc/conv_flat.cu. */
int nbe_conv3x3s2_flat_bf16(const void* xp, const void* wq, void* y,
                            int N, int H, int W, int Cin, int Cout, int y_cs, int64_t y_row_pitch, int64_t y_img_pitch,
                            const float* bias, float alpha, float gain, float clamp, const float* next_scale,
                            nbe_stream_t stream);

/* Transposed 3x3 convolution, stride 2, at its algorithmic cost (9 taps per INPUT pixel):
 *   T[n, 2Y+kh, 2X+kw, o] += dcoef[n,o] * sum_i wq[kh*3+kw][o][i] * x[n,Y,X,i]     -> T is (2H+1) x (2W+1)
 * = F.conv_transpose2d(x, w, stride=2) of the up-sampling path SG2/torch_utils/ops/conv2d_resample.py:124-138
 * (wq prepared with flip = 0).  x: [N, H, x_pitch >= W+1, x_cs] with zero gap columns (already modulated).
 * The four output parity classes are four TMEM accumulators of one pass over x. */
/* Cout: a multiple of 128 (one pass per 128 output channels). */
int nbe_convT3x3s2_flat_bf16(const void* x, const void* wq, void* t_out,
                             int N, int H, int W, int Cin, int x_cs, int x_pitch, int Cout, int t_cs,
                             int64_t t_row_pitch, int64_t t_img_pitch, const float* dcoef, nbe_stream_t stream);

/* 4x4 FIR (pad `pad`, filter gain fgain) followed by the SynthesisLayer epilogue, NHWC bf16 -> NHWC bf16:
 *   y = clamp(lrelu(fgain * FIR(T) * scale[n,c] + noise * noise_gain + bias[c]) * gain) * next_scale[n,c]
 * T: [N, TH, TW] valid extent (zero outside) with pitches; OH = TH + 2*pad - 3.  Replaces upfirdn2d after the transposed
 * conv (conv2d_resample.py:139) + fma/bias_act (SG2/training/networks.py:71-75, 386-390). */
int nbe_fir_act_nhwc_bf16(const void* t, const float* f, void* y, int N, int OH, int OW, int C, int TH, int TW, int pad,
                          int t_cs, int64_t t_row_pitch, int64_t t_img_pitch,
                          int y_cs, int64_t y_row_pitch, int64_t y_img_pitch, float fgain,
                          const float* scale, const float* noise, int64_t noise_sn, float noise_gain, const float* bias,
                          float alpha, float gain, float clamp, const float* next_scale, nbe_stream_t stream);

/* A whole up-sampling SynthesisLayer in ONE kernel: nbe_convT3x3s2_flat_bf16 followed by nbe_fir_act_nhwc_bf16 (pad 1) without
 * the T tensor ever reaching HBM -- the transposed conv runs on tcgen05 CTA pairs, its bf16 result goes into a per-CTA ring of the
 * last 16 T rows (`scratch`, L2-resident), and FIR warps of the same CTA filter the rows as they complete and apply
 *   y = clamp(lrelu(fgain * FIR(T) * dcoef[n,c] + noise * noise_gain + bias[c]) * gain) * next_scale[n,c]      -> [N, 2H, 2W, Cout]
 * Replaces: conv2d_resample's up-sampling path SG2/torch_utils/ops/conv2d_resample.py:124-142 (conv_transpose2d + upfirdn2d) and
 * the SynthesisLayer epilogue SG2/training/networks.py:386-390.  Bit-identical to the two-kernel sequence.
 * x: [N, H, x_pitch = W + 1, x_cs] zero-gapped, already modulated; wq: nbe_prepare_weights_bf16(flip = 0); f: 4x4 float32, must be
 * separable (outer product).  Needs Cout == 128, Cin <= 128, W % 8 == 0, W <= 120, gain > 0, 0 <= alpha <= 1: NBE_EUNSUPPORTED
 * otherwise (callers then run the two kernels).  scratch: nbe_up_layer_fused_scratch_bytes(W) bytes, zero-initialised ONCE by the
 * caller (guard columns stay zero; everything else is rewritten), not shared between concurrently running launches. */
int64_t nbe_up_layer_fused_scratch_bytes(int W);
int nbe_up_layer_fused_bf16(const void* x, const void* wq, const float* f, void* y, void* scratch, int64_t scratch_bytes,
                            int N, int H, int W, int Cin, int x_cs, int x_pitch, int Cout,
                            int y_cs, int64_t y_row_pitch, int64_t y_img_pitch, float fgain,
                            const float* dcoef, const float* noise, int64_t noise_sn, float noise_gain,
                            const float* bias, float alpha, float gain, float clamp, const float* next_scale,
                            nbe_stream_t stream);

/* The last synthesis layer with ToRGB fused into its epilogue (SynthesisBlock.forward networks.py:663-672 for the last
 * block): the 3x3 modulated conv of nbe_conv_tc_bf16 followed, per pixel and still in registers, by the 1x1 modulated
 * ToRGB (rgb_w [3,Cout] * rgb_styles [N,Cout], no demodulation), bias, clamp, softmax over the three UVS logits and the
 * triad colour mix (rgb_colors [N,3,3]) -> img / uvs [N,3,OH,OW] float32.  write_y = 0 skips storing the Cout-channel
 * feature map altogether (y may then be NULL).  Needs a 128-wide layer (OW % 128 == 0, OH even, Cin <= 128, Cout == 128);
 * returns NBE_EUNSUPPORTED otherwise (callers then run nbe_conv_tc_bf16 + nbe_torgb_triad). */
int nbe_conv_tc_bf16_torgb(const void* x, const void* wq, void* y,
                           int N, int OH, int OW, int Cin, int x_cs, int Cout, int y_cs, int valid,
                           const float* dcoef, const float* noise, int64_t noise_sn, float noise_gain,
                           const float* bias, float alpha, float gain, float clamp,
                           const float* rgb_w, const float* rgb_styles, const float* rgb_bias, const float* rgb_colors,
                           float rgb_clamp, float* img, float* uvs, int write_y, nbe_stream_t stream);

/* Fused ToRGB (1x1 modulated conv, no demodulation) + bias + clamp + softmax(3) + triad colour mix.
 *   t[k] = clamp(sum_c x[c] * w[k,c] * styles[n,c] + bias[k], +-clamp);  uvs = softmax(t);  img[c] = sum_k uvs[k]*colors[n,c,k]
 * Replaces: ToRGBColorTriadLayer.forward SG2/training/networks.py:451-485.
 * x: NHWC bf16 (x_is_bf16 = 1, pixel stride x_cs) or NCHW float32 (x_is_bf16 = 0); img/uvs: NCHW float32 [N,3,H,W]
 * (either may be NULL). */
int nbe_torgb_triad(const void* x, int x_is_bf16, int x_cs, const float* w, const float* styles, const float* bias,
                    const float* colors, float clamp, float* img, float* uvs, int N, int C, int H, int W,
                    nbe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * modulated_conv2d behind the reference's operator signature -- ONE call, NCHW in, NCHW out, dtype of x.
 * Replaces: modulated_conv2d(x, weight, styles, noise, up, down=1, padding, resample_filter, demodulate, flip_weight,
 *           fused_modconv) SG2/training/networks.py:31-88 (both of its formulations; they agree to ~1e-6) and, with
 *           styles = NULL / demodulate = 0, conv2d_resample(x, w, f, up, padding, flip_weight)
 *           SG2/torch_utils/ops/conv2d_resample.py:59-154 (cuDNN call sites conv2d_gradfix.py:38,43).
 *   y[n,o] = d[n,o] * conv2d_resample(x[n] * styles[n,:], weight, f, up, padding, flip_weight)[o] + noise[n]
 *   d[n,o] = rsqrt(sum_{i,k} (weight[o,i,k] * styles[n,i])^2 + 1e-8) if demodulate else 1
 * x: [N,Cin,H,W] dense NCHW of `dtype`; weight [Cout,Cin,K,K], styles [N,Cin] (NULL = ones), resample_filter (4x4, needed
 * for up = 2) float32; noise float32 or NULL, element (n,oy,ox) at noise[n*noise_sn + oy*OW + ox] (noise_sn = 0 broadcasts
 * one map); y: [N,Cout,OH,OW] of `dtype`, OH = H + 2*padding - K + 1 (up = 1) or 2H + 2*padding - 2 (up = 2, K = 3).
 * flip_weight != 0 = correlation (F.conv2d), 0 = true convolution (what the up-sampling SynthesisLayers pass).
 * dtype NBE_BF16 / NBE_F16 with K = 3, Cout % 128 == 0, padding 1 (or 0 for up = 1): tensor cores -- pack to zero-gapped
 *   NHWC bf16 with the modulation applied, nbe_conv3x3_flat_bf16 (up = 1) or nbe_convT3x3s2_flat_bf16 +
 *   nbe_fir_act_nhwc_bf16 (up = 2, the transposed conv at 9 taps per input pixel), unpack; fp32 accumulation.
 * dtype NBE_F32: true-FP32 direct convolution (nbe_conv2d_f32; up = 2 FIR-first through nbe_upfirdn2d).
 * Other 16-bit shapes return NBE_EUNSUPPORTED (the caller converts to float32).
 * workspace: caller-owned, 256-byte aligned, at least nbe_modulated_conv2d_workspace(...) bytes (that function returns
 * -1 for an unsupported combination); nothing in it needs to be initialised or kept between calls. */
int64_t nbe_modulated_conv2d_workspace(int dtype, int N, int Cin, int H, int W, int Cout, int K, int up, int padding);
int nbe_modulated_conv2d(const void* x, int dtype, const float* weight, const float* styles,
                         const float* noise, int64_t noise_sn, void* y,
                         int N, int Cin, int H, int W, int Cout, int K, int up, int padding,
                         const float* resample_filter, int demodulate, int flip_weight,
                         void* workspace, int64_t workspace_bytes, nbe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Geometry encoder, tensor-core path (forger/experimental/autoenc/simple_autoencoder.py:95-126,155-199,251-261).
 * Activations are NHWC bf16 buffers that carry their 1-pixel reflect padding explicitly, [N, H+2, W+2, cs].
 */

/* First layer: conv7x7(1 -> Cout, reflect pad 3) + folded eval-BatchNorm bias + LeakyReLU(neg_slope) on x [N,1,H,W] float32
 * (preproc: 0 none, 1 'inverse', 2 '-11inverse', base.py:32-58); writes the interior of y [N,H+2,W+2,y_cs] bf16.
 * w: [Cout,49] float32 with BN folded. */
int nbe_enc_conv7x7_bf16(const float* x, const float* w, const float* bias, void* y, int N, int H, int W, int Cout,
                         int y_cs, float neg_slope, int preproc, nbe_stream_t stream);

/* The same layer as one K = 64 tensor-core GEMM per 128-pixel tile (49 taps + 15 zero columns; the im2col tile is written
 * directly in the swizzled UMMA smem layout).  wq: [Cout,64] bf16 = BN-folded weights, taps kh*7+kw in columns 0..48, zeros after. */
int nbe_enc_conv7x7_tc_bf16(const float* x, const void* wq, const float* bias, void* y, int N, int H, int W, int Cout,
                            int y_cs, float neg_slope, int preproc, nbe_stream_t stream);

/* The same layer as a BLOCK-TOEPLITZ implicit GEMM (csrc/enc7x7_toeplitz.cu): a GEMM row is a group of 8 adjacent pixels, so the
 * A operand is fetched by one overlapping 5-D TMA box per tile from a reflect-padded bf16 copy of the image (no im2col by
 * threads), N = 8 px x 64 channels, K = 7 x 16.  Writes the interior AND the 1-pixel reflect border of y [N,H+2,W+2,64] bf16
 * (no nbe_reflect_border_nhwc_bf16 pass needed).  Needs Cout == y_cs == 64, H % 8 == 0, W % 128 == 0 (NBE_EUNSUPPORTED otherwise:
 * use nbe_enc_conv7x7_tc_bf16).  wt: [7][512][16] bf16 from nbe_enc_conv7x7_toeplitz_weights(w [Cout,49] float32, BN folded);
 * scratch: nbe_enc_conv7x7_toeplitz_scratch_bytes(N, H, W) bytes (the padded image; contents need not be preserved). */
int64_t nbe_enc_conv7x7_toeplitz_scratch_bytes(int N, int H, int W);
int nbe_enc_conv7x7_toeplitz_weights(const float* w, void* wt, int Cout, nbe_stream_t stream);
int nbe_enc_conv7x7_toeplitz_bf16(const float* x, const void* wt, const float* bias, void* y, void* scratch, int64_t scratch_bytes,
                                  int N, int H, int W, int Cout, int y_cs, float neg_slope, int preproc, nbe_stream_t stream);

/* Fill the 1-pixel border of buf [N,Hp,Wp,cs] (first C channels) by reflection of its interior (torch padding_mode='reflect'). */
int nbe_reflect_border_nhwc_bf16(void* buf, int N, int Hp, int Wp, int C, int cs, nbe_stream_t stream);

/* out [N,2h+2,2w+2,out_cs] = reflect_pad1( bilinear_x2(x [N,h,w,xs_c], align_corners=True) )   (ScaleUp.up + conv padding) */
int nbe_bilinear2x_pad_nhwc_bf16(const void* x, void* out, int N, int h, int w, int C, int xs_c, int out_cs,
                                 nbe_stream_t stream);

/* Eval-mode BatchNorm AFTER the activation (the encoder's --neg_slope variant, simple_autoencoder.py:100-103,128-148), for the
 * feature maps that leave the encoder: y[n,py,px,c] = (x[n,py,px,c] * scale[c] + shift[c]) * next_scale[n,c] (next_scale may be
 * NULL) on NHWC bf16 with independent pitches (in pixels) and channel strides, C % 8 == 0; and the same on dense NCHW float32. */
int nbe_affine_nhwc_bf16(const void* x, int x_cs, int64_t x_row_pitch, int64_t x_img_pitch,
                         void* y, int y_cs, int64_t y_row_pitch, int64_t y_img_pitch,
                         int N, int H, int W, int C, const float* scale, const float* shift, const float* next_scale,
                         nbe_stream_t stream);
int nbe_affine_nchw_f32(const float* x, float* y, int N, int C, int HW, const float* scale, const float* shift, nbe_stream_t stream);

/* FP32 parity mode of the same two steps on NCHW float32: y = reflect_pad(x, pad) (padding_mode='reflect',
 * simple_autoencoder.py:98), or, with upsample2x != 0, y = reflect_pad(bilinear_x2(x), pad) with align_corners=True
 * (ScaleUp, simple_autoencoder.py:117) without materialising the up-sampled map.
 * x: [NC, H, W]; y: [NC, S*H + 2 pad, S*W + 2 pad], S = 2 if upsample2x else 1. */
int nbe_reflect_pad_nchw_f32(const float* x, float* y, int64_t NC, int H, int W, int pad, int upsample2x,
                             nbe_stream_t stream);

/* ToRGB in the 'canvas' colour format (ToRGBColorTriadLayer with color_format == 'canvas', SG2/training/networks.py:433-481):
 *   t[k] = clamp(sum_c x[c] * w[k,c] * styles[n,c] + bias[k], +-clamp), k < 8 ;  uvs = softmax(t[0:3]) ; canvas = t[3:6] ;
 *   alpha = softmax(t[6:8]) ;  img[c] = alpha[0] * sum_k uvs[k] * colors[n,c,k] + alpha[1] * canvas[c]
 * x as in nbe_torgb_triad; w [8,C]; img / uvs / canvas [N,3,H,W], alpha [N,2,H,W] float32 (each may be NULL). */
int nbe_torgb_canvas(const void* x, int x_is_bf16, int x_cs, const float* w, const float* styles, const float* bias,
                     const float* colors, float clamp, float* img, float* uvs, float* canvas, float* alpha,
                     int N, int C, int H, int W, nbe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Engine composite and canvas placement.
 */

/* rgba[n,:,y,x] in [0,1] from uvs [N,3,H,W] and colors01 [N,3,3] (C x ncolors):
 *   optional UVS "clear background" remap with per-patch sfactor[n] (NULL = off), rgb = sum_k uvs_k colors_k,
 *   alpha = U+V (mode 0, 'clear') or 1 (mode 1, 'full').
 * Replaces: TriadGanPaintEngine._render_stroke_torch tail forger/ui/brush.py:763-792 and
 *           StyleUVSMapper._map_style_s forger/ui/mapper.py:53-72.
 * out_f32: [N,4,H,W] float32 or NULL.  out_u8: [N,H-2m,W-2m,4] uint8 HWC tiles, trunc(clip(255*v)) as
 * PaintingHelper.render_stroke forger/ui/brush.py:369-377, or NULL. */
int nbe_triad_composite(const float* uvs, const float* colors01, const float* sfactor, int mode,
                        float* out_f32, uint8_t* out_u8, int N, int H, int W, int crop_margin, nbe_stream_t stream);

/* CanvasPaintEngine._render_stroke_torch tail (forger/ui/brush.py:905-935), outputs as nbe_triad_composite:
 *   stroke = sum_k uvs_k colors01_k ; cv = (gen_canvas + 1) / 2 ; a = alpha_fg[n * alpha_sn + pix]
 *   mode 0 'clear': (stroke, a) ; 1 'stroke': (stroke, 1) ; 2 'canvas': (cv, 1) ; 3 'full': ((1 - a) cv + a stroke, 1). */
int nbe_canvas_composite(const float* uvs, const float* colors01, const float* alpha_fg, int64_t alpha_sn,
                         const float* gen_canvas, int mode, float* out_f32, uint8_t* out_u8,
                         int N, int H, int W, int crop_margin, nbe_stream_t stream);

/* counts[r*ncols + c] = number of stroke pixels (== 0) of the P x P crop at (r*stride, c*stride) of the padded guidance
 * (pixels outside the canvas do not count): the "more than 10 stroke pixels" crop filter of
 * forger/viz/style_transfer.py:43-47 for stitching modes other than 'all'. */
int nbe_count_stroke_pixels(const uint8_t* canvas, int canvas_h, int canvas_w, int P, int stride, int nrows, int ncols,
                            int32_t* counts, nbe_stream_t stream);

/* geom[n,0,y,x] = 1 - (255 - canvas[(cy[n]+y)*canvas_w + cx[n]+x]) / 255  for 0 <= y,x < P  (float32, 0 = stroke):
 * the crop + `255 - geom` + prepare_geom_input chain of forger/viz/paint_image_main.py:164-167 and
 * forger/ui/brush.py:672-681.  canvas: padded guidance, uint8 [canvas_h, canvas_w], 0 = stroke. crops: int32 [N,2] = (y,x). */
int nbe_gather_geom_patches(const uint8_t* canvas, int canvas_h, int canvas_w, const int32_t* crops, float* geom,
                            int N, int P, nbe_stream_t stream);

/* Place uint8 RGBA tiles [N,T,T,4] at (ty[n], tx[n]) into canvas [canvas_h, canvas_w, 4] with the reference's
 * last-writer-wins raster semantics (forger/viz/paint_image_main.py:163-177): a tile pixel is written only if
 * owner[y*canvas_w+x] == order[n] where owner is the int32 map of the highest raster index covering each pixel
 * (computed once per canvas by nbe_tile_owner_map). */
int nbe_tile_owner_map(const int32_t* tile_yx, int n_tiles, int T, int32_t* owner, int canvas_h, int canvas_w,
                       nbe_stream_t stream);
int nbe_place_tiles(const uint8_t* tiles, const int32_t* tile_yx, const int32_t* order, int N, int T,
                    const int32_t* owner, uint8_t* canvas, int canvas_h, int canvas_w, nbe_stream_t stream);

/* Feature blending against a persistent feature canvas for a batch of patches with DISJOINT windows (one wavefront of the crop
 * grid), on the flat generator path: brush.py:190-242 (saved-feature lookup, dirty-area alpha, core write-back) and
 * stitching.py:18-25 (the blend) in one pass over the block output x [N, R, x_pitch, x_cs] bf16 (channels 0..C-1):
 *   m = fmask[fy+r, fx+c];  alpha = m ? base_alpha[r,c] : 1;  x_b = bf16((1 - alpha) * fcanvas[fy+r, fx+c, :] + alpha * x)
 *   if (base_alpha > 0.99 or (m and base_alpha > 0)) and crop_margin <= r, c < R - crop_margin:  fcanvas <- x_b, fmask <- 1
 *   x <- bf16(x_b * scale[n, :])   (scale = the consuming layer's styles, or NULL)
 * fcanvas: [FH, FW, C] bf16, fmask: [FH, FW] uint8, fyx: [N, 2] int32 window origins (windows must stay inside the canvas). */
int nbe_blend_window_nhwc_bf16(void* x, int x_pitch, int x_cs, int R, int C, void* fcanvas, uint8_t* fmask, int FH, int FW,
                               const int32_t* fyx, const float* base_alpha, int crop_margin, const float* scale, int N,
                               nbe_stream_t stream);

/* x = alpha * saved + (1 - alpha) * x on NCHW float32 or NHWC bf16 features (BlendedFeatures.blend,
 * forger/train/stitching.py:24-25), alpha [H,W] shared over channels, per patch n (alpha_sn = 0 broadcasts). */
int nbe_blend_features(void* x, const void* saved, const float* alpha, int64_t alpha_sn, int N, int C, int H, int W,
                       int is_nhwc_bf16, int cs, nbe_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NBE_B200_H_ */
