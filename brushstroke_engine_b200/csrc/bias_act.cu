// bias_act forward: y = clamp(act(x + b) * gain).  HBM-bound: 2 * numel * sizeof(T) algorithmic bytes.
//
// Design (B200): 16-byte vector loads/stores that bypass L1, four independent vectors in flight per
// thread (the whole tensor is touched once, so latency is hidden by MLP, not by reuse), one 64-bit
// division per *vector* for the bias index (all lanes of a vector share a bias element whenever
// step_b % VEC == 0, which holds for every NCHW call the generator makes: step_b = H*W), grid sized
// as a multiple of the 148 SMs with a grid-stride loop.
#include "common.cuh"

namespace nbe {

template <class T, int ACT>
__device__ __forceinline__ T bias_act_one(T xv, float b, float alpha, float gain, float clamp) {
    float v = Cvt<T>::ld(xv) + b;
    v = apply_act(v, ACT, alpha) * gain;
    if (clamp >= 0.f) v = fminf(fmaxf(v, -clamp), clamp);
    return Cvt<T>::st(v);
}

template <class T, int ACT, bool UNIFORM_BIAS>
__global__ void __launch_bounds__(256)
bias_act_vec_kernel(const T* __restrict__ x, const T* __restrict__ b, T* __restrict__ y,
                    int64_t n_vec, int64_t size_b, int64_t step_b, float alpha, float gain, float clamp) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int UNROLL = 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < n_vec; v0 += stride * UNROLL) {
        int4 raw[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            int64_t v = v0 + u * stride;
            if (v < n_vec) raw[u] = ld_stream16(reinterpret_cast<const int4*>(x) + v);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            int64_t v = v0 + u * stride;
            if (v >= n_vec) continue;
            T* e = reinterpret_cast<T*>(&raw[u]);
            if (UNIFORM_BIAS) {
                float bb = 0.f;
                if (b) bb = Cvt<T>::ld(b[((v * VEC) / step_b) % size_b]);
#pragma unroll
                for (int k = 0; k < VEC; ++k) e[k] = bias_act_one<T, ACT>(e[k], bb, alpha, gain, clamp);
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    float bb = Cvt<T>::ld(b[((v * VEC + k) / step_b) % size_b]);
                    e[k] = bias_act_one<T, ACT>(e[k], bb, alpha, gain, clamp);
                }
            }
            st_stream16(reinterpret_cast<int4*>(y) + v, raw[u]);
        }
    }
}

// Scalar tail / unaligned / float64 path.
template <class T, class AccT>
__global__ void bias_act_scalar_kernel(const T* __restrict__ x, const T* __restrict__ b, T* __restrict__ y,
                                       int64_t begin, int64_t size_x, int64_t size_b, int64_t step_b,
                                       int act, float alpha, float gain, float clamp) {
    int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size_x) return;
    AccT v = (AccT)x[i];
    if (b) v += (AccT)b[(i / step_b) % size_b];
    if (sizeof(AccT) == 8) {
        double d = (double)v, a = alpha;
        switch (act) {
            case NBE_ACT_RELU: d = d > 0 ? d : 0; break;
            case NBE_ACT_LRELU: d = d > 0 ? d : d * a; break;
            case NBE_ACT_TANH: d = tanh(d); break;
            case NBE_ACT_SIGMOID: d = 1.0 / (1.0 + exp(-d)); break;
            case NBE_ACT_ELU: d = d >= 0 ? d : expm1(d); break;
            case NBE_ACT_SELU: d = 1.0507009873554804934193349852946 * (d >= 0 ? d : 1.6732632423543772848170429916717 * expm1(d)); break;
            case NBE_ACT_SOFTPLUS: d = d > 20 ? d : log1p(exp(d)); break;
            case NBE_ACT_SWISH: d = d / (1.0 + exp(-d)); break;
            default: break;
        }
        d *= (double)gain;
        if (clamp >= 0.f) d = fmin(fmax(d, -(double)clamp), (double)clamp);
        y[i] = (T)d;
    } else {
        float f = apply_act((float)v, act, alpha) * gain;
        if (clamp >= 0.f) f = fminf(fmaxf(f, -clamp), clamp);
        y[i] = (T)f;
    }
}

template <class T, int ACT>
static void launch_vec(const T* x, const T* b, T* y, int64_t n_vec, int64_t size_b, int64_t step_b,
                       float alpha, float gain, float clamp, cudaStream_t s) {
    constexpr int VEC = 16 / sizeof(T);
    int64_t blocks = (n_vec + 256 * 4 - 1) / (256 * 4);
    const int64_t cap = (int64_t)kNumSMs * 16;                 // 16 x 256 threads = full residency per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const bool uniform = (b == nullptr) || (step_b % VEC == 0);
    if (uniform)
        bias_act_vec_kernel<T, ACT, true><<<(int)blocks, 256, 0, s>>>(x, b, y, n_vec, size_b, step_b, alpha, gain, clamp);
    else
        bias_act_vec_kernel<T, ACT, false><<<(int)blocks, 256, 0, s>>>(x, b, y, n_vec, size_b, step_b, alpha, gain, clamp);
}

template <class T>
static int run_typed(const void* xv, const void* bv, void* yv, int64_t size_x, int64_t size_b, int64_t step_b,
                     int act, float alpha, float gain, float clamp, cudaStream_t s) {
    const T* x = (const T*)xv; const T* b = (const T*)bv; T* y = (T*)yv;
    constexpr int VEC = 16 / sizeof(T);
    int64_t n_vec = 0;
    const bool aligned = (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
    if (aligned) n_vec = size_x / VEC;
    if (n_vec > 0) {
        switch (act) {
#define NBE_CASE(A) case A: launch_vec<T, A>(x, b, y, n_vec, size_b, step_b, alpha, gain, clamp, s); break;
            NBE_CASE(NBE_ACT_LINEAR) NBE_CASE(NBE_ACT_RELU) NBE_CASE(NBE_ACT_LRELU) NBE_CASE(NBE_ACT_TANH)
            NBE_CASE(NBE_ACT_SIGMOID) NBE_CASE(NBE_ACT_ELU) NBE_CASE(NBE_ACT_SELU) NBE_CASE(NBE_ACT_SOFTPLUS)
            NBE_CASE(NBE_ACT_SWISH)
#undef NBE_CASE
            default: return fail(NBE_EINVAL, "bias_act: unknown activation %d", act);
        }
        int st = launched("bias_act_vec_kernel");
        if (st) return st;
    }
    const int64_t begin = n_vec * VEC;
    if (begin < size_x) {
        int64_t rem = size_x - begin;
        bias_act_scalar_kernel<T, float><<<(int)((rem + 255) / 256), 256, 0, s>>>(
            x, b, y, begin, size_x, size_b, step_b, act, alpha, gain, clamp);
        return launched("bias_act_scalar_kernel");
    }
    return NBE_OK;
}

}  // namespace nbe

extern "C" int nbe_bias_act(const void* x, const void* b, void* y, int64_t size_x, int64_t size_b, int64_t step_b,
                            int act, float alpha, float gain, float clamp, int dtype, nbe_stream_t stream) {
    using namespace nbe;
    NBE_REQUIRE(size_x >= 0, "bias_act: negative size");
    NBE_REQUIRE(size_x <= INT32_MAX, "bias_act: x is too large");           // bias_act.cpp:40
    if (size_x == 0) return NBE_OK;
    NBE_REQUIRE(x && y, "bias_act: null tensor");
    NBE_REQUIRE(act >= NBE_ACT_LINEAR && act <= NBE_ACT_SWISH, "bias_act: no CUDA kernel found for the specified activation func");
    if (size_b == 0) b = nullptr;
    if (b) NBE_REQUIRE(step_b >= 1, "bias_act: bad bias step");
    if (!b) { size_b = 1; step_b = 1; }
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case NBE_F32:  return run_typed<float>(x, b, y, size_x, size_b, step_b, act, alpha, gain, clamp, s);
        case NBE_F16:  return run_typed<__half>(x, b, y, size_x, size_b, step_b, act, alpha, gain, clamp, s);
        case NBE_BF16: return run_typed<__nv_bfloat16>(x, b, y, size_x, size_b, step_b, act, alpha, gain, clamp, s);
        case NBE_F64: {
            bias_act_scalar_kernel<double, double><<<(int)((size_x + 255) / 256), 256, 0, s>>>(
                (const double*)x, (const double*)b, (double*)y, 0, size_x, size_b, step_b, act, alpha, gain, clamp);
            return launched("bias_act_scalar_kernel<double>");
        }
        default: return fail(NBE_EINVAL, "bias_act: unsupported dtype %d", dtype);
    }
}
