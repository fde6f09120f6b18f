// dbaz_nn_kernels.cuh -- fused elementwise stages of the leaf-evaluation pipeline.  The dense
// contractions (conv / linear) stay in cuDNN / cuBLAS; everything between them that PyTorch
// eager would run as 3-4 separate passes (bias add, ReLU, eval-mode BatchNorm, dtype casts,
// log-softmax + exp, tanh) is one HBM pass here.
//
// Reference semantics: dots_boxes/dots_boxes_nn.py:85-98 (x = bn(relu(conv(x)))), nn.py:49-58,
// nn.py:155-160 (p = exp(log_softmax), v = tanh).
#pragma once
#include "dbaz_device.cuh"

namespace dbaz {

template <typename T> struct Vec8;  // 8 consecutive channels
template <> struct Vec8<__nv_bfloat16> {
    uint4 raw;
    __device__ __forceinline__ void unpack(float* f) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    __device__ __forceinline__ void pack(const float* f) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    }
};
template <> struct Vec8<__half> {
    uint4 raw;
    __device__ __forceinline__ void unpack(float* f) const {
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    __device__ __forceinline__ void pack(const float* f) {
        __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    }
};

// mode 0:  y = scale[c] * relu(x + bias[c]) + shift[c]            (SimpleNN: conv -> ReLU -> BN)
// mode 1:  y = relu(scale[c] * (x + bias[c]) + shift[c] + res)    (ResNetZero: conv -> BN -> (+res) -> ReLU), res may be null
// mode 2:  y = scale[c] * (x + bias[c]) + shift[c]                (affine only)
// x is [rows, C] with the channel innermost (NHWC conv output / linear output), in place.
//
// HBM-bound: one 16-byte vector (VW channels) per thread per iteration.  The host sizes the grid so
// that the total thread count is a multiple of C/VW: a thread then always sees the same VW channels and
// keeps their bias/scale/shift in registers for the whole grid-stride loop (the per-element parameter
// loads and the 64-bit modulo of a naive version made it LSU-bound at 1.5 TB/s).  UNROLL independent
// vectors are in flight per thread.
template <int MODE>
__device__ __forceinline__ float epi_apply(float x, float b, float s, float t, float r, bool has_res) {
    float y = x + b;
    if (MODE == 0) return fmaf(s, fmaxf(y, 0.0f), t);
    if (MODE == 1) { y = fmaf(s, y, t); if (has_res) y += r; return fmaxf(y, 0.0f); }
    return fmaf(s, y, t);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_nn_epilogue16(T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ bias,
                const float* __restrict__ scale, const float* __restrict__ shift, uint32_t n_vec, int C) {
    constexpr int UNROLL = 4;
    const uint32_t nthreads = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = (int)(tid % (uint32_t)(C >> 3)) << 3;
    float b[8], s[8], t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { b[j] = bias ? bias[c0 + j] : 0.0f; s[j] = scale[c0 + j]; t[j] = shift[c0 + j]; }
    uint4* xv = reinterpret_cast<uint4*>(x);
    const uint4* rv = reinterpret_cast<const uint4*>(res);
    for (uint32_t base = tid; base < n_vec; base += nthreads * UNROLL) {
        Vec8<T> v[UNROLL], r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint32_t i = base + u * nthreads;
            if (i < n_vec) { v[u].raw = xv[i]; if (res) r[u].raw = rv[i]; }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint32_t i = base + u * nthreads;
            if (i < n_vec) {
                float f[8], g[8];
                v[u].unpack(f);
                if (res) r[u].unpack(g);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = epi_apply<MODE>(f[j], b[j], s[j], t[j], res ? g[j] : 0.0f, res != nullptr);
                v[u].pack(f);
                xv[i] = v[u].raw;
            }
        }
    }
}

// fp32 variant (4 channels per 16-byte vector)
template <int MODE>
__global__ void __launch_bounds__(256)
k_nn_epilogue32(float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ bias,
                const float* __restrict__ scale, const float* __restrict__ shift, uint32_t n_vec, int C) {
    constexpr int UNROLL = 4;
    const uint32_t nthreads = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = (int)(tid % (uint32_t)(C >> 2)) << 2;
    float b[4], s[4], t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { b[j] = bias ? bias[c0 + j] : 0.0f; s[j] = scale[c0 + j]; t[j] = shift[c0 + j]; }
    float4* xv = reinterpret_cast<float4*>(x);
    const float4* rv = reinterpret_cast<const float4*>(res);
    for (uint32_t base = tid; base < n_vec; base += nthreads * UNROLL) {
        float4 v[UNROLL], r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint32_t i = base + u * nthreads;
            r[u] = make_float4(0, 0, 0, 0);
            if (i < n_vec) { v[u] = xv[i]; if (res) r[u] = rv[i]; }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint32_t i = base + u * nthreads;
            if (i < n_vec) {
                float4 o;
                o.x = epi_apply<MODE>(v[u].x, b[0], s[0], t[0], r[u].x, res != nullptr);
                o.y = epi_apply<MODE>(v[u].y, b[1], s[1], t[1], r[u].y, res != nullptr);
                o.z = epi_apply<MODE>(v[u].z, b[2], s[2], t[2], r[u].z, res != nullptr);
                o.w = epi_apply<MODE>(v[u].w, b[3], s[3], t[3], r[u].w, res != nullptr);
                xv[i] = o;
            }
        }
    }
}

// Heads: logits [n, ld] (policy logits in columns 0..A-1, value pre-activation in column A) ->
// priors float32 [n, A] = softmax (== exp(log_softmax), nn.py:159), values float32 [n] = tanh.
// One warp per row.
template <typename T>
__global__ void k_nn_heads(const T* __restrict__ logits, int ld, int A, float* __restrict__ priors,
                           float* __restrict__ values, int64_t n) {
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const T* row = logits + w * ld;
    float x[4];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int a = lane + 32 * k;
        x[k] = a < A ? (float)row[a] : -INFINITY;
        m = fmaxf(m, x[k]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { x[k] = (lane + 32 * k < A) ? __expf(x[k] - m) : 0.0f; s += x[k]; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float inv = 1.0f / s;
#pragma unroll
    for (int k = 0; k < 4; ++k) { int a = lane + 32 * k; if (a < A) priors[w * A + a] = x[k] * inv; }
    if (lane == 0) values[w] = tanhf((float)row[A]);
}

}  // namespace dbaz
