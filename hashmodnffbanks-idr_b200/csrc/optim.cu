// K6: optimiser step on the flat parameter / gradient bucket: global-norm clip (max_norm) fused with Adam.
// Replaces, for the data-parallel trainer, the caller-side sequence of idr_train.py:299-308
//   clip_grad_norm_(model.parameters(), 1.0)  +  torch.optim.Adam(lr).step()
// One pass reads g, m, v, p and writes m, v, p (HBM-bound: 28 B / parameter); the squared norm is
// produced by a separate reduction pass (4 B / parameter) and consumed from device memory, so the
// step needs no host synchronisation.  `grad_scale` folds the 1/world_size of the gradient all-reduce.
#include "common.cuh"

namespace idrk {

__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long k = i; k < n4; k += stride) {
        const float4 v = g4[k];
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    for (long long k = (n4 << 2) + i; k < n; k += stride) acc = fmaf(g[k], g[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) atomicAdd(out, v);
    }
}

// Order-independent variant for data-parallel replicas: every rank must derive the SAME clip factor from the (bitwise
// identical) all-reduced bucket, or the replicas drift apart in the last bits and never re-converge.  Atomics into one
// float do not give that; here CTA b writes its partial to partials[b] (fixed grid, fixed shuffle tree) and ONE block adds
// the partials in a fixed order.
__global__ void sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partials) {
    pdl_wait();
    pdl_trigger();
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long k = i; k < n4; k += stride) {
        const float4 v = g4[k];
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    for (long long k = (n4 << 2) + i; k < n; k += stride) acc = fmaf(g[k], g[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partials[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(1024)
sumsq_final_kernel(const float* __restrict__ partials, int n_partials, float* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float part[32];
    float acc = 0.f;
    for (int k = threadIdx.x; k < n_partials; k += 1024) acc += partials[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = part[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) out[0] = v;
    }
}

__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                                 float max_norm, const float* __restrict__ sumsq, float grad_scale) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    float coef = grad_scale;
    if (max_norm > 0.f) {
        const float total = sqrtf(*sumsq) * grad_scale;
        const float c = max_norm / (total + 1e-6f);
        coef *= c < 1.f ? c : 1.f;
    }
    const float step = lr / bc1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i] * coef;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= step * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

}  // namespace idrk

using namespace idrk;

extern "C" int idrk_sumsq(const float* g, int64_t n, float* out, void* stream) {
    if (!g || !out || n < 0) return IDRK_E_ARG;
    if (n == 0) return 0;
    if (!aligned16(g)) return IDRK_E_ALIGN;
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    IDRK_CUDA_TRY(launch_k(sumsq_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, g, n, out));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_sumsq_det(const float* g, int64_t n, float* out, float* partials, int32_t n_partials, void* stream) {
    if (!g || !out || !partials || n < 0 || n_partials < 1 || n_partials > 65536) return IDRK_E_ARG;
    if (!aligned16(g)) return IDRK_E_ALIGN;
    IDRK_CUDA_TRY(launch_k(sumsq_partial_kernel, dim3(n_partials), dim3(256), 0, (cudaStream_t)stream, g, (long long)n, partials));
    IDRK_CUDA_TRY(launch_k(sumsq_final_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, (const float*)partials, (int)n_partials, out));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                              float eps, int32_t step, float max_norm, const float* sumsq, float grad_scale, void* stream) {
    if (!p || !g || !m || !v || n < 0 || step < 1 || (max_norm > 0.f && !sumsq)) return IDRK_E_ARG;
    if (n == 0) return 0;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2 = 1.f - powf(beta2, (float)step);
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    IDRK_CUDA_TRY(launch_k(clip_adam_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2), max_norm,
                                                                   sumsq, grad_scale));
    IDRK_LAUNCH_CHECK();
    return 0;
}
