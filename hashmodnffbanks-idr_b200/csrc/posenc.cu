// NeRF-style positional encoding, forward and backward (memory-bound elementwise kernels).
// Replaces PositionalEncoding.embed (model/embeddings/frequency_enc.py:19-51): with include_input the
// row is [x | x | sin(f0 x) | cos(f0 x) | sin(f1 x) | ...] - the input appears twice, as in the reference.
#include "common.cuh"

namespace idrk {

struct Bands { float f[32]; int n; };

__global__ void posenc_fwd_kernel(const float* __restrict__ x, long long n, int d, int ldx, Bands b, int include_input,
                                  float* __restrict__ out, int ld_out, int width) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const long long total = n * (long long)ld_out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / ld_out;
        int c = (int)(i - p * ld_out);
        float v = 0.f;
        if (c < width) {
            const int head = include_input ? 2 * d : 0;
            if (c < head) {
                v = x[p * ldx + (c % d)];
            } else {
                c -= head;
                const int k = c / d, j = c - k * d;
                const float a = __fmul_rn(x[p * ldx + j], b.f[k >> 1]);
                v = (k & 1) ? cosf(a) : sinf(a);
            }
        }
        out[i] = v;
    }
}

__global__ void posenc_bwd_kernel(const float* __restrict__ x, long long n, int d, int ldx, Bands b, int include_input,
                                  const float* __restrict__ dy, int ld_dy, float* __restrict__ dx, int ld_dx) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const long long total = n * (long long)d;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / d;
        const int j = (int)(i - p * d);
        const float* g = dy + p * ld_dy;
        const float xv = x[p * ldx + j];
        float acc = 0.f;
        int base = 0;
        if (include_input) { acc = g[j] + g[d + j]; base = 2 * d; }
        for (int q = 0; q < b.n; ++q) {
            const float f = b.f[q];
            float sn, cs;
            sincosf(__fmul_rn(xv, f), &sn, &cs);
            acc = fmaf(g[base + (2 * q) * d + j] * cs - g[base + (2 * q + 1) * d + j] * sn, f, acc);
        }
        dx[p * ld_dx + j] = acc;
    }
}

}  // namespace idrk

using namespace idrk;

extern "C" int idrk_posenc_fwd(const float* x, int64_t n, int32_t d, int32_t ldx, const float* h_bands, int32_t n_bands,
                               int32_t include_input, float* out, int32_t ld_out, void* stream) {
    if (!x || !out || n < 0 || d < 1 || ldx < d || n_bands < 0 || n_bands > 32 || (n_bands > 0 && !h_bands)) return IDRK_E_ARG;
    const int width = d * ((include_input ? 2 : 0) + 2 * n_bands);
    if (ld_out < width) return IDRK_E_ARG;
    if (n == 0) return 0;
    Bands b; b.n = n_bands;
    for (int i = 0; i < 32; ++i) b.f[i] = i < n_bands ? h_bands[i] : 0.f;
    const long long total = n * (long long)ld_out;
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    IDRK_CUDA_TRY(launch_k(posenc_fwd_kernel, dim3((int)blocks), dim3(threads), 0, (cudaStream_t)stream, x, n, d, ldx, b, include_input, out, ld_out, width));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_posenc_bwd(const float* x, int64_t n, int32_t d, int32_t ldx, const float* h_bands, int32_t n_bands,
                               int32_t include_input, const float* dy, int32_t ld_dy, float* dx, int32_t ld_dx, void* stream) {
    if (!x || !dy || !dx || n < 0 || d < 1 || ldx < d || ld_dx < d || n_bands < 0 || n_bands > 32 || (n_bands > 0 && !h_bands)) return IDRK_E_ARG;
    const int width = d * ((include_input ? 2 : 0) + 2 * n_bands);
    if (ld_dy < width) return IDRK_E_ARG;
    if (n == 0) return 0;
    Bands b; b.n = n_bands;
    for (int i = 0; i < 32; ++i) b.f[i] = i < n_bands ? h_bands[i] : 0.f;
    const long long total = n * (long long)d;
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    IDRK_CUDA_TRY(launch_k(posenc_bwd_kernel, dim3((int)blocks), dim3(threads), 0, (cudaStream_t)stream, x, n, d, ldx, b, include_input, dy, ld_dy, dx, ld_dx));
    IDRK_LAUNCH_CHECK();
    return 0;
}
