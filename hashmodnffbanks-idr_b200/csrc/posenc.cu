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

// Backward of the map (x, dy) -> dx computed by posenc_bwd_kernel (the RECORDED backward of the encoding, i.e. the
// double backward of a loss on d y / d x - ImplicitNetwork.gradient with create_graph=True through the filter banks'
// positional encodings, nffb3d.py:170-173 under implicit_differentiable_renderer.py:116-128).  Given g = d L / d dx:
//   g_dy[c]  = g[c % d]                                   for the copied input columns
//   g_dy[sin_q, j] =  g_j f_q cos(f_q x_j),   g_dy[cos_q, j] = -g_j f_q sin(f_q x_j)
//   g_x[j]   = -g_j sum_q f_q^2 (dy[sin_q, j] sin(f_q x_j) + dy[cos_q, j] cos(f_q x_j))
__global__ void posenc_dx_bwd_kernel(const float* __restrict__ g, int ld_g, const float* __restrict__ x, long long n, int d, int ldx,
                                     Bands b, int include_input, const float* __restrict__ dy, int ld_dy,
                                     float* __restrict__ g_dy, int ld_gdy, int width, float* __restrict__ g_x, int ld_gx) {
    pdl_wait();
    pdl_trigger();
    const int head = include_input ? 2 * d : 0;
    if (g_dy != nullptr) {
        const long long total = n * (long long)ld_gdy;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long p = i / ld_gdy;
            int c = (int)(i - p * ld_gdy);
            float v = 0.f;
            if (c < width) {
                if (c < head) v = g[p * ld_g + (c % d)];
                else {
                    c -= head;
                    const int k = c / d, j = c - k * d;
                    const float f = b.f[k >> 1];
                    const float a = __fmul_rn(x[p * ldx + j], f);
                    v = g[p * ld_g + j] * f * ((k & 1) ? -sinf(a) : cosf(a));
                }
            }
            g_dy[i] = v;
        }
    }
    if (g_x != nullptr) {
        const long long total = n * (long long)d;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long p = i / d;
            const int j = (int)(i - p * d);
            const float* u = dy + p * ld_dy;
            const float xv = x[p * ldx + j];
            float acc = 0.f;
            for (int q = 0; q < b.n; ++q) {
                const float f = b.f[q];
                float sn, cs;
                sincosf(__fmul_rn(xv, f), &sn, &cs);
                acc = fmaf(-(u[head + (2 * q) * d + j] * sn + u[head + (2 * q + 1) * d + j] * cs), f * f, acc);
            }
            g_x[p * ld_gx + j] = g[p * ld_g + j] * acc;
        }
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

extern "C" int idrk_posenc_dx_bwd(const float* g, int32_t ld_g, const float* x, int64_t n, int32_t d, int32_t ldx, const float* h_bands,
                                  int32_t n_bands, int32_t include_input, const float* dy, int32_t ld_dy, float* g_dy, int32_t ld_gdy,
                                  float* g_x, int32_t ld_gx, void* stream) {
    if (!g || !x || !dy || (!g_dy && !g_x) || n < 0 || d < 1 || ldx < d || ld_g < d || n_bands < 0 || n_bands > 32 || (n_bands > 0 && !h_bands))
        return IDRK_E_ARG;
    const int width = d * ((include_input ? 2 : 0) + 2 * n_bands);
    if (ld_dy < width || (g_dy && ld_gdy < width) || (g_x && ld_gx < d)) return IDRK_E_ARG;
    if (n == 0) return 0;
    Bands b; b.n = n_bands;
    for (int i = 0; i < 32; ++i) b.f[i] = i < n_bands ? h_bands[i] : 0.f;
    const long long total = n * (long long)(g_dy ? ld_gdy : d);
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    IDRK_CUDA_TRY(launch_k(posenc_dx_bwd_kernel, dim3((int)blocks), dim3(threads), 0, (cudaStream_t)stream, g, ld_g, x, n, d, ldx, b,
                           include_input, dy, ld_dy, g_dy, ld_gdy, width, g_x, ld_gx));
    IDRK_LAUNCH_CHECK();
    return 0;
}
