// Small memory-bound helpers around the MLP tiles: TF32 hi/lo operand split, weight-norm forward /
// backward (legacy nn.utils.weight_norm, dim=0: W = g * v / ||v||_row), bias gradients (column sums)
// and the SDF head (last Linear row 0 + the Laplace-density squash of
// implicit_differentiable_renderer.py:112 / density_net.py:20-30).
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace idrk {

__device__ __forceinline__ float tf32_round(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

// 16-bit pair x ~= h + l * 2^-11 (fmt 0: fp16, clamped to the format's range; 1: bf16), see csrc/gemm_p16.cu
__device__ __forceinline__ void pair16(float x, int fmt, uint16_t& h, uint16_t& l) {
    if (fmt == 0) {
        x = fminf(fmaxf(x, -65504.f), 65504.f);
        const __half hh = __float2half_rn(x);
        const __half ll = __float2half_rn((x - __half2float(hh)) * 2048.f);
        h = __half_as_ushort(hh); l = __half_as_ushort(ll);
    } else {
        const __nv_bfloat16 hh = __float2bfloat16_rn(x);
        const __nv_bfloat16 ll = __float2bfloat16_rn((x - __bfloat162float(hh)) * 2048.f);
        h = __bfloat16_as_ushort(hh); l = __bfloat16_as_ushort(ll);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void split_tf32_kernel(const float* __restrict__ x, long long rows, int cols, int ldx, float scale,
                                  float* __restrict__ hi, float* __restrict__ lo, int ldo, int pad_cols,
                                  const int* __restrict__ m_count) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    long long r_eff = rows;
    if (m_count) { const long long c = *m_count; r_eff = c < rows ? c : rows; }
    const int w = cols + pad_cols;
    const long long total = r_eff * (long long)w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / w;
        const int c = (int)(i - r * w);
        float h = 0.f, l = 0.f;
        if (c < cols) {
            const float v = x[r * ldx + c] * scale;
            if (lo) { h = tf32_round(v); l = tf32_round(v - h); } else { h = v; }
        }
        hi[r * ldo + c] = h;
        if (lo) lo[r * ldo + c] = l;
    }
}

// one warp per output row n
__global__ void weight_norm_fwd_kernel(const float* __restrict__ g, const float* __restrict__ v, int N, int K, int ldv,
                                       float* __restrict__ W, float* __restrict__ W_hi, float* __restrict__ W_lo, int ldw,
                                       float* __restrict__ Wt, float* __restrict__ Wt_hi, float* __restrict__ Wt_lo, int ldwt,
                                       uint16_t* __restrict__ Wp_h, uint16_t* __restrict__ Wp_l, int ldp, int p_fmt) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const float* row = v + (long long)n * ldv;
    float ss = 0.f;
    for (int k = lane; k < K; k += 32) ss = fmaf(row[k], row[k], ss);
    ss = warp_sum(ss);
    const float scale = g ? g[n] / sqrtf(ss) : 1.f;
    for (int k = lane; k < ldw; k += 32) {
        const float w = k < K ? row[k] * scale : 0.f;
        const float h = tf32_round(w), l = tf32_round(w - h);
        if (W) W[(long long)n * ldw + k] = w;
        if (W_hi) { W_hi[(long long)n * ldw + k] = h; W_lo[(long long)n * ldw + k] = l; }
        if (Wp_h) { uint16_t ph, pl; pair16(w, p_fmt, ph, pl); Wp_h[(long long)n * ldp + k] = ph; Wp_l[(long long)n * ldp + k] = pl; }
        if (k < K) {
            if (Wt) Wt[(long long)k * ldwt + n] = w;
            if (Wt_hi) { Wt_hi[(long long)k * ldwt + n] = h; Wt_lo[(long long)k * ldwt + n] = l; }
        }
    }
}

__global__ void weight_norm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ v, const float* __restrict__ dW,
                                       int N, int K, int ldv, int lddw, float* __restrict__ dg, float* __restrict__ dv, int lddv,
                                       int accumulate) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const float* vr = v + (long long)n * ldv;
    const float* dr = dW + (long long)n * lddw;
    float ss = 0.f, dot = 0.f;
    for (int k = lane; k < K; k += 32) { ss = fmaf(vr[k], vr[k], ss); dot = fmaf(vr[k], dr[k], dot); }
    ss = warp_sum(ss); dot = warp_sum(dot);
    const float nrm = sqrtf(ss);
    const float gn = g[n];
    if (lane == 0) dg[n] = (accumulate ? dg[n] : 0.f) + dot / nrm;
    const float a = gn / nrm, b = gn * dot / (nrm * ss);
    for (int k = lane; k < K; k += 32) {
        float* o = dv + (long long)n * lddv + k;
        *o = (accumulate ? *o : 0.f) + a * dr[k] - b * vr[k];
    }
}

// out[c] += sum_r x[r, c]
__global__ void colsum_kernel(const float* __restrict__ x, long long rows, int cols, int ldx, float* __restrict__ out) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;                    // 8 row lanes
    __shared__ float part[8][33];
    float acc = 0.f;
    if (c < cols) {
        const long long chunk = (rows + gridDim.y - 1) / gridDim.y;
        const long long r0 = blockIdx.y * chunk, r1 = min(rows, r0 + chunk);
        for (long long r = r0 + ry; r < r1; r += 8) acc += x[r * ldx + c];
    }
    part[ry][threadIdx.x & 31] = acc;
    __syncthreads();
    if (ry == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
        atomicAdd(out + c, s);
    }
}

__device__ __forceinline__ float sdf_squash(float s, float beta) {
    // rho = (1/beta) * (0.5 + 0.5 * sign(s) * expm1(-|s| / beta))   (density_net.py:27)
    const float sg = s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f);
    const float rho = (1.f / beta) * (0.5f + 0.5f * sg * expm1f(-fabsf(s) / beta));
    return tanhf(s / (2.f + rho));
}

// one warp per point: sdf = squash(dot(h[p, :K], w) + b)
__global__ void sdf_head_kernel(const float* __restrict__ h, long long rows, int K, int ldh, const float* __restrict__ w,
                                const float* __restrict__ bias, float beta, float* __restrict__ out,
                                const int* __restrict__ m_count) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    long long r_eff = rows;
    if (m_count) { const long long c = *m_count; r_eff = c < rows ? c : rows; }
    const int lane = threadIdx.x & 31;
    const long long wpg = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long p = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); p < r_eff; p += wpg) {
        const float* row = h + p * ldh;
        float acc = 0.f;
        for (int k = lane; k < K; k += 32) acc = fmaf(row[k], __ldg(w + k), acc);
        acc = warp_sum(acc);
        if (lane == 0) out[p] = sdf_squash(acc + bias[0], beta);
    }
}

__global__ void sdf_squash_kernel(const float* __restrict__ s, long long n, float beta, float* __restrict__ out, float* __restrict__ dout) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = s[i];
        const float sg = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
        const float rho = (1.f / beta) * (0.5f + 0.5f * sg * expm1f(-fabsf(v) / beta));
        const float t = tanhf(v / (2.f + rho));
        out[i] = t;
        if (dout) dout[i] = (1.f - t * t) / (2.f + rho);
    }
}

// out = x with column 0 replaced by tanh(s / (2 + rho(s))) (the last step of ImplicitNetwork.forward,
// implicit_differentiable_renderer.py:108-113: rho is a constant for autograd), d = d out_0 / d s, d2 = d^2 out_0 / d s^2
__global__ void sdf_squash_rows_kernel(const float* __restrict__ x, long long rows, int cols, int ldx, float beta,
                                       float* __restrict__ out, int ld_out, float* __restrict__ d, float* __restrict__ d2) {
    pdl_wait();
    pdl_trigger();
    const long long total = rows * (long long)ld_out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ld_out;
        const int c = (int)(i - r * ld_out);
        float v = c < cols ? x[r * ldx + c] : 0.f;
        if (c == 0) {
            const float sg = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
            const float rho = (1.f / beta) * (0.5f + 0.5f * sg * expm1f(-fabsf(v) / beta));
            const float k = 1.f / (2.f + rho);
            const float t = tanhf(v * k);
            const float dd = (1.f - t * t) * k;
            if (d) d[r] = dd;
            if (d2) d2[r] = -2.f * t * k * dd;
            v = t;
        }
        out[i] = v;
    }
}

// backward of H = scale * act(Z), S = act'(Z):  dZ = dH * S * scale + dS * act''(Z), also written as a 3xTF32 operand pair
__global__ void act_bwd_kernel(const float* __restrict__ dH, int ld_dh, const float* __restrict__ dS, int ld_ds,
                               const float* __restrict__ S, int ld_s, const float* __restrict__ H, int ld_h,
                               long long rows, int cols, int mode, float act, float scale,
                               float* __restrict__ dZ, float* __restrict__ hi, float* __restrict__ lo, int ld_out,
                               uint16_t* __restrict__ ph, uint16_t* __restrict__ pl, int ldp, int p_fmt) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const long long total = rows * (long long)ld_out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ld_out;
        const int c = (int)(i - r * ld_out);
        float v = 0.f;
        if (c < cols) {
            const float s = S[r * ld_s + c];
            if (dH) v = dH[r * ld_dh + c] * s * scale;
            if (dS) {
                float s2 = 0.f;
                if (mode == IDRK_EPI_SOFTPLUS) s2 = act * s * (1.f - s);
                else if (mode == IDRK_EPI_SINE) s2 = -(act * act) * H[r * ld_h + c] / scale;
                else if (mode == IDRK_EPI_TANH) s2 = -2.f * H[r * ld_h + c] * s / scale;
                v = fmaf(dS[r * ld_ds + c], s2, v);
            }
        }
        dZ[i] = v;
        if (hi) { const float h = tf32_round(v); hi[i] = h; lo[i] = tf32_round(v - h); }
        if (ph) { uint16_t a, b; pair16(v, p_fmt, a, b); ph[r * ldp + c] = a; pl[r * ldp + c] = b; }
    }
}

// float4 form of act_bwd_kernel: every leading dimension a multiple of 4 and every pointer 16-byte aligned (the padded
// operand layout), one thread per 4 columns.  Same arithmetic per element.
__global__ void act_bwd_vec4_kernel(const float* __restrict__ dH, int ld_dh, const float* __restrict__ dS, int ld_ds,
                                    const float* __restrict__ S, int ld_s, const float* __restrict__ H, int ld_h,
                                    long long rows, int cols, int mode, float act, float scale,
                                    float* __restrict__ dZ, float* __restrict__ hi, float* __restrict__ lo, int ld_out,
                                    uint16_t* __restrict__ ph, uint16_t* __restrict__ pl, int ldp, int p_fmt) {
    pdl_wait();
    pdl_trigger();
    const int q = ld_out >> 2;
    const long long total = rows * (long long)q;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / q;
        const int c = (int)(i - r * q) << 2;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (c < cols) {
            const float4 s4 = *reinterpret_cast<const float4*>(S + r * ld_s + c);
            const float s[4] = {s4.x, s4.y, s4.z, s4.w};
            if (dH) {
                const float4 d4 = *reinterpret_cast<const float4*>(dH + r * ld_dh + c);
                v[0] = d4.x * s[0] * scale; v[1] = d4.y * s[1] * scale; v[2] = d4.z * s[2] * scale; v[3] = d4.w * s[3] * scale;
            }
            if (dS) {
                const float4 g4 = *reinterpret_cast<const float4*>(dS + r * ld_ds + c);
                const float g[4] = {g4.x, g4.y, g4.z, g4.w};
                float h[4] = {0.f, 0.f, 0.f, 0.f};
                if (mode == IDRK_EPI_SINE || mode == IDRK_EPI_TANH) {
                    const float4 h4 = *reinterpret_cast<const float4*>(H + r * ld_h + c);
                    h[0] = h4.x; h[1] = h4.y; h[2] = h4.z; h[3] = h4.w;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float s2 = 0.f;
                    if (mode == IDRK_EPI_SOFTPLUS) s2 = act * s[k] * (1.f - s[k]);
                    else if (mode == IDRK_EPI_SINE) s2 = -(act * act) * h[k] / scale;
                    else if (mode == IDRK_EPI_TANH) s2 = -2.f * h[k] * s[k] / scale;
                    v[k] = fmaf(g[k], s2, v[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) if (c + k >= cols) v[k] = 0.f;       // pad columns of the last group
        }
        *reinterpret_cast<float4*>(dZ + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
        if (hi) {
            float a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { a[k] = tf32_round(v[k]); b[k] = tf32_round(v[k] - a[k]); }
            *reinterpret_cast<float4*>(hi + i * 4) = make_float4(a[0], a[1], a[2], a[3]);
            *reinterpret_cast<float4*>(lo + i * 4) = make_float4(b[0], b[1], b[2], b[3]);
        }
        if (ph) {
            uint16_t a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) pair16(v[k], p_fmt, a[k], b[k]);
            *reinterpret_cast<uint2*>(ph + r * ldp + c) = make_uint2(a[0] | ((uint32_t)a[1] << 16), a[2] | ((uint32_t)a[3] << 16));
            *reinterpret_cast<uint2*>(pl + r * ldp + c) = make_uint2(b[0] | ((uint32_t)b[1] << 16), b[2] | ((uint32_t)b[3] << 16));
        }
    }
}

static inline int ew_blocks(long long total, int threads) {
    long long b = (total + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace idrk

using namespace idrk;

extern "C" int idrk_split_tf32(const float* x, int64_t rows, int32_t cols, int32_t ldx, float scale, float* hi, float* lo,
                               int32_t ld_out, int32_t pad_cols, const int32_t* m_count, void* stream) {
    if (!x || !hi || rows < 0 || cols < 1 || ldx < cols || pad_cols < 0 || ld_out < cols + pad_cols) return IDRK_E_ARG;
    if (rows == 0) return 0;
    IDRK_CUDA_TRY(launch_k(split_tf32_kernel, dim3(ew_blocks(rows * (long long)(cols + pad_cols), 256)), dim3(256), 0, (cudaStream_t)stream, 
        x, rows, cols, ldx, scale, hi, lo, ld_out, pad_cols, m_count));
    IDRK_LAUNCH_CHECK();
    return 0;
}

static int weight_norm_fwd_impl(const float* g, const float* v, int32_t N, int32_t K, int32_t ldv,
                               float* W, float* W_hi, float* W_lo, int32_t ldw,
                               float* Wt, float* Wt_hi, float* Wt_lo, int32_t ldwt,
                               void* Wp_h, void* Wp_l, int32_t ldp, int32_t p_fmt, void* stream) {
    if (!v || N < 1 || K < 1 || ldv < K || ldw < K) return IDRK_E_ARG;
    if ((W_hi == nullptr) != (W_lo == nullptr) || (Wt_hi == nullptr) != (Wt_lo == nullptr)) return IDRK_E_ARG;
    if ((Wt || Wt_hi) && ldwt < N) return IDRK_E_ARG;
    if ((Wp_h == nullptr) != (Wp_l == nullptr) || (Wp_h && (ldp < ldw || (p_fmt & ~1)))) return IDRK_E_ARG;
    IDRK_CUDA_TRY(launch_k(weight_norm_fwd_kernel, dim3((N + 7) / 8), dim3(256), 0, (cudaStream_t)stream, g, v, N, K, ldv, W, W_hi, W_lo, ldw,
                           Wt, Wt_hi, Wt_lo, ldwt, (uint16_t*)Wp_h, (uint16_t*)Wp_l, ldp, p_fmt));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_weight_norm_fwd(const float* g, const float* v, int32_t N, int32_t K, int32_t ldv,
                                    float* W, float* W_hi, float* W_lo, int32_t ldw,
                                    float* Wt, float* Wt_hi, float* Wt_lo, int32_t ldwt, void* stream) {
    return weight_norm_fwd_impl(g, v, N, K, ldv, W, W_hi, W_lo, ldw, Wt, Wt_hi, Wt_lo, ldwt, nullptr, nullptr, 0, 0, stream);
}

extern "C" int idrk_weight_norm_fwd_p16(const float* g, const float* v, int32_t N, int32_t K, int32_t ldv, float* W, int32_t ldw,
                                        void* W_h, void* W_l, int32_t ldp, int32_t fmt, void* stream) {
    if (!W_h || !W_l) return IDRK_E_ARG;
    return weight_norm_fwd_impl(g, v, N, K, ldv, W, nullptr, nullptr, ldw, nullptr, nullptr, nullptr, 0, W_h, W_l, ldp, fmt, stream);
}

extern "C" int idrk_weight_norm_bwd(const float* g, const float* v, const float* dW, int32_t N, int32_t K, int32_t ldv,
                                    int32_t lddw, float* dg, float* dv, int32_t lddv, int32_t accumulate, void* stream) {
    if (!g || !v || !dW || !dg || !dv || N < 1 || K < 1 || ldv < K || lddw < K || lddv < K) return IDRK_E_ARG;
    IDRK_CUDA_TRY(launch_k(weight_norm_bwd_kernel, dim3((N + 7) / 8), dim3(256), 0, (cudaStream_t)stream, g, v, dW, N, K, ldv, lddw, dg, dv, lddv, accumulate));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_colsum(const float* x, int64_t rows, int32_t cols, int32_t ldx, float* out, void* stream) {
    if (!x || !out || rows < 0 || cols < 1 || ldx < cols) return IDRK_E_ARG;
    if (rows == 0) return 0;
    long long ysplit = rows / 64;                 // ~8 rows per thread: enough CTAs to cover the SMs
    if (ysplit < 1) ysplit = 1;
    if (ysplit > 128) ysplit = 128;
    dim3 grid((cols + 31) / 32, (unsigned)ysplit);
    IDRK_CUDA_TRY(launch_k(colsum_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, rows, cols, ldx, out));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_sdf_head(const float* h, int64_t rows, int32_t K, int32_t ldh, const float* w, const float* bias,
                             float beta, float* out, const int32_t* m_count, void* stream) {
    if (!h || !w || !bias || !out || rows < 0 || K < 1 || ldh < K || !(beta > 0.f)) return IDRK_E_ARG;
    if (rows == 0) return 0;
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    IDRK_CUDA_TRY(launch_k(sdf_head_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, h, rows, K, ldh, w, bias, beta, out, m_count));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_sdf_squash(const float* s, int64_t n, float beta, float* out, float* dout, void* stream) {
    if (!s || !out || n < 0 || !(beta > 0.f)) return IDRK_E_ARG;
    if (n == 0) return 0;
    IDRK_CUDA_TRY(launch_k(sdf_squash_kernel, dim3(ew_blocks(n, 256)), dim3(256), 0, (cudaStream_t)stream, s, n, beta, out, dout));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_sdf_squash_rows(const float* x, int64_t rows, int32_t cols, int32_t ldx, float beta, float* out, int32_t ld_out,
                                    float* d, float* d2, void* stream) {
    if (!x || !out || rows < 0 || cols < 1 || ldx < cols || ld_out < cols || !(beta > 0.f)) return IDRK_E_ARG;
    if (rows == 0) return 0;
    IDRK_CUDA_TRY(launch_k(sdf_squash_rows_kernel, dim3(ew_blocks(rows * (long long)ld_out, 256)), dim3(256), 0, (cudaStream_t)stream,
                           x, (long long)rows, (int)cols, (int)ldx, beta, out, (int)ld_out, d, d2));
    IDRK_LAUNCH_CHECK();
    return 0;
}

static int act_bwd_impl(const float* dH, int32_t ld_dh, const float* dS, int32_t ld_ds, const float* S, int32_t ld_s,
                        const float* H, int32_t ld_h, int64_t rows, int32_t cols, int32_t mode, float act, float scale,
                        float* dZ, float* dZ_hi, float* dZ_lo, int32_t ld_out, void* P_h, void* P_l, int32_t ldp, int32_t p_fmt,
                        void* stream) {
    if (!S || !dZ || (!dH && !dS) || rows < 0 || cols < 1 || ld_out < cols || ld_s < cols) return IDRK_E_ARG;
    if ((dZ_hi == nullptr) != (dZ_lo == nullptr)) return IDRK_E_ARG;
    if ((P_h == nullptr) != (P_l == nullptr) || (P_h && (ldp < ld_out || (p_fmt & ~1)))) return IDRK_E_ARG;
    if (dS && (mode == IDRK_EPI_SINE || mode == IDRK_EPI_TANH) && !H) return IDRK_E_ARG;
    if (rows == 0) return 0;
    const bool need_h = dS && (mode == IDRK_EPI_SINE || mode == IDRK_EPI_TANH);
    const bool vec4 = (ld_out & 3) == 0 && (ld_s & 3) == 0 && aligned16(S) && aligned16(dZ) &&
                      (!dH || ((ld_dh & 3) == 0 && aligned16(dH))) && (!dS || ((ld_ds & 3) == 0 && aligned16(dS))) &&
                      (!need_h || ((ld_h & 3) == 0 && aligned16(H))) && (!dZ_hi || (aligned16(dZ_hi) && aligned16(dZ_lo))) &&
                      ld_s >= ((cols + 3) & ~3) && (!dH || ld_dh >= ((cols + 3) & ~3)) && (!dS || ld_ds >= ((cols + 3) & ~3)) &&
                      (!need_h || ld_h >= ((cols + 3) & ~3)) &&
                      (!P_h || ((ldp & 3) == 0 && (reinterpret_cast<uintptr_t>(P_h) & 7) == 0 && (reinterpret_cast<uintptr_t>(P_l) & 7) == 0));
    if (vec4) {
        IDRK_CUDA_TRY(launch_k(act_bwd_vec4_kernel, dim3(ew_blocks(rows * (long long)(ld_out >> 2), 256)), dim3(256), 0, (cudaStream_t)stream,
            dH, ld_dh, dS, ld_ds, S, ld_s, H, ld_h, rows, cols, mode, act, scale, dZ, dZ_hi, dZ_lo, ld_out,
            (uint16_t*)P_h, (uint16_t*)P_l, ldp, p_fmt));
        IDRK_LAUNCH_CHECK();
        return 0;
    }
    IDRK_CUDA_TRY(launch_k(act_bwd_kernel, dim3(ew_blocks(rows * (long long)ld_out, 256)), dim3(256), 0, (cudaStream_t)stream,
        dH, ld_dh, dS, ld_ds, S, ld_s, H, ld_h, rows, cols, mode, act, scale, dZ, dZ_hi, dZ_lo, ld_out,
        (uint16_t*)P_h, (uint16_t*)P_l, ldp, p_fmt));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_act_bwd(const float* dH, int32_t ld_dh, const float* dS, int32_t ld_ds, const float* S, int32_t ld_s,
                            const float* H, int32_t ld_h, int64_t rows, int32_t cols, int32_t mode, float act, float scale,
                            float* dZ, float* dZ_hi, float* dZ_lo, int32_t ld_out, void* stream) {
    return act_bwd_impl(dH, ld_dh, dS, ld_ds, S, ld_s, H, ld_h, rows, cols, mode, act, scale, dZ, dZ_hi, dZ_lo, ld_out,
                        nullptr, nullptr, 0, 0, stream);
}

extern "C" int idrk_act_bwd_p16(const float* dH, int32_t ld_dh, const float* dS, int32_t ld_ds, const float* S, int32_t ld_s,
                                const float* H, int32_t ld_h, int64_t rows, int32_t cols, int32_t mode, float act, float scale,
                                float* dZ, int32_t ld_out, void* dZ_h, void* dZ_l, int32_t ld_pair, int32_t fmt, void* stream) {
    if (!dZ_h || !dZ_l) return IDRK_E_ARG;
    return act_bwd_impl(dH, ld_dh, dS, ld_ds, S, ld_s, H, ld_h, rows, cols, mode, act, scale, dZ, nullptr, nullptr, ld_out,
                        dZ_h, dZ_l, ld_pair, fmt, stream);
}
