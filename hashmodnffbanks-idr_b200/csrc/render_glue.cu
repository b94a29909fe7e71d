// O(rays) ends of the IDR step as single launches (sm_100a):
//
//   idrk_camera_rays   pixel -> world ray + bounding-sphere intersection
//                      utils/rend_util.py:48-75 (get_camera_params), :87-100 (lift), :141-162 (get_sphere_intersection)
//   idrk_idr_loss      IDRLoss.forward and its gradient w.r.t. the three network outputs
//                      model/loss.py:5-71
//
// The reference runs these as ~25 and ~40 eager tensor ops (two K = 3 / K = 4 batched matmuls, boolean-mask indexing with
// a host synchronisation, a dozen scalar reductions).  Here each is one kernel with the reference's operation order in
// fp32 (explicit _rn intrinsics: no fused multiply-add where torch rounds twice), masks as bytes, and sums reduced in a
// fixed order by ONE thread block - rays per step are 2 K - 64 K, so a single block is latency-optimal and makes the
// loss deterministic run to run.
#include "hash_common.cuh"

namespace idrk {

// ------------------------------------------------------------------------------------------
// camera rays + sphere intersection
// ------------------------------------------------------------------------------------------
__global__ void camera_rays_kernel(const float* __restrict__ uv, const float* __restrict__ pose, const float* __restrict__ intr,
                                   int n_images, int n_pixels, float r2, float* __restrict__ dirs, float* __restrict__ cam_out,
                                   float* __restrict__ t_sph, uint8_t* __restrict__ hit) {
    pdl_wait();
    pdl_trigger();
    const long long total = (long long)n_images * n_pixels;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / n_pixels);
        const float* P = pose + 16 * b;
        const float* K = intr + 16 * b;
        const float fx = K[0], sk = K[1], cx = K[2], fy = K[5], cy = K[6];
        const float x = uv[2 * i], y = uv[2 * i + 1];
        // lift (rend_util.py:87-100), z = 1:  (x - cx + cy*sk/fy - sk*y/fy) / fx * z,  (y - cy) / fy * z
        float xl = __fsub_rn(x, cx);
        xl = __fadd_rn(xl, __fdiv_rn(__fmul_rn(cy, sk), fy));
        xl = __fsub_rn(xl, __fdiv_rn(__fmul_rn(sk, y), fy));
        xl = __fdiv_rn(xl, fx);
        const float yl = __fdiv_rn(__fsub_rn(y, cy), fy);
        // world = pose @ (xl, yl, 1, 1): a K = 4 dot product accumulated in order, as the batched matmul does
        float w[3];
#pragma unroll
        for (int rrow = 0; rrow < 3; ++rrow) {
            float acc = __fmul_rn(P[4 * rrow], xl);
            acc = __fmaf_rn(P[4 * rrow + 1], yl, acc);
            acc = __fmaf_rn(P[4 * rrow + 2], 1.f, acc);
            acc = __fmaf_rn(P[4 * rrow + 3], 1.f, acc);
            w[rrow] = acc;
        }
        const float c0 = P[3], c1 = P[7], c2 = P[11];
        float d0 = __fsub_rn(w[0], c0), d1 = __fsub_rn(w[1], c1), d2 = __fsub_rn(w[2], c2);
        // F.normalize: v / max(||v||_2, 1e-12)
        const float nrm = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2))), 1e-12f);
        d0 = __fdiv_rn(d0, nrm); d1 = __fdiv_rn(d1, nrm); d2 = __fdiv_rn(d2, nrm);
        dirs[3 * i] = d0; dirs[3 * i + 1] = d1; dirs[3 * i + 2] = d2;
        if (i % n_pixels == 0) { cam_out[3 * b] = c0; cam_out[3 * b + 1] = c1; cam_out[3 * b + 2] = c2; }
        if (t_sph != nullptr) {
            // get_sphere_intersection (rend_util.py:141-162)
            float dot = __fmul_rn(d0, c0);
            dot = __fmaf_rn(d1, c1, dot);
            dot = __fmaf_rn(d2, c2, dot);
            const float cn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, c0), __fmul_rn(c1, c1)), __fmul_rn(c2, c2)));
            const float under = __fsub_rn(__fmul_rn(dot, dot), __fsub_rn(__fmul_rn(cn, cn), r2));
            const bool h = under > 0.f;
            float ta = 0.f, tb = 0.f;
            if (h) {
                const float root = __fsqrt_rn(under);
                ta = fmaxf(__fsub_rn(-root, dot), 0.f);
                tb = fmaxf(__fsub_rn(root, dot), 0.f);
            }
            t_sph[2 * i] = ta; t_sph[2 * i + 1] = tb;
            hit[i] = h ? 1 : 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// IDR loss + gradient
// ------------------------------------------------------------------------------------------
constexpr int LOSS_THREADS = 1024;

__device__ __forceinline__ float block_sum(float v, float* s_red) {
    // fixed-order reduction: lanes by xor-shuffle, warps by one warp's shuffle
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < (LOSS_THREADS >> 5)) ? s_red[threadIdx.x] : 0.f;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) s_red[32] = t;
    }
    __syncthreads();
    return s_red[32];
}

__global__ void __launch_bounds__(LOSS_THREADS)
idr_loss_kernel(const float* __restrict__ rgb, int ld_rgb, const float* __restrict__ gt, const uint8_t* __restrict__ net_mask,
                const uint8_t* __restrict__ obj_mask, const float* __restrict__ sdf, int ld_sdf, long long n_rays,
                const float* __restrict__ gth, int ld_g, long long n_grad, float eik_w, float mask_w, float alpha,
                float* __restrict__ out4, float* __restrict__ d_rgb, float* __restrict__ d_sdf, float* __restrict__ d_g) {
    pdl_wait();
    pdl_trigger();
    __shared__ float s_red[33];
    const float inv_n = 1.f / (float)n_rays;
    float s_rgb = 0.f, s_mask = 0.f, s_eik = 0.f;
    for (long long i = threadIdx.x; i < n_rays; i += LOSS_THREADS) {
        const bool both = net_mask[i] && obj_mask[i];
        // masked L1 (loss.py:13-20)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float diff = rgb[i * ld_rgb + c] - gt[3 * i + c];
            if (both) s_rgb += fabsf(diff);
            if (d_rgb) d_rgb[3 * i + c] = both ? (diff > 0.f ? inv_n : (diff < 0.f ? -inv_n : 0.f)) : 0.f;
        }
        // mask term on the other rays (loss.py:41-49): BCE-with-logits of -alpha * sdf against the object mask
        const float s = sdf[i * ld_sdf];
        const float l = -alpha * s, t = obj_mask[i] ? 1.f : 0.f;
        const float e = expf(-fabsf(l));
        const float bce = fmaxf(l, 0.f) - l * t + log1pf(e);
        if (!both) s_mask += bce;
        if (d_sdf) {
            const float sig = l >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
            d_sdf[i] = both ? 0.f : -mask_w * (sig - t) * inv_n;           // (1/alpha) * (-alpha) * (sigmoid - t) / N * w
        }
    }
    const float inv_m = n_grad > 0 ? 1.f / (float)n_grad : 0.f;
    for (long long i = threadIdx.x; i < n_grad; i += LOSS_THREADS) {
        // eikonal term (loss.py:22-39): mean((||g|| - 1)^2)
        const float a = gth[i * ld_g], b = gth[i * ld_g + 1], c = gth[i * ld_g + 2];
        const float nrm = sqrtf(a * a + b * b + c * c);
        const float dlt = nrm - 1.f;
        s_eik += dlt * dlt;
        if (d_g) {
            const float k = nrm > 0.f ? eik_w * 2.f * dlt * inv_m / nrm : 0.f;
            d_g[3 * i] = k * a; d_g[3 * i + 1] = k * b; d_g[3 * i + 2] = k * c;
        }
    }
    const float rgb_loss = block_sum(s_rgb, s_red) * inv_n;
    const float mask_loss = (1.f / alpha) * block_sum(s_mask, s_red) * inv_n;
    const float eik_loss = block_sum(s_eik, s_red) * inv_m;
    if (threadIdx.x == 0) {
        out4[0] = rgb_loss + eik_w * eik_loss + mask_w * mask_loss;
        out4[1] = rgb_loss; out4[2] = eik_loss; out4[3] = mask_loss;
    }
}

// y[i] = s[0] * x[i] for three buffers at once (the loss gradients scaled by the incoming d loss)
__global__ void scale3_kernel(const float* __restrict__ s, const float* a, float* ya, long long na, const float* b, float* yb,
                              long long nb, const float* c, float* yc, long long nc) {
    pdl_wait();
    pdl_trigger();
    const float k = *s;
    const long long total = na + nb + nc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (i < na) ya[i] = k * a[i];
        else if (i < na + nb) yb[i - na] = k * b[i - na];
        else yc[i - na - nb] = k * c[i - na - nb];
    }
}

// ------------------------------------------------------------------------------------------
// differentiable d/dx of the Fourier prefix (recorded backward pass of ImplicitNetwork.gradient)
// ------------------------------------------------------------------------------------------
// Row layout of dy: [dy_x(3) | dy_sin(C) | dy_cos(C) | ...].  With xp_j = 2 pi x . B[:, j]:
//   forward   dx   = dy_x + 2 pi * sum_j (dy_sin_j cos xp_j - dy_cos_j sin xp_j) B[:, j]
//   backward  t_j  = 2 pi (G . B[:, j]);  g_dy_x = G,  g_dy_sin_j = t_j cos xp_j,  g_dy_cos_j = -t_j sin xp_j,
//             g_x  = 2 pi * sum_j t_j (-dy_sin_j sin xp_j - dy_cos_j cos xp_j) B[:, j]
// One thread per row; C <= 64.  Same sin / cos evaluation as the encoder kernels (sincos_fast).
__global__ void fourier_dx_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int ld_dy,
                                      const float* __restrict__ B, int C, long long n, float* __restrict__ dx) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float s_B[];
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_B[i] = B[i];
    __syncthreads();
    const float two_pi = 6.283185307179586f;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const float x0 = x[r * ldx], x1 = x[r * ldx + 1], x2 = x[r * ldx + 2];
        const float* d = dy + r * ld_dy;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int j = 0; j < C; ++j) {
            const float b0 = s_B[j], b1 = s_B[C + j], b2 = s_B[2 * C + j];
            float xp = __fmul_rn(__fmul_rn(x0, two_pi), b0);
            xp = __fmaf_rn(__fmul_rn(x1, two_pi), b1, xp);
            xp = __fmaf_rn(__fmul_rn(x2, two_pi), b2, xp);
            float sn, cs;
            sincos_fast(xp, &sn, &cs);
            const float q = d[3 + j] * cs - d[3 + C + j] * sn;
            a0 = fmaf(q, b0, a0); a1 = fmaf(q, b1, a1); a2 = fmaf(q, b2, a2);
        }
        dx[3 * r] = d[0] + two_pi * a0; dx[3 * r + 1] = d[1] + two_pi * a1; dx[3 * r + 2] = d[2] + two_pi * a2;
    }
}

__global__ void fourier_dx_bwd_kernel(const float* __restrict__ G, int ld_g, const float* __restrict__ x, int ldx,
                                      const float* __restrict__ dy, int ld_dy, const float* __restrict__ B, int C, long long n,
                                      float* __restrict__ g_dy, int ld_gdy, int width, float* __restrict__ g_x) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float s_B[];
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_B[i] = B[i];
    __syncthreads();
    const float two_pi = 6.283185307179586f;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const float x0 = x[r * ldx], x1 = x[r * ldx + 1], x2 = x[r * ldx + 2];
        const float g0 = G[r * ld_g], g1 = G[r * ld_g + 1], g2 = G[r * ld_g + 2];
        const float* d = dy + r * ld_dy;
        float* o = g_dy + r * ld_gdy;
        o[0] = g0; o[1] = g1; o[2] = g2;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int j = 0; j < C; ++j) {
            const float b0 = s_B[j], b1 = s_B[C + j], b2 = s_B[2 * C + j];
            float xp = __fmul_rn(__fmul_rn(x0, two_pi), b0);
            xp = __fmaf_rn(__fmul_rn(x1, two_pi), b1, xp);
            xp = __fmaf_rn(__fmul_rn(x2, two_pi), b2, xp);
            float sn, cs;
            sincos_fast(xp, &sn, &cs);
            const float t = two_pi * (g0 * b0 + g1 * b1 + g2 * b2);
            o[3 + j] = t * cs;
            o[3 + C + j] = -t * sn;
            if (g_x != nullptr) {
                const float q = t * (-d[3 + j] * sn - d[3 + C + j] * cs);
                a0 = fmaf(q, b0, a0); a1 = fmaf(q, b1, a1); a2 = fmaf(q, b2, a2);
            }
        }
        for (int c = 3 + 2 * C; c < width; ++c) o[c] = 0.f;       // the level columns do not depend on x (reference mode)
        if (g_x != nullptr) { g_x[3 * r] = two_pi * a0; g_x[3 * r + 1] = two_pi * a1; g_x[3 * r + 2] = two_pi * a2; }
    }
}

}  // namespace idrk

using namespace idrk;

extern "C" int idrk_fourier_dx_fwd(const float* x, int32_t ldx, const float* dy, int32_t ld_dy, const float* B, int32_t n_fourier,
                                   int64_t n, float* dx, void* stream) {
    if (!x || !dy || !B || !dx || n < 0 || ldx < 3 || n_fourier < 1 || n_fourier > 64 || ld_dy < 3 + 2 * n_fourier) return IDRK_E_ARG;
    if (n == 0) return 0;
    long long blocks = (n + 127) / 128;
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    IDRK_CUDA_TRY(launch_k(fourier_dx_fwd_kernel, dim3((unsigned)blocks), dim3(128), (size_t)3 * n_fourier * sizeof(float),
                           (cudaStream_t)stream, x, (int)ldx, dy, (int)ld_dy, B, (int)n_fourier, (long long)n, dx));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_fourier_dx_bwd(const float* g, int32_t ld_g, const float* x, int32_t ldx, const float* dy, int32_t ld_dy,
                                   const float* B, int32_t n_fourier, int64_t n, float* g_dy, int32_t ld_gdy, int32_t width,
                                   float* g_x, void* stream) {
    if (!g || !x || !dy || !B || !g_dy || n < 0 || ldx < 3 || ld_g < 3 || n_fourier < 1 || n_fourier > 64) return IDRK_E_ARG;
    if (ld_dy < 3 + 2 * n_fourier || width < 3 + 2 * n_fourier || ld_gdy < width) return IDRK_E_ARG;
    if (n == 0) return 0;
    long long blocks = (n + 127) / 128;
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    IDRK_CUDA_TRY(launch_k(fourier_dx_bwd_kernel, dim3((unsigned)blocks), dim3(128), (size_t)3 * n_fourier * sizeof(float),
                           (cudaStream_t)stream, g, (int)ld_g, x, (int)ldx, dy, (int)ld_dy, B, (int)n_fourier, (long long)n, g_dy,
                           (int)ld_gdy, (int)width, g_x));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_camera_rays(const float* uv, const float* pose, const float* intrinsics, int32_t n_images, int32_t n_pixels,
                                float radius, float* ray_dirs, float* cam_loc, float* t_sph, uint8_t* hit, void* stream) {
    if (!uv || !pose || !intrinsics || !ray_dirs || !cam_loc || n_images < 0 || n_pixels < 0) return IDRK_E_ARG;
    if ((t_sph == nullptr) != (hit == nullptr)) return IDRK_E_ARG;
    const long long total = (long long)n_images * n_pixels;
    if (total == 0) return 0;
    const double r2 = (double)radius * (double)radius;          // r ** 2 is a Python float in the reference
    long long blocks = (total + 255) / 256;
    if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
    IDRK_CUDA_TRY(launch_k(camera_rays_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, uv, pose, intrinsics,
                           (int)n_images, (int)n_pixels, (float)r2, ray_dirs, cam_loc, t_sph, hit));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_idr_loss(const float* rgb_values, int32_t ld_rgb, const float* rgb_gt, const uint8_t* network_object_mask,
                             const uint8_t* object_mask, const float* sdf_output, int32_t ld_sdf, int64_t n_rays,
                             const float* grad_theta, int32_t ld_grad, int64_t n_grad, float eikonal_weight, float mask_weight,
                             float alpha, float* out_losses, float* d_rgb, float* d_sdf, float* d_grad_theta, void* stream) {
    if (!rgb_values || !rgb_gt || !network_object_mask || !object_mask || !sdf_output || !out_losses) return IDRK_E_ARG;
    if (n_rays <= 0 || n_grad < 0 || ld_rgb < 3 || ld_sdf < 1 || (n_grad > 0 && (!grad_theta || ld_grad < 3))) return IDRK_E_ARG;
    if (!(alpha > 0.f)) return IDRK_E_ARG;
    IDRK_CUDA_TRY(launch_k(idr_loss_kernel, dim3(1), dim3(LOSS_THREADS), 0, (cudaStream_t)stream, rgb_values, (int)ld_rgb, rgb_gt,
                           network_object_mask, object_mask, sdf_output, (int)ld_sdf, (long long)n_rays, grad_theta, (int)ld_grad,
                           (long long)n_grad, eikonal_weight, mask_weight, alpha, out_losses, d_rgb, d_sdf, d_grad_theta));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_scale3(const float* scale, const float* a, float* ya, int64_t na, const float* b, float* yb, int64_t nb,
                           const float* c, float* yc, int64_t nc, void* stream) {
    if (!scale || na < 0 || nb < 0 || nc < 0) return IDRK_E_ARG;
    if ((na && (!a || !ya)) || (nb && (!b || !yb)) || (nc && (!c || !yc))) return IDRK_E_ARG;
    const long long total = na + nb + nc;
    if (total == 0) return 0;
    long long blocks = (total + 255) / 256;
    if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
    IDRK_CUDA_TRY(launch_k(scale3_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, scale, a, ya, (long long)na,
                           b, yb, (long long)nb, c, yc, (long long)nc));
    IDRK_LAUNCH_CHECK();
    return 0;
}
