// K7: fused Fourier-filter-bank encoder (NFFB / StyleModNFFB), forward only, for the ray tracer's no-grad queries.
//
// Restates FourierFilterBanks.forward of the live configuration (freq_enc_type = PositionalEncodingNET,
// layers_type = SIREN, has_out = False; model/embeddings/nffb3d.py:122-194, style block styleMod.py:17-44):
//     z0 = p / bound,  u = (p + bound) / (2 bound)
//     g  = grid_enc(u)[:, 3:]                      (Fourier sin | cos | level features; chunks of 2F columns)
//     z_{j+1} = sin(w0 (A_j z_j + a_j))                                        j = 0 .. NL-1
//     for j >= 1:  E = PositionalEncoding(chunk_{j-1});  e = style ? rownorm(M E + m) : E
//                  f += O (e + z_{j+1}) + o
//     out = [u | f / L]
// The module path runs this as ~45 launches (tiny 56-wide contractions, positional encodings, adds) per SDF query;
// the widths are far below a tensor-core tile, so here ONE warp walks one point through all layers with FP32 FMAs:
// layer weights live transposed in shared memory (lane o reads Wt[k][o]: conflict free), the running vectors in a
// warp-private shared buffer (broadcast reads).  Accurate sinf / cosf for the SIREN and positional terms (arguments
// reach |w0 z| ~ 1e2), the hash-grid columns with the exact arithmetic of hash_encode.cu.
#include "hash_common.cuh"

namespace idrk {

constexpr int NFFB_MAX_W = 64;          // filter-bank width (2 outputs per lane)
constexpr int NFFB_MAX_LAYERS = 16;
constexpr int NFFB_WARPS = 8;

struct NffbDev {
    GridDev grid;
    float bands[32];
    int n_bands, include_input, n_lin, width, chunk, style;
    float bound, w0, levels_div, eps;
    const float* lin_w[NFFB_MAX_LAYERS];      // [width, in]  (in = 3 for layer 0, width after)
    const float* lin_b[NFFB_MAX_LAYERS];
    const float* out_w; const float* out_b;   // [width, width]
    const float* sty_w; const float* sty_b;   // [width, width] (style only)
};

__device__ __forceinline__ float warp_sum_all(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y[o] = b[o] + sum_k Wt[k][o] v[k] for the lane's two outputs (o = lane, lane + 32); v in shared memory
__device__ __forceinline__ void matvec(const float* __restrict__ Wt, const float* __restrict__ bias, const float* v, int n_in,
                                       int W, int lane, float& y0, float& y1) {
    float a0 = 0.f, a1 = 0.f;
    const float* w = Wt + lane;
#pragma unroll 4
    for (int k = 0; k < n_in; ++k) {
        const float vk = v[k];
        a0 = fmaf(w[k * NFFB_MAX_W], vk, a0);
        a1 = fmaf(w[k * NFFB_MAX_W + 32], vk, a1);
    }
    y0 = a0 + (lane < W ? bias[lane] : 0.f);
    y1 = a1 + (lane + 32 < W ? bias[lane + 32] : 0.f);
}

__global__ void __launch_bounds__(NFFB_WARPS * 32)
nffb_encode_fwd_kernel(const NffbDev d, const float* __restrict__ x, long long n, int ldx, float* __restrict__ out, int ld_out,
                       const int* __restrict__ m_count) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float smem[];
    if (m_count != nullptr) { const long long c = *m_count; n = c < n ? c : n; }
    if (n <= 0) return;                     // gated tracer query: skip the weight staging
    const int W = d.width, NL = d.n_lin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // shared layout: transposed weights [mat][k][NFFB_MAX_W] (zero padded), biases [mat][NFFB_MAX_W], per-warp vectors
    const int n_mats = NL + 1 + (d.style ? 1 : 0);                 // SIREN layers, out layer, style transform
    float* s_w = smem;
    float* s_b = s_w + (size_t)n_mats * NFFB_MAX_W * NFFB_MAX_W;
    float* s_v = s_b + n_mats * NFFB_MAX_W + warp * (2 * NFFB_MAX_W + 32);     // [z | e | chunk columns]
    for (int m = 0; m < n_mats; ++m) {
        const float* src_w = m < NL ? d.lin_w[m] : (m == NL ? d.out_w : d.sty_w);
        const float* src_b = m < NL ? d.lin_b[m] : (m == NL ? d.out_b : d.sty_b);
        const int n_in = m == 0 ? 3 : W;
        float* dw = s_w + (size_t)m * NFFB_MAX_W * NFFB_MAX_W;
        for (int i = threadIdx.x; i < NFFB_MAX_W * NFFB_MAX_W; i += blockDim.x) {
            const int k = i / NFFB_MAX_W, o = i - k * NFFB_MAX_W;
            dw[i] = (k < n_in && o < W) ? src_w[o * n_in + k] : 0.f;
        }
        for (int i = threadIdx.x; i < NFFB_MAX_W; i += blockDim.x) s_b[m * NFFB_MAX_W + i] = i < W ? src_b[i] : 0.f;
    }
    __syncthreads();
    float* zs = s_v;                       // current SIREN activations
    float* es = s_v + NFFB_MAX_W;          // encoded chunk / (e + z)
    float* gs = s_v + 2 * NFFB_MAX_W;      // the grid columns the chunks are cut from
    const GridDev& g = d.grid;
    const int C = g.n_fourier, F = g.n_feat;
    const int n_cols = (NL - 1) * d.chunk;                          // grid columns consumed by chunks 0 .. NL-2
    const int dch = d.chunk;
    const int head = d.include_input ? 2 * dch : 0;

    const long long wstride = (long long)gridDim.x * NFFB_WARPS;
    for (long long p = (long long)blockIdx.x * NFFB_WARPS + warp; p < n; p += wstride) {
        const float p0 = x[p * ldx + 0], p1 = x[p * ldx + 1], p2 = x[p * ldx + 2];
        const float z00 = p0 / d.bound, z01 = p1 / d.bound, z02 = p2 / d.bound;
        const float den = 2.f * d.bound;
        const float u0 = (p0 + d.bound) / den, u1 = (p1 + d.bound) / den, u2 = (p2 + d.bound) / den;
        // ---- grid columns (grid_enc(u)[:, 3:]): Fourier sin | cos | level features, arithmetic of hash_encode.cu
        for (int c = lane; c < n_cols; c += 32) {
            float v;
            if (c < 2 * C) {
                const int j = c < C ? c : c - C;
                float xp = __fmul_rn(__fmul_rn(u0, 6.283185307179586f), g.B[j]);
                xp = __fmaf_rn(__fmul_rn(u1, 6.283185307179586f), g.B[C + j], xp);
                xp = __fmaf_rn(__fmul_rn(u2, 6.283185307179586f), g.B[2 * C + j], xp);
                float sn, cs;
                sincos_fast(xp, &sn, &cs);
                v = c < C ? sn : cs;
            } else {
                const int l = (c - 2 * C) / F, f = (c - 2 * C) - l * F;
                const float r = g.res[l];
                const uint32_t h = hash3(trunc_u32(__fmul_rn(u0, r)), trunc_u32(__fmul_rn(u1, r)), trunc_u32(__fmul_rn(u2, r)));
                v = __ldg(g.tables[l] + (size_t)wrap(h, g.rows[l], g.pow2mask[l], g.magic[l]) * F + f);
            }
            gs[c] = v;
        }
        if (lane < 3) zs[lane] = lane == 0 ? z00 : (lane == 1 ? z01 : z02);
        __syncwarp();
        float f0 = 0.f, f1 = 0.f;                                   // the lane's two output features
        for (int j = 0; j < NL; ++j) {
            float y0, y1;
            matvec(s_w + (size_t)j * NFFB_MAX_W * NFFB_MAX_W, s_b + j * NFFB_MAX_W, zs, j == 0 ? 3 : W, W, lane, y0, y1);
            const float z0 = sinf(y0 * d.w0), z1 = sinf(y1 * d.w0);
            __syncwarp();
            zs[lane] = z0; zs[lane + 32] = z1;
            if (j > 0) {
                // E = PositionalEncoding(chunk_{j-1}): [c | c | sin(b0 c) | cos(b0 c) | sin(b1 c) | ...]
                const float* ch = gs + (j - 1) * dch;
                float e0 = 0.f, e1 = 0.f;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int o = lane + 32 * t;
                    float v = 0.f;
                    if (o < W) {
                        if (o < head) {
                            v = ch[o % dch];
                        } else {
                            const int q = (o - head) / dch, jx = (o - head) - q * dch;
                            const float a = __fmul_rn(ch[jx], d.bands[q >> 1]);
                            v = (q & 1) ? cosf(a) : sinf(a);
                        }
                    }
                    if (t == 0) e0 = v; else e1 = v;
                }
                if (d.style) {
                    es[lane] = e0; es[lane + 32] = e1;
                    __syncwarp();
                    float s0, s1;
                    matvec(s_w + (size_t)(NL + 1) * NFFB_MAX_W * NFFB_MAX_W, s_b + (NL + 1) * NFFB_MAX_W, es, W, W, lane, s0, s1);
                    const float m0 = lane < W ? s0 : 0.f, m1 = lane + 32 < W ? s1 : 0.f;
                    const float mu = warp_sum_all(m0 + m1) / (float)W;
                    const float c0 = lane < W ? s0 - mu : 0.f, c1 = lane + 32 < W ? s1 - mu : 0.f;
                    const float var = warp_sum_all(c0 * c0 + c1 * c1) / (float)W;
                    const float inv = 1.f / sqrtf(var + d.eps);
                    e0 = c0 * inv; e1 = c1 * inv;
                    __syncwarp();
                }
                es[lane] = e0 + z0; es[lane + 32] = e1 + z1;
                __syncwarp();
                float o0, o1;
                matvec(s_w + (size_t)NL * NFFB_MAX_W * NFFB_MAX_W, s_b + NL * NFFB_MAX_W, es, W, W, lane, o0, o1);
                f0 += o0; f1 += o1;
            }
            __syncwarp();
        }
        float* orow = out + p * (long long)ld_out;
        if (lane < 3) orow[lane] = lane == 0 ? u0 : (lane == 1 ? u1 : u2);
        if (lane < W) orow[3 + lane] = f0 / d.levels_div;
        if (lane + 32 < W) orow[3 + lane + 32] = f1 / d.levels_div;
        for (int c = 3 + W + lane; c < ld_out; c += 32) orow[c] = 0.f;
        __syncwarp();
    }
}

}  // namespace idrk

using namespace idrk;

extern "C" int idrk_nffb_encode_fwd(const idrk_nffb_t* h, const float* x, int64_t n, int32_t ldx, float* out, int32_t ld_out,
                                    const int32_t* m_count, void* stream) {
    if (!h || !x || !out || n < 0 || ldx < 3) return IDRK_E_ARG;
    NffbDev d;
    int rc = fill_grid(&h->grid, d.grid);
    if (rc) return rc;
    if (h->grid.frac_mode != IDRK_HASH_REFERENCE) return IDRK_E_UNSUP;
    if (h->width < 1 || h->width > NFFB_MAX_W || h->n_lin < 2 || h->n_lin > NFFB_MAX_LAYERS) return IDRK_E_UNSUP;
    if (h->n_bands < 0 || h->n_bands > 32 || h->chunk < 1 || !(h->bound > 0.f) || h->n_levels_div < 1) return IDRK_E_ARG;
    if (h->width != h->chunk * ((h->include_input ? 2 : 0) + 2 * h->n_bands)) return IDRK_E_ARG;
    const int n_cols = (h->n_lin - 1) * h->chunk;
    if (n_cols > 32 || n_cols > 2 * h->grid.n_fourier + h->grid.n_levels * h->grid.n_feat) return IDRK_E_UNSUP;
    if (ld_out < 3 + h->width || !h->out_w || !h->out_b) return IDRK_E_ARG;
    if (h->style && (!h->style_w || !h->style_b)) return IDRK_E_ARG;
    if (n == 0) return 0;
    for (int i = 0; i < 32; ++i) d.bands[i] = i < h->n_bands ? h->bands[i] : 0.f;
    d.n_bands = h->n_bands; d.include_input = h->include_input; d.n_lin = h->n_lin; d.width = h->width; d.chunk = h->chunk;
    d.style = h->style; d.bound = h->bound; d.w0 = h->w0; d.levels_div = (float)h->n_levels_div; d.eps = h->eps;
    for (int i = 0; i < NFFB_MAX_LAYERS; ++i) {
        d.lin_w[i] = i < h->n_lin ? h->lin_w[i] : nullptr;
        d.lin_b[i] = i < h->n_lin ? h->lin_b[i] : nullptr;
        if (i < h->n_lin && (!d.lin_w[i] || !d.lin_b[i])) return IDRK_E_ARG;
    }
    d.out_w = h->out_w; d.out_b = h->out_b; d.sty_w = h->style_w; d.sty_b = h->style_b;
    const int n_mats = h->n_lin + 1 + (h->style ? 1 : 0);
    const size_t smem = ((size_t)n_mats * NFFB_MAX_W * NFFB_MAX_W + (size_t)n_mats * NFFB_MAX_W +
                         (size_t)NFFB_WARPS * (2 * NFFB_MAX_W + 32)) * sizeof(float);
    if (smem > 220 * 1024) return IDRK_E_UNSUP;
    IDRK_CUDA_TRY(cudaFuncSetAttribute(nffb_encode_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nffb_encode_fwd_kernel, NFFB_WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long grid = (long long)per_sm * sm_count();
    const long long need = (n + NFFB_WARPS - 1) / NFFB_WARPS;
    if (grid > need) grid = need;
    IDRK_CUDA_TRY(launch_k(nffb_encode_fwd_kernel, dim3((unsigned)grid), dim3(NFFB_WARPS * 32), smem, (cudaStream_t)stream,
                           d, x, (long long)n, (int)ldx, out, (int)ld_out, m_count));
    IDRK_LAUNCH_CHECK();
    return 0;
}
