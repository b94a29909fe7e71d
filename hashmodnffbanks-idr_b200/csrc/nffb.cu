// K7: fused Fourier-filter-bank encoder (NFFB / StyleModNFFB), forward only, for the ray tracer's no-grad queries.
//
// Restates FourierFilterBanks.forward of the live configuration (freq_enc_type = PositionalEncodingNET,
// layers_type = SIREN, has_out = False; model/embeddings/nffb3d.py:122-194, style block styleMod.py:17-44):
//     z0 = p / bound,  u = (p + bound) / (2 bound)
//     g  = grid_enc(u)[:, 3:]                      (Fourier sin | cos | level features; chunks of 2F columns)
//     z_{j+1} = sin(w0 (A_j z_j + a_j))                                        j = 0 .. NL-1
//     for j >= 1:  E = PositionalEncoding(chunk_{j-1});  e = style ? rownorm(M E + m) : E
//                  f += O (e + z_{j+1}) + o            (evaluated as O (sum_j (e + z_{j+1})) + (NL - 1) o: out_layer is linear)
//     out = [u | f / L]
// The module path runs this as ~45 launches (tiny 56-wide contractions, positional encodings, adds) per SDF query;
// the widths are far below a tensor-core tile, so here ONE warp walks one point through all layers with FP32 FMAs:
// layer weights live transposed in shared memory (lane o reads Wt[k][o]: conflict free), the running vectors in a
// warp-private shared buffer (broadcast reads).  Accurate sinf / cosf for the SIREN and positional terms (arguments
// reach |w0 z| ~ 1e2), the hash-grid columns with the exact arithmetic of hash_encode.cu.
#include <cuda_fp16.h>
#include "hash_common.cuh"

namespace idrk {

constexpr int NFFB_MAX_W = 64;          // filter-bank width (2 outputs per lane)
constexpr int NFFB_LDW = NFFB_MAX_W + 1;   // pitch of a staged (transposed) weight row: the staging writes of consecutive k for
                                           // one output column land in different banks, the mat-vec reads (consecutive columns
                                           // for one k) stay conflict-free
constexpr int NFFB_MAT = NFFB_MAX_W * NFFB_LDW;    // floats per staged matrix (4160: the buffers behind stay 16-byte aligned)
constexpr int NFFB_MAX_LAYERS = 16;
constexpr int NFFB_WARPS = 16;
constexpr int NFFB_P = 4;               // points a warp walks through the layers together

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

// Optional second form of the output: the fp16 pair (h, l = (v - h) 2^11) the SDF pipeline's contraction consumes, plus a
// scaled second copy (the skip connection's columns) - written instead of the fp32 row when h != nullptr.
struct NffbPairOut {
    __half* h; __half* l; int ld, pad;
    __half* h2; __half* l2; int ld2, pad2; float scale2;
};

__device__ __forceinline__ void store_pair(__half* h, __half* l, long long o, float v) {
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half hv = __float2half_rn(v);
    h[o] = hv;
    l[o] = __float2half_rn((v - __half2float(hv)) * 2048.f);
}

__device__ __forceinline__ float warp_sum_all(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y[i][.] = b + sum_k Wt[k][o] v_i[k] for the lane's two outputs (o = lane, lane + 32) and the warp's NFFB_P points.
// The kernel is bound by shared-memory wavefronts, not FMAs: with one point per warp every k cost 3 of them (2 weight
// reads + 1 vector broadcast) for 2 FMAs.  Here the points' vectors are interleaved in shared memory (v[k][point]), so
// one 16-byte broadcast serves 4 points and the two weight reads are shared by them: 3 wavefronts for 8 FMAs.  Each
// point's sum still runs over k in order: bit-identical to the one-point form.
__device__ __forceinline__ void matvec4(const float* __restrict__ Wt, const float* __restrict__ bias, const float* v, int n_in,
                                        int W, int lane, float (&y0)[NFFB_P], float (&y1)[NFFB_P]) {
    float a0[NFFB_P] = {0.f, 0.f, 0.f, 0.f}, a1[NFFB_P] = {0.f, 0.f, 0.f, 0.f};
    const float* w = Wt + lane;
    const float4* v4 = reinterpret_cast<const float4*>(v);
#pragma unroll 4
    for (int k = 0; k < n_in; ++k) {
        const float4 vk = v4[k];
        const float w0 = w[k * NFFB_LDW], w1 = w[k * NFFB_LDW + 32];
        a0[0] = fmaf(w0, vk.x, a0[0]); a0[1] = fmaf(w0, vk.y, a0[1]); a0[2] = fmaf(w0, vk.z, a0[2]); a0[3] = fmaf(w0, vk.w, a0[3]);
        a1[0] = fmaf(w1, vk.x, a1[0]); a1[1] = fmaf(w1, vk.y, a1[1]); a1[2] = fmaf(w1, vk.z, a1[2]); a1[3] = fmaf(w1, vk.w, a1[3]);
    }
    const float b0 = lane < W ? bias[lane] : 0.f, b1 = lane + 32 < W ? bias[lane + 32] : 0.f;
#pragma unroll
    for (int i = 0; i < NFFB_P; ++i) { y0[i] = a0[i] + b0; y1[i] = a1[i] + b1; }
}

__global__ void __launch_bounds__(NFFB_WARPS * 32, 1)
nffb_encode_fwd_kernel(const NffbDev d, const float* __restrict__ x, long long n, int ldx, float* __restrict__ out, int ld_out,
                       const int* __restrict__ m_count, const NffbPairOut po) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    if (m_count != nullptr) { const long long c = *m_count; n = c < n ? c : n; }
    if (n <= 0) return;                     // gated tracer query: skip the weight staging
    const int W = d.width, NL = d.n_lin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // shared layout: transposed weights [mat][k][NFFB_MAX_W] (zero padded), biases [mat][NFFB_MAX_W], per-warp vectors
    const int n_mats = NL + 1 + (d.style ? 1 : 0);                 // SIREN layers, out layer, style transform
    float* s_w = smem;
    float* s_b = s_w + (size_t)n_mats * NFFB_MAT;
    float* s_v = s_b + n_mats * NFFB_MAX_W + warp * (NFFB_P * (2 * NFFB_MAX_W + 32));     // [z | e | chunk columns] x points
    for (int m = 0; m < n_mats; ++m) {
        const float* src_w = m < NL ? d.lin_w[m] : (m == NL ? d.out_w : d.sty_w);
        const float* src_b = m < NL ? d.lin_b[m] : (m == NL ? d.out_b : d.sty_b);
        const int n_in = m == 0 ? 3 : W;
        float* dw = s_w + (size_t)m * NFFB_MAT;
        // consecutive threads read consecutive k of one output row (coalesced; the strided form cost ~40 us per launch, most
        // of a 4096-point tracer query) and write the transposed element
        for (int i = threadIdx.x; i < NFFB_MAX_W * NFFB_MAX_W; i += blockDim.x) {
            const int o = i / NFFB_MAX_W, k = i - o * NFFB_MAX_W;
            dw[k * NFFB_LDW + o] = (k < n_in && o < W) ? src_w[o * n_in + k] : 0.f;
        }
        for (int i = threadIdx.x; i < NFFB_MAX_W; i += blockDim.x) s_b[m * NFFB_MAX_W + i] = i < W ? src_b[i] : 0.f;
    }
    __syncthreads();
    float* zs = s_v;                                   // current SIREN activations       [k][point]
    float* es = s_v + NFFB_P * NFFB_MAX_W;             // encoded chunk / (e + z)         [k][point]
    float* gs = s_v + 2 * NFFB_P * NFFB_MAX_W;         // grid columns the chunks are cut from [c][point]
    const GridDev& g = d.grid;
    const int C = g.n_fourier, F = g.n_feat;
    const int n_cols = (NL - 1) * d.chunk;                          // grid columns consumed by chunks 0 .. NL-2
    const int dch = d.chunk;
    const int head = d.include_input ? 2 * dch : 0;
    const float den = 2.f * d.bound;

    const long long n_groups = (n + NFFB_P - 1) / NFFB_P;
    const long long wstride = (long long)gridDim.x * NFFB_WARPS;
    for (long long grp = (long long)blockIdx.x * NFFB_WARPS + warp; grp < n_groups; grp += wstride) {
        const long long pbase = grp * NFFB_P;
        float u[NFFB_P][3];
#pragma unroll
        for (int i = 0; i < NFFB_P; ++i) {
            const long long p = pbase + i < n ? pbase + i : n - 1;          // tail: recompute the last point, never stored
            const float p0 = x[p * ldx + 0], p1 = x[p * ldx + 1], p2 = x[p * ldx + 2];
            u[i][0] = (p0 + d.bound) / den; u[i][1] = (p1 + d.bound) / den; u[i][2] = (p2 + d.bound) / den;
            if (lane < 3) zs[lane * NFFB_P + i] = (lane == 0 ? p0 : (lane == 1 ? p1 : p2)) / d.bound;
            // ---- grid columns (grid_enc(u)[:, 3:]): Fourier sin | cos | level features, arithmetic of hash_encode.cu
            for (int c = lane; c < n_cols; c += 32) {
                float v;
                if (c < 2 * C) {
                    const int j = c < C ? c : c - C;
                    float xp = __fmul_rn(__fmul_rn(u[i][0], 6.283185307179586f), g.B[j]);
                    xp = __fmaf_rn(__fmul_rn(u[i][1], 6.283185307179586f), g.B[C + j], xp);
                    xp = __fmaf_rn(__fmul_rn(u[i][2], 6.283185307179586f), g.B[2 * C + j], xp);
                    float sn, cs;
                    sincos_fast(xp, &sn, &cs);
                    v = c < C ? sn : cs;
                } else {
                    const int l = (c - 2 * C) / F, f = (c - 2 * C) - l * F;
                    const float r = g.res[l];
                    const uint32_t h = hash3(trunc_u32(__fmul_rn(u[i][0], r)), trunc_u32(__fmul_rn(u[i][1], r)), trunc_u32(__fmul_rn(u[i][2], r)));
                    v = __ldg(g.tables[l] + (size_t)wrap(h, g.rows[l], g.pow2mask[l], g.magic[l]) * F + f);
                }
                gs[c * NFFB_P + i] = v;
            }
        }
        __syncwarp();
        float f0[NFFB_P] = {0.f, 0.f, 0.f, 0.f}, f1[NFFB_P] = {0.f, 0.f, 0.f, 0.f};     // the lane's two output features per point
        for (int j = 0; j < NL; ++j) {
            float y0[NFFB_P], y1[NFFB_P];
            matvec4(s_w + (size_t)j * NFFB_MAT, s_b + j * NFFB_MAX_W, zs, j == 0 ? 3 : W, W, lane, y0, y1);
            float z0[NFFB_P], z1[NFFB_P];
#pragma unroll
            for (int i = 0; i < NFFB_P; ++i) { z0[i] = sinf(y0[i] * d.w0); z1[i] = sinf(y1[i] * d.w0); }
            __syncwarp();
            reinterpret_cast<float4*>(zs)[lane] = make_float4(z0[0], z0[1], z0[2], z0[3]);
            reinterpret_cast<float4*>(zs)[lane + 32] = make_float4(z1[0], z1[1], z1[2], z1[3]);
            if (j > 0) {
                // E = PositionalEncoding(chunk_{j-1}): [c | c | sin(b0 c) | cos(b0 c) | sin(b1 c) | ...].  Column o and column
                // o + dch share their argument (sin / cos of the same band x chunk value), and lanes of one warp would run
                // both the sinf and the cosf path (the band index changes every dch lanes): instead one lane evaluates
                // sincosf ONCE per (band, chunk column) for the 4 points and the row is assembled in shared memory
                // (dch x n_bands calls per point instead of 2 W divergent ones).
                const float* ch = gs + (j - 1) * dch * NFFB_P;
                float4* es4 = reinterpret_cast<float4*>(es);
                for (int o = lane; o < head; o += 32) es4[o] = reinterpret_cast<const float4*>(ch)[o % dch];
                for (int idx = lane; idx < dch * d.n_bands; idx += 32) {
                    const int m = idx / dch, jx = idx - m * dch;
                    const float4 c4 = reinterpret_cast<const float4*>(ch)[jx];
                    const float band = d.bands[m];
                    float sx, sy, sz, sw, cx, cy, cz, cw;
                    sincosf(__fmul_rn(c4.x, band), &sx, &cx);
                    sincosf(__fmul_rn(c4.y, band), &sy, &cy);
                    sincosf(__fmul_rn(c4.z, band), &sz, &cz);
                    sincosf(__fmul_rn(c4.w, band), &sw, &cw);
                    es4[head + (2 * m) * dch + jx] = make_float4(sx, sy, sz, sw);
                    es4[head + (2 * m + 1) * dch + jx] = make_float4(cx, cy, cz, cw);
                }
                for (int o = W + lane; o < NFFB_MAX_W; o += 32) es4[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncwarp();
                float e0[NFFB_P], e1[NFFB_P];
                if (!d.style) {
                    const float4 a4 = es4[lane], b4 = es4[lane + 32];
                    e0[0] = a4.x; e0[1] = a4.y; e0[2] = a4.z; e0[3] = a4.w;
                    e1[0] = b4.x; e1[1] = b4.y; e1[2] = b4.z; e1[3] = b4.w;
                    __syncwarp();
                } else {
                    float s0[NFFB_P], s1[NFFB_P];
                    matvec4(s_w + (size_t)(NL + 1) * NFFB_MAT, s_b + (NL + 1) * NFFB_MAX_W, es, W, W, lane, s0, s1);
#pragma unroll
                    for (int i = 0; i < NFFB_P; ++i) {
                        const float m0 = lane < W ? s0[i] : 0.f, m1 = lane + 32 < W ? s1[i] : 0.f;
                        const float mu = warp_sum_all(m0 + m1) / (float)W;
                        const float c0 = lane < W ? s0[i] - mu : 0.f, c1 = lane + 32 < W ? s1[i] - mu : 0.f;
                        const float var = warp_sum_all(c0 * c0 + c1 * c1) / (float)W;
                        const float inv = 1.f / sqrtf(var + d.eps);
                        e0[i] = c0 * inv; e1[i] = c1 * inv;
                    }
                    __syncwarp();
                }
                // out_layer is linear: sum (e + z) over the levels, apply it once after the loop (as the module path does)
#pragma unroll
                for (int i = 0; i < NFFB_P; ++i) { f0[i] += e0[i] + z0[i]; f1[i] += e1[i] + z1[i]; }
            }
            __syncwarp();
        }
        {
            reinterpret_cast<float4*>(es)[lane] = make_float4(f0[0], f0[1], f0[2], f0[3]);
            reinterpret_cast<float4*>(es)[lane + 32] = make_float4(f1[0], f1[1], f1[2], f1[3]);
            __syncwarp();
            float o0[NFFB_P], o1[NFFB_P];
            matvec4(s_w + (size_t)NL * NFFB_MAT, s_b + NL * NFFB_MAX_W, es, W, W, lane, o0, o1);
            // matvec4 added the bias once: sum_j (O u_j + o) = O sum_j u_j + (NL - 1) o
            const float b0 = lane < W ? s_b[NL * NFFB_MAX_W + lane] : 0.f, b1 = lane + 32 < W ? s_b[NL * NFFB_MAX_W + lane + 32] : 0.f;
            const float nb = (float)(NL - 2);
#pragma unroll
            for (int i = 0; i < NFFB_P; ++i) { f0[i] = fmaf(nb, b0, o0[i]); f1[i] = fmaf(nb, b1, o1[i]); }
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < NFFB_P; ++i) {
            if (pbase + i >= n) break;
            if (po.h != nullptr) {
                // the fp16 pair of the SDF pipeline's operand (+ the skip connection's scaled copy), pads zero-filled
                const long long o1 = (pbase + i) * (long long)po.ld, o2 = (pbase + i) * (long long)po.ld2;
                const __half zero = __float2half_rn(0.f);
                if (lane < 3) {
                    store_pair(po.h, po.l, o1 + lane, u[i][lane == 0 ? 0 : (lane == 1 ? 1 : 2)]);
                    if (po.h2) store_pair(po.h2, po.l2, o2 + lane, u[i][lane == 0 ? 0 : (lane == 1 ? 1 : 2)] * po.scale2);
                }
                if (lane < W) {
                    store_pair(po.h, po.l, o1 + 3 + lane, f0[i] / d.levels_div);
                    if (po.h2) store_pair(po.h2, po.l2, o2 + 3 + lane, f0[i] / d.levels_div * po.scale2);
                }
                if (lane + 32 < W) {
                    store_pair(po.h, po.l, o1 + 3 + lane + 32, f1[i] / d.levels_div);
                    if (po.h2) store_pair(po.h2, po.l2, o2 + 3 + lane + 32, f1[i] / d.levels_div * po.scale2);
                }
                for (int c = 3 + W + lane; c < 3 + W + po.pad; c += 32) { po.h[o1 + c] = zero; po.l[o1 + c] = zero; }
                if (po.h2) for (int c = 3 + W + lane; c < 3 + W + po.pad2; c += 32) { po.h2[o2 + c] = zero; po.l2[o2 + c] = zero; }
                continue;
            }
            float* orow = out + (pbase + i) * (long long)ld_out;
            if (lane < 3) orow[lane] = lane == 0 ? u[i][0] : (lane == 1 ? u[i][1] : u[i][2]);
            if (lane < W) orow[3 + lane] = f0[i] / d.levels_div;
            if (lane + 32 < W) orow[3 + lane + 32] = f1[i] / d.levels_div;
            for (int c = 3 + W + lane; c < ld_out; c += 32) orow[c] = 0.f;
        }
        __syncwarp();
    }
}

}  // namespace idrk

using namespace idrk;

static int nffb_encode_impl(const idrk_nffb_t* h, const float* x, int64_t n, int32_t ldx, float* out, int32_t ld_out,
                            const int32_t* m_count, const NffbPairOut& po, void* stream) {
    if (!h || !x || (!out && !po.h) || n < 0 || ldx < 3) return IDRK_E_ARG;
    NffbDev d;
    int rc = fill_grid(&h->grid, d.grid);
    if (rc) return rc;
    if (h->grid.frac_mode != IDRK_HASH_REFERENCE) return IDRK_E_UNSUP;
    if (h->width < 1 || h->width > NFFB_MAX_W || h->n_lin < 2 || h->n_lin > NFFB_MAX_LAYERS) return IDRK_E_UNSUP;
    if (h->n_bands < 0 || h->n_bands > 32 || h->chunk < 1 || !(h->bound > 0.f) || h->n_levels_div < 1) return IDRK_E_ARG;
    if (h->width != h->chunk * ((h->include_input ? 2 : 0) + 2 * h->n_bands)) return IDRK_E_ARG;
    const int n_cols = (h->n_lin - 1) * h->chunk;
    if (n_cols > 32 || n_cols > 2 * h->grid.n_fourier + h->grid.n_levels * h->grid.n_feat) return IDRK_E_UNSUP;
    if ((out && ld_out < 3 + h->width) || !h->out_w || !h->out_b) return IDRK_E_ARG;
    if (po.h && (!po.l || po.pad < 0 || po.ld < 3 + h->width + po.pad)) return IDRK_E_ARG;
    if (po.h2 && (!po.h || !po.l2 || po.pad2 < 0 || po.ld2 < 3 + h->width + po.pad2)) return IDRK_E_ARG;
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
    const size_t smem = ((size_t)n_mats * NFFB_MAT + (size_t)n_mats * NFFB_MAX_W +
                         (size_t)NFFB_WARPS * NFFB_P * (2 * NFFB_MAX_W + 32)) * sizeof(float);
    if (smem > 220 * 1024) return IDRK_E_UNSUP;
    IDRK_CUDA_TRY(cudaFuncSetAttribute(nffb_encode_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nffb_encode_fwd_kernel, NFFB_WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long grid = (long long)per_sm * sm_count();
    const long long need = (n + NFFB_WARPS * NFFB_P - 1) / (NFFB_WARPS * NFFB_P);
    if (grid > need) grid = need;
    IDRK_CUDA_TRY(launch_k(nffb_encode_fwd_kernel, dim3((unsigned)grid), dim3(NFFB_WARPS * 32), smem, (cudaStream_t)stream,
                           d, x, (long long)n, (int)ldx, out, (int)ld_out, m_count, po));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_nffb_encode_fwd(const idrk_nffb_t* h, const float* x, int64_t n, int32_t ldx, float* out, int32_t ld_out,
                                    const int32_t* m_count, void* stream) {
    if (!out) return IDRK_E_ARG;
    NffbPairOut po = {};
    return nffb_encode_impl(h, x, n, ldx, out, ld_out, m_count, po, stream);
}

extern "C" int idrk_nffb_encode_f16pair(const idrk_nffb_t* h, const float* x, int64_t n, int32_t ldx, const int32_t* m_count,
                                        void* out_h, void* out_l, int32_t ld_out, int32_t pad_cols,
                                        void* out_h2, void* out_l2, int32_t ld_out2, int32_t pad_cols2, float scale2, void* stream) {
    if (!out_h || !out_l) return IDRK_E_ARG;
    if ((out_h2 == nullptr) != (out_l2 == nullptr)) return IDRK_E_ARG;
    NffbPairOut po = {(__half*)out_h, (__half*)out_l, ld_out, pad_cols, (__half*)out_h2, (__half*)out_l2, ld_out2, pad_cols2, scale2};
    return nffb_encode_impl(h, x, n, ldx, nullptr, 0, m_count, po, stream);
}
