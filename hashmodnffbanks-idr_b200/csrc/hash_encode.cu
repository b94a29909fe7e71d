// K1 / K2: multi-resolution hash-grid encode, forward and backward, for sm_100a.
//
// Behaviour follows the reference's live PyTorch path
//   model/embeddings/hashGridEmbedding.py:32-40   hash_func  (primes {1, 3, 2654435761}, uint32, % T)
//   model/embeddings/hashGridEmbedding.py:81-102  _HashGridMLP.forward
//   model/embeddings/hashGridEmbedding.py:150-155 MultiResHashGridMLP.forward (Fourier prefix ++ levels)
//   model/embeddings/frequency_enc.py:63-67       FourierFeature.forward
// IDRK_HASH_REFERENCE reproduces it bit-exactly (the reference's fractional part is identically 0, so a
// level's output is the floor-corner row); IDRK_HASH_TRILINEAR is the 8-corner interpolation.
//
// Design (bound by the SM <-> L2 request rate: every random 8-byte table read or reduction costs a full 32-byte
// sector request, so the lever is requests per point, not bytes):
//   * a warp owns 32 consecutive points, a lane one (point, level) element at a time; every output float is
//     produced by the lane that stores it, in row-major order -> coalesced row segments with no shared-memory
//     staging, no block barrier, occupancy limited by registers only;
//   * stores are re-aligned to 8-byte words with one shuffle (the level columns start on an odd column), so each
//     output sector is written by one full request; KB independent gathers are issued before the first store;
//   * tables are read through the read-only path (ld.global.nc): coarse levels live in L1, the rest in the 126 MB L2;
//     persistent grid = (resident CTAs per SM) x (SM count);
//   * backward: table gradients go out as vector reductions (red.global.add.v2/v4.f32).  Levels whose whole table
//     fits a shared-memory budget are first accumulated per CTA in shared memory and flushed once (kills the
//     contention on tiny tables such as the reference configs' T = 32); dL/dx partials are shuffle-reduced per row.
#include <cuda_fp16.h>
#include "hash_common.cuh"
#include "sort_common.cuh"

namespace idrk {

struct GradDev {
    float* grad[IDRK_MAX_LEVELS];
    int small_off[IDRK_MAX_LEVELS];         // offset (floats) into the CTA's shared accumulator, -1 = global
    int small_total;                        // floats
    int any_grad;                           // 0 when only dL/dx is wanted
};

template <int F>
__device__ __forceinline__ void scatter_add(float* table, uint32_t idx, const float (&v)[F]) {
    float* p = table + (size_t)idx * F;
    if constexpr (F == 1) {
        atomicAdd(p, v[0]);
    } else if constexpr (F == 2) {
        red_add_v2(p, v[0], v[1]);
    } else if constexpr (F == 4) {
        red_add_v4(p, v[0], v[1], v[2], v[3]);
    } else {
        red_add_v4(p, v[0], v[1], v[2], v[3]);
        red_add_v4(p + 4, v[4], v[5], v[6], v[7]);
    }
}

// ------------------------------------------------------------------------------------------
// element-per-lane kernels
// ------------------------------------------------------------------------------------------
// A warp owns 32 consecutive points; a lane owns one (point, level) element at a time (flat index e = lane,
// lane + 32, ... over [32 points][L levels]; the Fourier prefix likewise over [32 points][C frequencies]).
// Every output float is produced by the lane that stores it, in row-major order, so the stores (and the dL/dy
// loads of the backward) are coalesced 128-byte row segments WITHOUT shared-memory staging: no block barrier,
// ~0.7 KB of shared memory per warp, occupancy limited by registers only.  When L (or C) divides 32 a lane's
// level (frequency) never changes, so its constants - resolution, rows, fastmod magic, table pointer, B column -
// live in registers for the whole kernel.
struct LevelC {                      // 48 bytes, read as 16-byte shared loads (the third only in IDRK_HASH_NGP kernels)
    float res; uint32_t rows, mask, soff;
    unsigned long long magic; const float* tab;
    uint32_t ngp_res, ngp_dense, pad0, pad1;
};

template <bool NGP = false>
__device__ __forceinline__ LevelC load_level(const LevelC* s_lev, int l) {
    const uint4 a = reinterpret_cast<const uint4*>(s_lev)[3 * l], b = reinterpret_cast<const uint4*>(s_lev)[3 * l + 1];
    LevelC c;
    c.res = __uint_as_float(a.x); c.rows = a.y; c.mask = a.z; c.soff = a.w;
    c.magic = ((unsigned long long)b.y << 32) | b.x;
    c.tab = reinterpret_cast<const float*>(((unsigned long long)b.w << 32) | b.z);
    c.ngp_res = 0; c.ngp_dense = 0; c.pad0 = 0; c.pad1 = 0;
    if constexpr (NGP) { const uint4 d = reinterpret_cast<const uint4*>(s_lev)[3 * l + 2]; c.ngp_res = d.x; c.ngp_dense = d.y; }
    return c;
}

// The 8 corner rows and the 3 interpolation weights of one (point, level) element in the 8-corner modes.
//   IDRK_HASH_TRILINEAR: the reference's hash (primes 1, 3, 2654435761) on floor(x * res).
//   IDRK_HASH_NGP: tiny-cuda-nn's grid (pos = fma(x, scale, 0.5); dense stride index while the level fits its table,
//   else the coherent prime hash 1, 2654435761, 805459861); corner bit 0 = x, bit 1 = y, bit 2 = z in both.
template <int MODE, bool FAST>
__device__ __forceinline__ void cell_of(const LevelC& lc, float x0, float x1, float x2, uint32_t& c0, uint32_t& c1, uint32_t& c2,
                                        float& w0, float& w1, float& w2) {
    float s0, s1, s2;
    if constexpr (MODE == IDRK_HASH_NGP) { s0 = fmaf(x0, lc.res, 0.5f); s1 = fmaf(x1, lc.res, 0.5f); s2 = fmaf(x2, lc.res, 0.5f); }
    else { s0 = __fmul_rn(x0, lc.res); s1 = __fmul_rn(x1, lc.res); s2 = __fmul_rn(x2, lc.res); }
    const float f0 = floorf(s0), f1 = floorf(s1), f2 = floorf(s2);
    w0 = s0 - f0; w1 = s1 - f1; w2 = s2 - f2;
    if constexpr (FAST) { c0 = (uint32_t)__float2int_rz(f0); c1 = (uint32_t)__float2int_rz(f1); c2 = (uint32_t)__float2int_rz(f2); }
    else { c0 = trunc_u32(f0); c1 = trunc_u32(f1); c2 = trunc_u32(f2); }
}

template <int MODE>
__device__ __forceinline__ void rows_of_cell(const LevelC& lc, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t (&idx)[8]) {
    if constexpr (MODE == IDRK_HASH_NGP) {
        const bool dense = lc.ngp_dense != 0;
        const uint32_t R = lc.ngp_res, py = dense ? R : 2654435761u, pz = dense ? R * R : 805459861u;
        const uint32_t hx0 = c0, hx1 = c0 + 1u, hy0 = c1 * py, hy1 = hy0 + py, hz0 = c2 * pz, hz1 = hz0 + pz;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t a = (k & 1) ? hx1 : hx0, b = (k & 2) ? hy1 : hy0, c = (k & 4) ? hz1 : hz0;
            idx[k] = wrap(dense ? a + b + c : a ^ b ^ c, lc.rows, lc.mask, lc.magic);
        }
    } else {
        // hash3 is an xor of three per-dimension terms: form the 2 x 3 terms once, xor per corner
        const uint32_t hx0 = c0, hx1 = c0 + 1u, hy0 = c1 * 3u, hy1 = hy0 + 3u;
        const uint32_t hz0 = c2 * 2654435761u, hz1 = hz0 + 2654435761u;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            idx[k] = wrap(((k & 1) ? hx1 : hx0) ^ ((k & 2) ? hy1 : hy0) ^ ((k & 4) ? hz1 : hz0), lc.rows, lc.mask, lc.magic);
    }
}

template <int MODE, bool FAST>
__device__ __forceinline__ void corner_rows(const LevelC& lc, float x0, float x1, float x2, uint32_t (&idx)[8],
                                            float& w0, float& w1, float& w2, uint32_t& c0_out) {
    uint32_t c0, c1, c2;
    cell_of<MODE, FAST>(lc, x0, x1, x2, c0, c1, c2, w0, w1, w2);
    c0_out = c0;
    rows_of_cell<MODE>(lc, c0, c1, c2, idx);
}
template <bool FAST> __device__ __forceinline__ uint32_t trunc_sel(float v) {
    if constexpr (FAST) return (uint32_t)__float2int_rz(v); else return trunc_u32(v);
}

// x-neighbour pairing (8-corner backward, F == 2).  hash3's x term has prime 1, so for an EVEN floor coordinate c0
// the two x-neighbours hash to h and h ^ 1, and with an even row count (h mod T) keeps them in one aligned pair of
// rows = one 16-byte slot of the gradient table.  Such an element issues 4 reductions (RED.128) instead of 8; an odd
// c0 keeps a separate 8-byte reduction for its x+1 corner.  The table-gradient backward is bound by L2 reduction
// operations (one per touched sector), so this is -25 % of them on average: 0.87 -> 1.16-1.21 Gpts/s at T = 2^19,
// 0.65 -> 0.79 at T = 2^22 (profiles/r01_hash_pair_ab.txt).  The same pairing of the forward's gathers (LDG.128)
// was measured and did not pay: no change while the tables sit in L2, slower for T >= 2^22.

// one (point, level) element of the forward
template <int F, int MODE, bool FAST>
__device__ __forceinline__ void fwd_element(const LevelC& lc, float x0, float x1, float x2, float (&acc)[F]) {
    const float s0 = __fmul_rn(x0, lc.res), s1 = __fmul_rn(x1, lc.res), s2 = __fmul_rn(x2, lc.res);
    if constexpr (MODE == IDRK_HASH_REFERENCE) {
        const uint32_t h = hash3(trunc_sel<FAST>(s0), trunc_sel<FAST>(s1), trunc_sel<FAST>(s2));
        gather<F>(lc.tab, wrap(h, lc.rows, lc.mask, lc.magic), acc);
    } else {
        float w0, w1, w2;
        uint32_t idx[8], c0;
        corner_rows<MODE, FAST>(lc, x0, x1, x2, idx, w0, w1, w2, c0);
        float v[8][F];
#pragma unroll
        for (int k = 0; k < 8; ++k) gather<F>(lc.tab, idx[k], v[k]);
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float wk = ((k & 1) ? w0 : 1.f - w0) * ((k & 2) ? w1 : 1.f - w1) * ((k & 4) ? w2 : 1.f - w2);
#pragma unroll
            for (int f = 0; f < F; ++f) acc[f] = fmaf(wk, v[k][f], acc[f]);
        }
    }
}

template <int F>
__device__ __forceinline__ void store_feat(float* o, const float (&v)[F], bool vec) {
    if (vec) {
        if constexpr (F == 2) { __stcs(reinterpret_cast<float2*>(o), make_float2(v[0], v[1])); return; }
        if constexpr (F == 4) { __stcs(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3])); return; }
        if constexpr (F == 8) {
            __stcs(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
            __stcs(reinterpret_cast<float4*>(o) + 1, make_float4(v[4], v[5], v[6], v[7]));
            return;
        }
    }
#pragma unroll
    for (int f = 0; f < F; ++f) __stcs(o + f, v[f]);
}

// Generic tile walk (any L, ragged tiles, coordinates outside the int32 range): one element at a time.
template <int F, int MODE>
__device__ __forceinline__ void fwd_levels_generic(const LevelC* s_lev, const float4* xs, int L, int lane, int rows_here,
                                                   float* __restrict__ orow0, int ld_out, bool vec) {
    const int total = rows_here * L;
    for (int e = lane; e < total; e += 32) {
        const int row = e / L, l = e - row * L;
        const LevelC lc = load_level<MODE == IDRK_HASH_NGP>(s_lev, l);
        const float4 xv = xs[row];
        float acc[F];
        fwd_element<F, MODE, false>(lc, xv.x, xv.y, xv.z, acc);
        store_feat<F>(orow0 + (long long)__float_as_int(xv.w) * ld_out + l * F, acc, vec);
    }
}

// Full 32-row tile with L | 32: the lane's level l = lane % L and its constants are fixed for the whole kernel; the
// tile takes L passes of 32 / L rows.  KB passes are batched so that KB (8 * KB in the 8-corner mode) independent
// gathers are in flight before the first store waits on one - that is what hides the L2 / HBM latency.
// SHIFT (F == 2, level columns starting on an odd column, a pad column after the last level): a lane's two features
// straddle an 8-byte boundary, so it stores (own f1, next level's f0) as one aligned float2 - every 32-byte sector
// of the row is then written by one full request instead of two half-filled ones.
template <int F, int MODE, bool SHIFT>
__device__ __forceinline__ void fwd_levels_full(const LevelC& lc, const float4* xs, int L, int lane,
                                                float* __restrict__ orow0, int ld_out, bool vec, bool tail = true) {
    constexpr int KB = (MODE == IDRK_HASH_REFERENCE) ? (F <= 2 ? 4 : 2) : (F <= 2 ? 2 : 1);
    const int rstep = 32 / L, row0 = lane / L, l = lane - row0 * L;
    float* __restrict__ o0 = orow0 + l * F;
    for (int pass = 0; pass < L; pass += KB) {
        float acc[KB][F];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            if (pass + k < L) {
                const float4 xv = xs[row0 + (pass + k) * rstep];
                fwd_element<F, MODE, true>(lc, xv.x, xv.y, xv.z, acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            if (pass + k < L) {
                // row of the output buffer relative to the tile's first point: re-read here (one shared-memory word) rather
                // than carried across the gathers - registers held there cost gathers in flight (T >= 2^22 is latency-bound)
                const int ro = __float_as_int(reinterpret_cast<const float*>(xs + row0 + (pass + k) * rstep)[3]);
                float* o = o0 + (long long)ro * ld_out;
                if constexpr (SHIFT) {
                    float nx = __shfl_down_sync(0xffffffffu, acc[k][0], 1);
                    if (l == L - 1) nx = 0.f;                                  // the pad column
                    if (l == L - 1 && !tail) __stcs(o + 1, acc[k][1]);         // a later level window owns the next column
                    else __stcs(reinterpret_cast<float2*>(o + 1), make_float2(acc[k][1], nx));
                    if (l == 0) __stcs(o, acc[k][0]);
                } else {
                    store_feat<F>(o, acc[k], vec);
                }
            }
        }
    }
}

template <int F, int MODE, int WARPS, int MINB = 0>
__global__ void __launch_bounds__(WARPS * 32, MINB)
hash_encode_fwd_elem_kernel(const GridDev g, const float* __restrict__ x, long long n, int ldx,
                            float* __restrict__ out, int ld_out, uint32_t* __restrict__ idx_dbg,
                            const int* __restrict__ m_count, const int* __restrict__ perm) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    extern __shared__ uint4 smem_u4[];
    if (m_count != nullptr) { const long long c = *m_count; n = c < n ? c : n; }
    if (n <= 0) return;                     // gated tracer query: nothing to encode
    const int C = g.n_fourier, L = g.n_levels;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LevelC* s_lev = reinterpret_cast<LevelC*>(smem_u4);                               // [L]
    float4* xs = reinterpret_cast<float4*>(smem_u4 + 3 * L) + warp * 32;              // [32] per warp
    float* s_B = reinterpret_cast<float*>(smem_u4 + 3 * L + WARPS * 32);              // [3][C]
    for (int i = threadIdx.x; i < L; i += WARPS * 32) {
        LevelC c; c.res = g.res[i]; c.rows = g.rows[i]; c.mask = g.pow2mask[i]; c.soff = 0; c.magic = g.magic[i]; c.tab = g.tables[i];
        c.ngp_res = g.ngp_res[i]; c.ngp_dense = g.ngp_dense[i]; c.pad0 = 0; c.pad1 = 0;
        s_lev[i] = c;
    }
    for (int i = threadIdx.x; i < 3 * C; i += WARPS * 32) s_B[i] = g.B[i];
    __syncthreads();
    const int pre = g.pre_cols;                             // prefix columns (+ the levels of earlier windows)
    const int fa = F >= 4 ? 4 : F;                          // floats a vector store must be aligned to
    const bool al8 = (ld_out & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & 7u) == 0;
    const bool vec = F > 1 && (pre % fa) == 0 && (ld_out % fa) == 0 && (reinterpret_cast<uintptr_t>(out) % (4 * fa)) == 0;
    const bool tail = g.tail != 0;
    const bool has_pad = ld_out > g.width || !tail;         // a column follows this launch's last level
    const bool l_fixed = L > 0 && (32 % L) == 0;            // lane <-> level is fixed
    const bool shift = F == 2 && l_fixed && (pre & 1) && has_pad && al8;
    const bool c_fixed = C >= 2 && (32 % C) == 0 && al8;    // lane <-> frequency is fixed, aligned pair stores
    float res_max = 0.f;
    for (int l = 0; l < L; ++l) res_max = fmaxf(res_max, fabsf(g.res[l]));
    LevelC lc = load_level<MODE == IDRK_HASH_NGP>(s_lev, L > 0 ? lane % L : 0);
    // Fourier prefix, c_fixed: column 3 + j is even for odd j, so odd lanes store the aligned pairs (s_j, s_j+1) and
    // (c_j, c_j+1) - the partner value comes from the next lane - lane j = 0 stores (x2, s_0) and one even lane
    // (x0, x1): every sector of the prefix is written in full 8-byte words.
    const int j = C > 0 ? lane % C : 0, jr0 = C > 0 ? lane / C : 0, jrstep = C > 0 ? 32 / C : 0;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if (C > 0) { b0 = s_B[j]; b1 = s_B[C + j]; b2 = s_B[2 * C + j]; }
    const bool j_odd = j & 1, j_last = j == C - 1, j_zero = j == 0, j_x = j == (C > 2 ? 2 : 1);
    const bool st_a = j_odd || j_zero || j_x, st_b = j_odd && !j_last;
    const int off_a = j_odd ? 3 + j : (j_zero ? 2 : 0);

    const long long n_wtiles = (n + 31) / 32;
    const long long wstride = (long long)gridDim.x * WARPS;
    for (long long wt = (long long)blockIdx.x * WARPS + warp; wt < n_wtiles; wt += wstride) {
        const long long p0 = wt * 32, p = p0 + lane;
        const int rows_here = (int)min(32LL, n - p0);
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;
        // `perm`: the i-th point processed is point perm[i] of the caller's buffers (its x row, its output row), so an
        // unordered batch is walked in Z-order without being copied.  xs[].w = that row relative to the tile's first point.
        int roff = lane;
        if (lane < rows_here) {
            if (perm != nullptr) roff = (int)((long long)perm[p] - p0);
            const float* xr = x + (p0 + roff) * (long long)ldx;
            x0 = xr[0]; x1 = xr[1]; x2 = xr[2];
        }
        __syncwarp();
        xs[lane] = make_float4(x0, x1, x2, __int_as_float(roff));
        __syncwarp();
        float* orow0 = out + p0 * (long long)ld_out;
        const bool full = rows_here == 32;
        if (C > 0 && c_fixed && full) {
#pragma unroll 2
            for (int pass = 0; pass < C; ++pass) {
                const int row = jr0 + pass * jrstep;
                const float4 xv = xs[row];
                float xp = __fmul_rn(__fmul_rn(xv.x, 6.283185307179586f), b0);
                xp = __fmaf_rn(__fmul_rn(xv.y, 6.283185307179586f), b1, xp);
                xp = __fmaf_rn(__fmul_rn(xv.z, 6.283185307179586f), b2, xp);
                float sn, cs;
                sincos_fast(xp, &sn, &cs);
                const float sn_n = __shfl_down_sync(0xffffffffu, sn, 1), cs_n = __shfl_down_sync(0xffffffffu, cs, 1);
                const float cs_0 = __shfl_sync(0xffffffffu, cs, lane & ~(C - 1));
                float* o = orow0 + (long long)__float_as_int(xv.w) * ld_out;
                const float a0 = j_odd ? sn : (j_zero ? xv.z : xv.x);
                const float a1 = j_odd ? (j_last ? cs_0 : sn_n) : (j_zero ? sn : xv.y);
                if (st_a) __stcs(reinterpret_cast<float2*>(o + off_a), make_float2(a0, a1));
                if (st_b) __stcs(reinterpret_cast<float2*>(o + 3 + C + j), make_float2(cs, cs_n));
                if (j_last) __stcs(o + 3 + C + j, cs);
                if (C == 2 && j_zero) __stcs(reinterpret_cast<float2*>(o), make_float2(xv.x, xv.y));   // C == 2: lane 1 is both "odd" and j_x
            }
        } else if (C > 0) {
            for (int i = lane; i < 3 * rows_here; i += 32) {
                const int r = i / 3, c = i - 3 * r;
                __stcs(orow0 + (long long)__float_as_int(xs[r].w) * ld_out + c, reinterpret_cast<const float*>(xs + r)[c]);
            }
            const int total = rows_here * C;
            for (int e = lane; e < total; e += 32) {
                const int row = e / C, jj = e - row * C;
                const float4 xv = xs[row];
                float xp = __fmul_rn(__fmul_rn(xv.x, 6.283185307179586f), s_B[jj]);
                xp = __fmaf_rn(__fmul_rn(xv.y, 6.283185307179586f), s_B[C + jj], xp);
                xp = __fmaf_rn(__fmul_rn(xv.z, 6.283185307179586f), s_B[2 * C + jj], xp);
                float sn, cs;
                sincos_fast(xp, &sn, &cs);
                float* o = orow0 + (long long)__float_as_int(xv.w) * ld_out + 3 + jj;
                __stcs(o, sn);
                __stcs(o + C, cs);
            }
        }
        bool pad_done = false;
        if (L > 0) {
            // .long() of a scaled coordinate fits 32 bits for every level unless a point is astronomically far out
            const float amax = fmaxf(fabsf(x0), fmaxf(fabsf(x1), fabsf(x2))) * res_max;
            const bool fast = __all_sync(0xffffffffu, amax < 2147483520.f);
            if (fast && full && l_fixed) {
                if (F == 2 && shift) { fwd_levels_full<F, MODE, F == 2>(lc, xs, L, lane, orow0 + pre, ld_out, vec, tail); pad_done = true; }
                else fwd_levels_full<F, MODE, false>(lc, xs, L, lane, orow0 + pre, ld_out, vec);
            } else {
                fwd_levels_generic<F, MODE>(s_lev, xs, L, lane, rows_here, orow0 + pre, ld_out, vec);
            }
        }
        if (tail && lane < rows_here)
            for (int c = g.width + (pad_done ? 1 : 0); c < ld_out; ++c) __stcs(orow0 + (long long)roff * ld_out + c, 0.f);

        if (idx_dbg != nullptr) {               // debug / parity output: table row of all 8 corners
            for (int e = lane; e < rows_here * L; e += 32) {
                const int row = e / L, l = e - row * L;
                const float4 xv = xs[row];
                const float r = g.res[l];
                const float s0 = __fmul_rn(xv.x, r), s1 = __fmul_rn(xv.y, r), s2 = __fmul_rn(xv.z, r);
                uint32_t c0, c1, c2;
                if constexpr (MODE == IDRK_HASH_NGP) {
                    const LevelC lcd = load_level<true>(s_lev, l);
                    uint32_t id8[8], cc; float q0, q1, q2;
                    corner_rows<MODE, false>(lcd, xv.x, xv.y, xv.z, id8, q0, q1, q2, cc);
                    for (int k = 0; k < 8; ++k) idx_dbg[((p0 + row) * L + l) * 8 + k] = id8[k];
                    continue;
                }
                if constexpr (MODE == IDRK_HASH_REFERENCE) { c0 = trunc_u32(s0); c1 = trunc_u32(s1); c2 = trunc_u32(s2); }
                else { c0 = trunc_u32(floorf(s0)); c1 = trunc_u32(floorf(s1)); c2 = trunc_u32(floorf(s2)); }
                for (int k = 0; k < 8; ++k)
                    idx_dbg[((p0 + row) * L + l) * 8 + k] =
                        wrap(hash3(c0 + (k & 1), c1 + ((k >> 1) & 1), c2 + ((k >> 2) & 1)), g.rows[l], g.pow2mask[l], g.magic[l]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
// Backward, element-per-lane (see the forward above): lane (row, l) reads its dL/dy features straight from the row
// (aligned float2 + one shuffle when the level columns start on an odd column), forms the table row(s) again and
// issues the vector reductions.  dL/dx partials are reduced across a row's lanes with xor-shuffles when L (C)
// divides 32, otherwise through shared-memory float atomics on a per-warp [32][3] tile.
template <int F>
__device__ __forceinline__ void scatter_sel(float* gtab, float* s_acc, uint32_t soff, uint32_t idx, const float (&v)[F]) {
    if (soff != 0xffffffffu) {
#pragma unroll
        for (int f = 0; f < F; ++f) atomicAdd(s_acc + soff + idx * F + f, v[f]);
    } else {
        scatter_add<F>(gtab, idx, v);
    }
}

template <int F, int MODE, bool FAST, bool WANT_DX, bool PAIR = false>
__device__ __forceinline__ void bwd_element(const LevelC& lc, float* gtab, float* s_acc, float x0, float x1, float x2,
                                            const float (&gy)[F], float (&d)[3], bool pair_lane = false) {
    const float s0 = __fmul_rn(x0, lc.res), s1 = __fmul_rn(x1, lc.res), s2 = __fmul_rn(x2, lc.res);
    if constexpr (MODE == IDRK_HASH_REFERENCE) {
        if (gtab == nullptr) return;
        const uint32_t h = hash3(trunc_sel<FAST>(s0), trunc_sel<FAST>(s1), trunc_sel<FAST>(s2));
        scatter_sel<F>(gtab, s_acc, lc.soff, wrap(h, lc.rows, lc.mask, lc.magic), gy);
    } else {
        float w0, w1, w2;
        uint32_t idx[8], c0;
        corner_rows<MODE, FAST>(lc, x0, x1, x2, idx, w0, w1, w2, c0);
        float t[8][F];
        if constexpr (WANT_DX) {
#pragma unroll
            for (int k = 0; k < 8; ++k) gather<F>(lc.tab, idx[k], t[k]);
        }
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float a0 = (k & 1) ? w0 : 1.f - w0, a1 = (k & 2) ? w1 : 1.f - w1, a2 = (k & 4) ? w2 : 1.f - w2;
            if (gtab != nullptr) {
                const float wk = a0 * a1 * a2;
                float v[F];
#pragma unroll
                for (int f = 0; f < F; ++f) v[f] = wk * gy[f];
                if (PAIR && F == 2 && pair_lane) {
                    // paired x-neighbours (see the note above fwd_element): the x corner goes out as a 16-byte reduction on its
                    // aligned row pair, carrying the x+1 corner in the other half when c0 is even (zeros otherwise)
                    if ((k & 1) == 0) {
                        const float wn = ((c0 & 1u) == 0) ? w0 * a1 * a2 : 0.f;
                        const float nx = wn * gy[0], ny = wn * gy[1];
                        const bool o = idx[k] & 1u;
                        red_add_v4(gtab + (size_t)(idx[k] & ~1u) * F, o ? nx : v[0], o ? ny : v[1], o ? v[0] : nx, o ? v[1] : ny);
                    } else if (c0 & 1u) {
                        red_add_v2(gtab + (size_t)idx[k] * F, v[0], v[1]);
                    }
                } else {
                    scatter_sel<F>(gtab, s_acc, lc.soff, idx[k], v);
                }
            }
            if constexpr (WANT_DX) {
                float dot = 0.f;
#pragma unroll
                for (int f = 0; f < F; ++f) dot = fmaf(t[k][f], gy[f], dot);
                d0 = fmaf(((k & 1) ? 1.f : -1.f) * a1 * a2, dot, d0);
                d1 = fmaf(((k & 2) ? 1.f : -1.f) * a0 * a2, dot, d1);
                d2 = fmaf(((k & 4) ? 1.f : -1.f) * a0 * a1, dot, d2);
            }
        }
        if constexpr (WANT_DX) { d[0] = d0 * lc.res; d[1] = d1 * lc.res; d[2] = d2 * lc.res; }
    }
}

// In-lane run aggregation (8-corner modes, table gradients only).  On the full-tile path a lane owns ONE level and walks the
// points of its warp's tile range in order, so when consecutive points fall into the same cell of that level (spatially
// ordered input: samples along a ray, Morton-sorted points; always true for the coarse levels of dense batches) their
// 8 corner contributions are summed in registers and leave as ONE set of reductions when the cell changes.  The table-
// gradient backward is bound by the number of L2 reduction operations (193 G/s device-wide whatever their width,
// scripts/probes/atomics_probe.cu), so every merged element saves its 4-8 of them; an element that starts a new cell costs
// what it cost before (one compare + the same reductions, issued one element later).
template <int F>
struct CellAcc {
    uint32_t c0, c1, c2;
    int valid;
    float acc[8][F];
};

template <int F, int MODE, bool PAIR>
__device__ __forceinline__ void flush_cell(const LevelC& lc, float* gtab, float* s_acc, const CellAcc<F>& st, bool pair_lane) {
    uint32_t idx[8];
    rows_of_cell<MODE>(lc, st.c0, st.c1, st.c2, idx);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (PAIR && F == 2 && pair_lane) {
            if ((k & 1) == 0) {                       // same pairing rule as bwd_element
                const bool even = (st.c0 & 1u) == 0;
                const float nx = even ? st.acc[k + 1][0] : 0.f, ny = even ? st.acc[k + 1][F - 1] : 0.f;
                const bool o = idx[k] & 1u;
                red_add_v4(gtab + (size_t)(idx[k] & ~1u) * F, o ? nx : st.acc[k][0], o ? ny : st.acc[k][F - 1],
                           o ? st.acc[k][0] : nx, o ? st.acc[k][F - 1] : ny);
            } else if (st.c0 & 1u) {
                red_add_v2(gtab + (size_t)idx[k] * F, st.acc[k][0], st.acc[k][F - 1]);
            }
        } else {
            scatter_sel<F>(gtab, s_acc, lc.soff, idx[k], st.acc[k]);
        }
    }
}

// one (point, level) element of the aggregated table-gradient backward
template <int F, int MODE, bool PAIR>
__device__ __forceinline__ void bwd_element_agg(const LevelC& lc, float* gtab, float* s_acc, float x0, float x1, float x2,
                                                const float (&gy)[F], CellAcc<F>& st, bool pair_lane) {
    uint32_t c0, c1, c2;
    float w0, w1, w2;
    cell_of<MODE, true>(lc, x0, x1, x2, c0, c1, c2, w0, w1, w2);
    const bool same = st.valid && c0 == st.c0 && c1 == st.c1 && c2 == st.c2;
    if (!same) {
        if (st.valid) flush_cell<F, MODE, PAIR>(lc, gtab, s_acc, st, pair_lane);
        st.c0 = c0; st.c1 = c1; st.c2 = c2; st.valid = 1;
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int f = 0; f < F; ++f) st.acc[k][f] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float wk = ((k & 1) ? w0 : 1.f - w0) * ((k & 2) ? w1 : 1.f - w1) * ((k & 4) ? w2 : 1.f - w2);
#pragma unroll
        for (int f = 0; f < F; ++f) st.acc[k][f] = fmaf(wk, gy[f], st.acc[k][f]);
    }
}

// adds (v0, v1, v2) of every lane into dxs[row]: xor-shuffle over the aligned group of `grp` lanes that share the row
// (grp = L or C when it divides 32), else shared-memory atomics.
__device__ __forceinline__ void row_reduce_add(float* dxs, int row, bool active, int grp, bool grp_ok, int lane,
                                               float v0, float v1, float v2) {
    if (grp_ok) {
        for (int off = grp >> 1; off > 0; off >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, off);
            v1 += __shfl_xor_sync(0xffffffffu, v1, off);
            v2 += __shfl_xor_sync(0xffffffffu, v2, off);
        }
        if (active && (lane & (grp - 1)) == 0) { dxs[4 * row] += v0; dxs[4 * row + 1] += v1; dxs[4 * row + 2] += v2; }
    } else if (active) {
        atomicAdd(dxs + 4 * row, v0); atomicAdd(dxs + 4 * row + 1, v1); atomicAdd(dxs + 4 * row + 2, v2);
    }
}

// Generic tile walk (any L, ragged tiles, coordinates outside the int32 range): one element at a time.
template <int F, int MODE, bool WANT_DX>
__device__ __forceinline__ void bwd_levels_generic(const LevelC* s_lev, float* const* s_grad, float* s_acc, const float4* xs,
                                                   float* dxs, int L, int lane, int rows_here,
                                                   const float* __restrict__ drow0, int ld_dy) {
    const int total = rows_here * L;
    for (int e = lane; e < total; e += 32) {
        const int row = e / L, l = e - row * L;
        const LevelC lc = load_level<MODE == IDRK_HASH_NGP>(s_lev, l);
        const float4 xv = xs[row];
        const float* o = drow0 + (long long)__float_as_int(xv.w) * ld_dy + l * F;
        float gy[F];
#pragma unroll
        for (int f = 0; f < F; ++f) gy[f] = __ldg(o + f);
        float d[3] = {0.f, 0.f, 0.f};
        bwd_element<F, MODE, false, WANT_DX>(lc, s_grad[l], s_acc, xv.x, xv.y, xv.z, gy, d);
        if constexpr (WANT_DX && MODE != IDRK_HASH_REFERENCE) {
            atomicAdd(dxs + 4 * row, d[0]); atomicAdd(dxs + 4 * row + 1, d[1]); atomicAdd(dxs + 4 * row + 2, d[2]);
        }
    }
}

// Full 32-row tile with L | 32 (see fwd_levels_full).  SHIFT: dL/dy features are read as the aligned pair
// (own f1, next level's f0); the lane's f0 arrives from the previous lane by shuffle.
template <int F, int MODE, bool WANT_DX, bool SHIFT, bool PAIR = false>
__device__ __forceinline__ void bwd_levels_full(const LevelC& lc, float* gtab, float* s_acc, const float4* xs, float* dxs,
                                                int L, int lane, const float* __restrict__ drow0, int ld_dy, bool pair_lane = false) {
    constexpr int KB = (MODE == IDRK_HASH_REFERENCE) ? (F <= 2 ? 4 : 2) : (WANT_DX ? 1 : 2);
    const int rstep = 32 / L, row0 = lane / L, l = lane - row0 * L;
    const float* __restrict__ o0 = drow0 + l * F;
    for (int pass = 0; pass < L; pass += KB) {
        float gy[KB][F];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            if (pass + k < L) {
                const float* o = o0 + (long long)__float_as_int(xs[row0 + (pass + k) * rstep].w) * ld_dy;
                if constexpr (SHIFT) {
                    const float2 pr = __ldcs(reinterpret_cast<const float2*>(o + 1));
                    float first = 0.f;
                    if (l == 0) first = __ldcs(o);
                    const float up = __shfl_up_sync(0xffffffffu, pr.y, 1);
                    gy[k][0] = (l == 0) ? first : up;
                    gy[k][1] = pr.x;
                } else {
#pragma unroll
                    for (int f = 0; f < F; ++f) gy[k][f] = __ldg(o + f);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            if (pass + k < L) {
                const int row = row0 + (pass + k) * rstep;
                const float4 xv = xs[row];
                float d[3] = {0.f, 0.f, 0.f};
                bwd_element<F, MODE, true, WANT_DX, PAIR>(lc, gtab, s_acc, xv.x, xv.y, xv.z, gy[k], d, pair_lane);
                if constexpr (WANT_DX && MODE != IDRK_HASH_REFERENCE) row_reduce_add(dxs, row, true, L, true, lane, d[0], d[1], d[2]);
            }
        }
    }
}

// Full 32-row tile, aggregated table-gradient form (no dL/dx): same walk as bwd_levels_full, the lane's open cell lives
// in `st` across passes AND tiles.
template <int F, int MODE, bool SHIFT, bool PAIR>
__device__ __forceinline__ void bwd_levels_full_agg(const LevelC& lc, float* gtab, float* s_acc, const float4* xs, int L, int lane,
                                                    const float* __restrict__ drow0, int ld_dy, bool pair_lane, CellAcc<F>& st) {
    constexpr int KB = 2;
    const int rstep = 32 / L, row0 = lane / L, l = lane - row0 * L;
    const float* __restrict__ o0 = drow0 + l * F;
    for (int pass = 0; pass < L; pass += KB) {
        float gy[KB][F];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            if (pass + k < L) {
                const float* o = o0 + (long long)__float_as_int(xs[row0 + (pass + k) * rstep].w) * ld_dy;
                if constexpr (SHIFT) {
                    const float2 pr = __ldcs(reinterpret_cast<const float2*>(o + 1));
                    float first = 0.f;
                    if (l == 0) first = __ldcs(o);
                    const float up = __shfl_up_sync(0xffffffffu, pr.y, 1);
                    gy[k][0] = (l == 0) ? first : up;
                    gy[k][1] = pr.x;
                } else {
#pragma unroll
                    for (int f = 0; f < F; ++f) gy[k][f] = __ldg(o + f);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            if (pass + k < L) {
                const float4 xv = xs[row0 + (pass + k) * rstep];
                bwd_element_agg<F, MODE, PAIR>(lc, gtab, s_acc, xv.x, xv.y, xv.z, gy[k], st, pair_lane);
            }
        }
    }
}

template <int F, int MODE, int WARPS, bool AGG = false>
__global__ void __launch_bounds__(WARPS * 32)
hash_encode_bwd_elem_kernel(const GridDev g, const GradDev gd, const float* __restrict__ x, long long n, int ldx,
                            const float* __restrict__ dy, int ld_dy, float* __restrict__ dx, const int* __restrict__ perm) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    extern __shared__ uint4 smem_u4[];
    const int C = g.n_fourier, L = g.n_levels;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LevelC* s_lev = reinterpret_cast<LevelC*>(smem_u4);                               // [L]
    float4* xs = reinterpret_cast<float4*>(smem_u4 + 3 * L) + warp * 32;              // [32] per warp
    float* dxs = reinterpret_cast<float*>(smem_u4 + 3 * L + WARPS * 32) + warp * 128; // [32][4] per warp
    float** s_grad = reinterpret_cast<float**>(smem_u4 + 3 * L + 2 * WARPS * 32);     // [L] (padded to even)
    float* s_B = reinterpret_cast<float*>(s_grad + ((L + 1) & ~1));                   // [3][C]
    float* s_acc = s_B + ((3 * C + 3) & ~3);                                          // [small_total]
    for (int i = threadIdx.x; i < L; i += WARPS * 32) {
        LevelC c; c.res = g.res[i]; c.rows = g.rows[i]; c.mask = g.pow2mask[i];
        c.soff = gd.small_off[i] >= 0 ? (uint32_t)gd.small_off[i] : 0xffffffffu; c.magic = g.magic[i]; c.tab = g.tables[i];
        c.ngp_res = g.ngp_res[i]; c.ngp_dense = g.ngp_dense[i]; c.pad0 = 0; c.pad1 = 0;
        s_lev[i] = c;
        s_grad[i] = gd.grad[i];
    }
    for (int i = threadIdx.x; i < 3 * C; i += WARPS * 32) s_B[i] = g.B[i];
    for (int i = threadIdx.x; i < gd.small_total; i += WARPS * 32) s_acc[i] = 0.f;
    __syncthreads();
    const int pre = g.pre_cols;
    const bool has_pad = ld_dy > g.width || !g.tail;
    const bool l_fixed = L > 0 && (32 % L) == 0;
    const bool shift = F == 2 && l_fixed && (pre & 1) && has_pad && (ld_dy & 1) == 0 && (reinterpret_cast<uintptr_t>(dy) & 7u) == 0;
    const LevelC lc = load_level<MODE == IDRK_HASH_NGP>(s_lev, L > 0 ? lane % L : 0);
    float* gtab = L > 0 ? s_grad[lane % L] : nullptr;
    const bool want_dx = dx != nullptr;
    // paired reductions need the gradient table (not the value table) on a 16-byte boundary
    constexpr bool PAIRK = MODE == IDRK_HASH_TRILINEAR && F == 2;      // compile-time: other kernels carry no pairing code
    bool pair_lane = false;                                            // per lane = per level on the full-tile path
    if constexpr (PAIRK)
        pair_lane = g.pair_x && L > 0 && gd.any_grad && (lc.rows & 1u) == 0 && lc.soff == 0xffffffffu &&
                    (reinterpret_cast<uintptr_t>(gtab) & 15u) == 0;
    float res_max = 0.f;
    for (int l = 0; l < L; ++l) res_max = fmaxf(res_max, fabsf(g.res[l]));
    const int jstep = C > 0 ? 32 % C : 0, jrstep = C > 0 ? 32 / C : 0;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if (C > 0) { const int j = lane % C; b0 = s_B[j]; b1 = s_B[C + j]; b2 = s_B[2 * C + j]; }

    const long long n_wtiles = (n + 31) / 32;
    long long wstride = (long long)gridDim.x * WARPS;
    long long wt_begin = (long long)blockIdx.x * WARPS + warp, wt_end = n_wtiles;
    // run aggregation (8-corner modes, table gradients without dL/dx): a warp walks a CONTIGUOUS range of tiles so that
    // a lane sees the points in input order; the strided assignment stays for everything else
    constexpr bool AGGK = AGG && MODE != IDRK_HASH_REFERENCE && F == 2;      // its own instantiation: 124 vs 80 registers
    const bool agg = AGGK && !want_dx && gd.any_grad && l_fixed;
    CellAcc<AGGK ? F : 1> st;
    st.valid = 0; st.c0 = st.c1 = st.c2 = 0;
    if (agg) {
        const long long per = (n_wtiles + wstride - 1) / wstride;
        wt_begin = wt_begin * per;
        wt_end = min(n_wtiles, wt_begin + per);
        wstride = 1;
    }
    for (long long wt = wt_begin; wt < wt_end; wt += wstride) {
        const long long p0 = wt * 32, p = p0 + lane;
        const int rows_here = (int)min(32LL, n - p0);
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;
        int roff = lane;                    // see the forward kernel: row of x / dy / dx relative to the tile's first point
        if (lane < rows_here) {
            if (perm != nullptr) roff = (int)((long long)perm[p] - p0);
            const float* xr = x + (p0 + roff) * (long long)ldx;
            x0 = xr[0]; x1 = xr[1]; x2 = xr[2];
        }
        const float* drow0 = dy + p0 * (long long)ld_dy;
        __syncwarp();
        xs[lane] = make_float4(x0, x1, x2, __int_as_float(roff));
        if (want_dx) {
            float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (C > 0 && lane < rows_here) {           // identity columns of the prefix
                const float* o = drow0 + (long long)roff * ld_dy;
                d0.x = __ldg(o); d0.y = __ldg(o + 1); d0.z = __ldg(o + 2);
            }
            reinterpret_cast<float4*>(dxs)[lane] = d0;
        }
        __syncwarp();
        if (want_dx && C > 0) {
            // d/dx of [sin(2 pi x B), cos(2 pi x B)]
            int row = lane / C, j = lane - row * C;
            const int total = rows_here * C;
            for (int base = 0; base < total; base += 32) {
                const bool act = base + lane < total;
                float v0 = 0.f, v1 = 0.f, v2 = 0.f;
                const int r_here = row;
                if (act) {
                    const float4 xv = xs[row];
                    float xp = __fmul_rn(__fmul_rn(xv.x, 6.283185307179586f), b0);
                    xp = __fmaf_rn(__fmul_rn(xv.y, 6.283185307179586f), b1, xp);
                    xp = __fmaf_rn(__fmul_rn(xv.z, 6.283185307179586f), b2, xp);
                    float sn, cs;
                    sincos_fast(xp, &sn, &cs);
                    const float* o = drow0 + (long long)__float_as_int(xv.w) * ld_dy + 3 + j;
                    const float dxp = (__ldg(o) * cs - __ldg(o + C) * sn) * 6.283185307179586f;
                    v0 = dxp * b0; v1 = dxp * b1; v2 = dxp * b2;
                }
                row_reduce_add(dxs, r_here & 31, act, C, jstep == 0, lane, v0, v1, v2);
                row += jrstep;
                if (jstep != 0) {
                    j += jstep;
                    if (j >= C) { j -= C; ++row; }
                    b0 = s_B[j]; b1 = s_B[C + j]; b2 = s_B[2 * C + j];
                }
            }
            if (jstep != 0) { const int j0 = lane % C; b0 = s_B[j0]; b1 = s_B[C + j0]; b2 = s_B[2 * C + j0]; }
        }
        __syncwarp();
        if (L > 0 && (MODE != IDRK_HASH_REFERENCE || gd.any_grad)) {
            const float amax = fmaxf(fabsf(x0), fmaxf(fabsf(x1), fabsf(x2))) * res_max;
            const bool fast = __all_sync(0xffffffffu, amax < 2147483520.f);
            if constexpr (AGGK) {
                if (agg && !(fast && rows_here == 32) && st.valid) {           // leaving the full-tile path: close the open cell
                    flush_cell<F, MODE, PAIRK>(lc, gtab, s_acc, st, pair_lane);
                    st.valid = 0;
                }
            }
            if (AGGK && agg && fast && rows_here == 32) {
                if constexpr (AGGK) {
                    if (shift) bwd_levels_full_agg<F, MODE, true, PAIRK>(lc, gtab, s_acc, xs, L, lane, drow0 + pre, ld_dy, pair_lane, st);
                    else       bwd_levels_full_agg<F, MODE, false, PAIRK>(lc, gtab, s_acc, xs, L, lane, drow0 + pre, ld_dy, pair_lane, st);
                }
            } else if (fast && rows_here == 32 && l_fixed) {
                if (F == 2 && shift) {
                    if (want_dx) bwd_levels_full<F, MODE, true, F == 2, PAIRK>(lc, gtab, s_acc, xs, dxs, L, lane, drow0 + pre, ld_dy, pair_lane);
                    else         bwd_levels_full<F, MODE, false, F == 2, PAIRK>(lc, gtab, s_acc, xs, dxs, L, lane, drow0 + pre, ld_dy, pair_lane);
                } else {
                    if (want_dx) bwd_levels_full<F, MODE, true, false, PAIRK>(lc, gtab, s_acc, xs, dxs, L, lane, drow0 + pre, ld_dy, pair_lane);
                    else         bwd_levels_full<F, MODE, false, false, PAIRK>(lc, gtab, s_acc, xs, dxs, L, lane, drow0 + pre, ld_dy, pair_lane);
                }
            } else {
                if (want_dx) bwd_levels_generic<F, MODE, true>(s_lev, s_grad, s_acc, xs, dxs, L, lane, rows_here, drow0 + pre, ld_dy);
                else         bwd_levels_generic<F, MODE, false>(s_lev, s_grad, s_acc, xs, dxs, L, lane, rows_here, drow0 + pre, ld_dy);
            }
        }
        if (want_dx) {
            __syncwarp();
            if (lane < rows_here) {
                float* dr = dx + (p0 + roff) * 3;
                dr[0] = dxs[4 * lane]; dr[1] = dxs[4 * lane + 1]; dr[2] = dxs[4 * lane + 2];
            }
        }
    }
    if constexpr (AGGK) {
        if (agg && st.valid) flush_cell<F, MODE, PAIRK>(lc, gtab, s_acc, st, pair_lane);
    }
    // flush the CTA-local accumulators of the small tables
    if (gd.small_total > 0) {
        __syncthreads();
        for (int l = 0; l < L; ++l) {
            const int soff = gd.small_off[l];
            if (soff < 0) continue;
            const int cnt = (int)g.rows[l] * F;
            for (int i = threadIdx.x; i < cnt; i += WARPS * 32) {
                const float v = s_acc[soff + i];
                if (v != 0.f) atomicAdd(gd.grad[l] + i, v);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------
// encode -> fp16 pair (tracer queries)
// ------------------------------------------------------------------------------------------
// The ray tracer's SDF queries feed the fp16-pair contraction (gemm.cu): instead of writing the fp32 embedding and
// splitting it in a second launch, one thread per (point, output column) forms the column with exactly the arithmetic
// of the element kernels above (same helpers, same operation order -> bit-identical values) and stores it as the pair
// h = fp16(v), l = fp16((v - h) * 2^11) - optionally a second scaled copy (the skip connection's 1/sqrt(2) embedding
// columns inside the layer-4 operand).  Row counts here are <= 32 K and the row is 27-67 columns, so the redundant
// re-evaluation of a level by its F column threads is cheaper than a second launch on the query's critical path.
template <int F, int MODE>
__global__ void hash_encode_pair_kernel(const GridDev g, const float* __restrict__ x, long long n, int ldx,
                                        const int* __restrict__ m_count, __half* __restrict__ h, __half* __restrict__ l,
                                        int ldo, int pad_cols, __half* __restrict__ h2, __half* __restrict__ l2, int ldo2,
                                        int pad_cols2, float scale2) {
    pdl_wait();
    pdl_trigger();
    if (m_count != nullptr) { const long long c = *m_count; n = c < n ? c : n; }
    if (n <= 0) return;
    const int C = g.n_fourier, L = g.n_levels;
    const int pre = C > 0 ? 3 + 2 * C : 0;
    const int width = pre + L * F;
    const int pmax = (h2 != nullptr && pad_cols2 > pad_cols) ? pad_cols2 : pad_cols;
    const int wt = width + pmax;
    const long long total = n * (long long)wt;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / wt;
        const int c = (int)(i - r * wt);
        float v = 0.f;
        if (c < width) {
            const float x0 = x[r * ldx], x1 = x[r * ldx + 1], x2 = x[r * ldx + 2];
            if (c < pre) {
                if (c < 3) {
                    v = c == 0 ? x0 : (c == 1 ? x1 : x2);
                } else {
                    const int j = c - 3 < C ? c - 3 : c - 3 - C;
                    float xp = __fmul_rn(__fmul_rn(x0, 6.283185307179586f), g.B[j]);
                    xp = __fmaf_rn(__fmul_rn(x1, 6.283185307179586f), g.B[C + j], xp);
                    xp = __fmaf_rn(__fmul_rn(x2, 6.283185307179586f), g.B[2 * C + j], xp);
                    float sn, cs;
                    sincos_fast(xp, &sn, &cs);
                    v = c - 3 < C ? sn : cs;
                }
            } else {
                const int lev = (c - pre) / F, f = (c - pre) - lev * F;
                LevelC lc;
                lc.res = g.res[lev]; lc.rows = g.rows[lev]; lc.mask = g.pow2mask[lev]; lc.soff = 0; lc.magic = g.magic[lev];
                lc.tab = g.tables[lev]; lc.ngp_res = g.ngp_res[lev]; lc.ngp_dense = g.ngp_dense[lev]; lc.pad0 = 0; lc.pad1 = 0;
                float acc[F];
                fwd_element<F, MODE, false>(lc, x0, x1, x2, acc);
                v = acc[0];
#pragma unroll
                for (int k = 1; k < F; ++k) if (f == k) v = acc[k];
            }
        }
        if (c < width + pad_cols) {
            const float u = fminf(fmaxf(v, -65504.f), 65504.f);
            const __half hv = __float2half_rn(u);
            h[r * ldo + c] = hv;
            l[r * ldo + c] = __float2half_rn((u - __half2float(hv)) * 2048.f);
        }
        if (h2 != nullptr && c < width + pad_cols2) {
            const float u = fminf(fmaxf(v * scale2, -65504.f), 65504.f);
            const __half hv = __float2half_rn(u);
            h2[r * ldo2 + c] = hv;
            l2[r * ldo2 + c] = __float2half_rn((u - __half2float(hv)) * 2048.f);
        }
    }
}

// ------------------------------------------------------------------------------------------
// deterministic table-gradient pass
// ------------------------------------------------------------------------------------------
// Floating-point reductions into the tables land in scheduling order, so two runs of the passes above differ in the last
// bits.  This pass fixes the order: every (point, level, corner) contribution gets the key (level, table row) and its
// own index as value; a STABLE radix sort (csrc/point_sort.cu) groups the contributions of a row in ascending index
// order; one thread then sums a row's segment front to back and is the only writer of that row.  Same arithmetic per
// contribution (w_k * dL/dy), bit-identical results run to run and independent of the grid size.  Cost: 8 B of key /
// value traffic per contribution and pass (1 - 4 passes: the key has ceil(log2 L) + ceil(log2 max rows) bits) and serial
// segments on the coarse levels - meant for reproducible training steps (10^3 - 10^5 points), not for 2^24-point batches.
struct DetGrad { float* grad[IDRK_MAX_LEVELS]; };

template <int MODE>
__device__ __forceinline__ LevelC level_consts(const GridDev& g, int lev) {
    LevelC lc;
    lc.res = g.res[lev]; lc.rows = g.rows[lev]; lc.mask = g.pow2mask[lev]; lc.soff = 0; lc.magic = g.magic[lev];
    lc.tab = g.tables[lev]; lc.ngp_res = g.ngp_res[lev]; lc.ngp_dense = g.ngp_dense[lev]; lc.pad0 = 0; lc.pad1 = 0;
    return lc;
}

template <int MODE>
__global__ void det_keys_kernel(const GridDev g, const float* __restrict__ x, long long n, int ldx, int row_bits,
                                uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    pdl_wait();
    pdl_trigger();
    constexpr int G = MODE == IDRK_HASH_REFERENCE ? 1 : 8;
    const int L = g.n_levels;
    const long long total = n * L;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long p = e / L;
        const int l = (int)(e - p * L);
        const LevelC lc = level_consts<MODE>(g, l);
        const float x0 = x[p * ldx], x1 = x[p * ldx + 1], x2 = x[p * ldx + 2];
        if constexpr (MODE == IDRK_HASH_REFERENCE) {
            const float s0 = __fmul_rn(x0, lc.res), s1 = __fmul_rn(x1, lc.res), s2 = __fmul_rn(x2, lc.res);
            const uint32_t row = wrap(hash3(trunc_u32(s0), trunc_u32(s1), trunc_u32(s2)), lc.rows, lc.mask, lc.magic);
            keys[e] = ((uint32_t)l << row_bits) | row;
            vals[e] = (uint32_t)e;
        } else {
            uint32_t idx[8], c0;
            float w0, w1, w2;
            corner_rows<MODE, false>(lc, x0, x1, x2, idx, w0, w1, w2, c0);
#pragma unroll
            for (int k = 0; k < 8; ++k) { keys[e * G + k] = ((uint32_t)l << row_bits) | idx[k]; vals[e * G + k] = (uint32_t)(e * G + k); }
        }
    }
}

template <int F, int MODE>
__global__ void det_reduce_kernel(const GridDev g, const DetGrad gd, const float* __restrict__ x, int ldx,
                                  const float* __restrict__ dy, int ld_dy, long long m, int row_bits,
                                  const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals) {
    pdl_wait();
    pdl_trigger();
    constexpr int G = MODE == IDRK_HASH_REFERENCE ? 1 : 8;
    const int L = g.n_levels, pre = g.pre_cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t key = keys[i];
        if (i > 0 && keys[i - 1] == key) continue;                   // not the head of a row's segment
        const int l = (int)(key >> row_bits);
        const uint32_t row = key & ((1u << row_bits) - 1u);
        const LevelC lc = level_consts<MODE>(g, l);
        float acc[F];
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = 0.f;
        for (long long j = i; j < m && keys[j] == key; ++j) {        // ascending contribution index: the sort is stable
            const uint32_t id = vals[j];
            const long long e = id / G;
            const int k = (int)(id - e * G);
            const long long p = e / L;
            const float* gy = dy + p * (long long)ld_dy + pre + l * F;
            float wk = 1.f;
            if constexpr (MODE != IDRK_HASH_REFERENCE) {
                uint32_t c0, c1, c2;
                float w0, w1, w2;
                cell_of<MODE, false>(lc, x[p * ldx], x[p * ldx + 1], x[p * ldx + 2], c0, c1, c2, w0, w1, w2);
                wk = ((k & 1) ? w0 : 1.f - w0) * ((k & 2) ? w1 : 1.f - w1) * ((k & 4) ? w2 : 1.f - w2);
            }
#pragma unroll
            for (int f = 0; f < F; ++f) acc[f] += __fmul_rn(wk, __ldg(gy + f));
        }
        float* o = gd.grad[l] + (size_t)row * F;
#pragma unroll
        for (int f = 0; f < F; ++f) o[f] += acc[f];                  // the row's only writer
    }
}

static int ceil_log2(unsigned long long v) { int b = 0; while ((1ull << b) < v) ++b; return b; }

// Persistent grid = resident CTAs per SM x SM count.  Also pins the L1 / shared-memory split to what the kernel
// needs: the random table reads live off L1 + L2, and with the default preference the driver keeps the large
// shared-memory carve-out left behind by a preceding GEMM launch (measured: -35 % on the 8-corner forward).
template <typename K>
static int persistent_grid(K kernel, int threads, size_t smem, long long n_tiles) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int pct = (int)(((smem + 1024) * (size_t)per_sm * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    long long g = (long long)per_sm * sm_count();
    if (g > n_tiles) g = n_tiles;
    if (g < 1) g = 1;
    return (int)g;
}

template <int F, int MODE, int WARPS>
static int launch_fwd(const GridDev& g, const float* x, long long n, int ldx, float* out, int ld_out,
                      uint32_t* idx_dbg, const int* m_count, const int* perm, cudaStream_t st) {
    const size_t smem = ((size_t)3 * g.n_levels + WARPS * 32) * 16 + (size_t)3 * g.n_fourier * sizeof(float);
    auto kern = hash_encode_fwd_elem_kernel<F, MODE, WARPS>;
    // 8-corner, F = 2: capping the kernel at 4 resident CTAs per SM (118 registers instead of 80) lets ptxas keep both
    // batched elements' 16 gathers in flight: +5 % with L2-resident tables, +27 % / +12 % at T = 2^22 / 2^24 where every
    // gather is a DRAM sector read (profiles/r01_hash_pair_ab.txt)
    if constexpr (F == 2 && MODE != IDRK_HASH_REFERENCE) kern = hash_encode_fwd_elem_kernel<F, MODE, WARPS, 4>;
    const int grid = persistent_grid(kern, WARPS * 32, smem, (n + WARPS * 32 - 1) / (WARPS * 32));
    IDRK_CUDA_TRY(launch_k(kern, dim3(grid), dim3(WARPS * 32), smem, st, g, x, n, ldx, out, ld_out, idx_dbg, m_count, perm));
    IDRK_LAUNCH_CHECK();
    return 0;
}

template <int F, int MODE, int WARPS>
static int launch_bwd(const GridDev& g, const GradDev& gd, const float* x, long long n, int ldx, const float* dy,
                      int ld_dy, float* dx, bool ordered, const int* perm, cudaStream_t st) {
    const int L = g.n_levels, C = g.n_fourier;
    const size_t smem = ((size_t)3 * L + 2 * WARPS * 32) * 16 + (size_t)((L + 1) & ~1) * 8 +
                        (size_t)((3 * C + 3) & ~3) * 4 + (size_t)gd.small_total * 4;
    auto kern = hash_encode_bwd_elem_kernel<F, MODE, WARPS>;
    // spatially ordered input (caller's hint): the run-aggregating instantiation, see CellAcc
    if constexpr (MODE != IDRK_HASH_REFERENCE && F == 2) {
        if (ordered && g.agg_runs && dx == nullptr && gd.any_grad) kern = hash_encode_bwd_elem_kernel<F, MODE, WARPS, true>;
    }
    IDRK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = persistent_grid(kern, WARPS * 32, smem, (n + WARPS * 32 - 1) / (WARPS * 32));
    IDRK_CUDA_TRY(launch_k(kern, dim3(grid), dim3(WARPS * 32), smem, st, g, gd, x, n, ldx, dy, ld_dy, dx, perm));
    IDRK_LAUNCH_CHECK();
    return 0;
}

#define IDRK_DISPATCH_F_MODE(CALL)                                                             \
    switch (g.n_feat * 2 + mode) {                                                             \
        case 1 * 2 + 0: return CALL(1, IDRK_HASH_REFERENCE);                                   \
        case 1 * 2 + 1: return CALL(1, IDRK_HASH_TRILINEAR);                                   \
        case 2 * 2 + 0: return CALL(2, IDRK_HASH_REFERENCE);                                   \
        case 2 * 2 + 1: return CALL(2, IDRK_HASH_TRILINEAR);                                   \
        case 4 * 2 + 0: return CALL(4, IDRK_HASH_REFERENCE);                                   \
        case 4 * 2 + 1: return CALL(4, IDRK_HASH_TRILINEAR);                                   \
        case 8 * 2 + 0: return CALL(8, IDRK_HASH_REFERENCE);                                   \
        case 8 * 2 + 1: return CALL(8, IDRK_HASH_TRILINEAR);                                   \
        case 2 * 2 + 2: return CALL(2, IDRK_HASH_NGP);                                         \
        default: return IDRK_E_UNSUP;                                                          \
    }

}  // namespace idrk

using namespace idrk;

extern "C" int idrk_version(void) { return 4; }   // 4: idrk_gemm_p16 / idrk_split_p16 / idrk_weight_norm_fwd_p16 / idrk_act_bwd_p16; 3: idrk_camera_rays / idrk_idr_loss / idrk_scale3 (render_glue.cu); 2: idrk_epilogue_f16_t gained dot_w / dot_out / ld_dot, IDRK_HASH_NGP, idrk_hash_encode_f16pair

extern "C" int idrk_device_sm_count(int* out_sms) {
    if (!out_sms) return IDRK_E_ARG;
    int dev = 0;
    IDRK_CUDA_TRY(cudaGetDevice(&dev));
    IDRK_CUDA_TRY(cudaDeviceGetAttribute(out_sms, cudaDevAttrMultiProcessorCount, dev));
    return 0;
}

extern "C" int idrk_hash_encode_fwd(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx,
                                    float* out, int32_t ld_out, uint32_t* idx_debug, const int32_t* m_count,
                                    const int32_t* perm, void* stream) {
    GridDev g;
    int rc = fill_grid(h_grid, g);
    if (rc) return rc;
    if (n < 0 || ldx < 3 || x == nullptr || out == nullptr || ld_out < g.width) return IDRK_E_ARG;
    if (perm != nullptr && (idx_debug != nullptr || n >= (1LL << 31))) return IDRK_E_ARG;
    if (n == 0) return 0;
    if ((ld_out & 3) == 0 && !aligned16(out)) return IDRK_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int mode = h_grid->frac_mode;
#define CALL(F, M) launch_fwd<F, M, 4>(g, x, n, ldx, out, ld_out, idx_debug, m_count, perm, st)
    auto run = [&](const GridDev& g) -> int { IDRK_DISPATCH_F_MODE(CALL) };
#undef CALL
    int win[IDRK_MAX_LEVELS + 1];
    const int nw = (idx_debug != nullptr || mode == IDRK_HASH_REFERENCE) ? 1 : level_groups(g, n, 512LL << 20, win);
    if (nw <= 1) return run(g);
    for (int i = 0; i < nw; ++i) {                         // tables beyond L2: one launch per level window
        rc = run(level_window(g, win[i], win[i + 1]));
        if (rc) return rc;
    }
    return 0;
}

template <int F, int MODE>
static int launch_pair(const GridDev& g, const float* x, long long n, int ldx, const int* m_count, __half* h, __half* l,
                       int ldo, int pad_cols, __half* h2, __half* l2, int ldo2, int pad_cols2, float scale2, cudaStream_t st) {
    const int pmax = (h2 && pad_cols2 > pad_cols) ? pad_cols2 : pad_cols;
    const long long total = n * (long long)(g.width + pmax);
    long long b = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    IDRK_CUDA_TRY(launch_k(hash_encode_pair_kernel<F, MODE>, dim3((int)b), dim3(256), 0, st, g, x, n, ldx, m_count, h, l, ldo,
                           pad_cols, h2, l2, ldo2, pad_cols2, scale2));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_hash_encode_f16pair(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx,
                                        const int32_t* m_count, void* h, void* l, int32_t ld_out, int32_t pad_cols,
                                        void* h2, void* l2, int32_t ld_out2, int32_t pad_cols2, float scale2, void* stream) {
    GridDev g;
    int rc = fill_grid(h_grid, g);
    if (rc) return rc;
    if (n < 0 || ldx < 3 || x == nullptr || h == nullptr || l == nullptr || pad_cols < 0 || ld_out < g.width + pad_cols) return IDRK_E_ARG;
    if ((h2 == nullptr) != (l2 == nullptr)) return IDRK_E_ARG;
    if (h2 && (pad_cols2 < 0 || ld_out2 < g.width + pad_cols2)) return IDRK_E_ARG;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int mode = h_grid->frac_mode;
    if (g.n_feat != 2) return IDRK_E_UNSUP;
#define PAIR(M) launch_pair<2, M>(g, x, n, ldx, m_count, (__half*)h, (__half*)l, ld_out, pad_cols, (__half*)h2, (__half*)l2, ld_out2, pad_cols2, scale2, st)
    if (mode == IDRK_HASH_REFERENCE) return PAIR(IDRK_HASH_REFERENCE);
    if (mode == IDRK_HASH_TRILINEAR) return PAIR(IDRK_HASH_TRILINEAR);
    return PAIR(IDRK_HASH_NGP);
#undef PAIR
}

extern "C" int idrk_hash_encode_bwd(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx,
                                    const float* dy, int32_t ld_dy, float* const* h_grad_tables, float* dx,
                                    int32_t flags, const int32_t* perm, void* stream) {
    GridDev g;
    int rc = fill_grid(h_grid, g);
    if (rc) return rc;
    if (n < 0 || ldx < 3 || x == nullptr || dy == nullptr || ld_dy < g.width) return IDRK_E_ARG;
    if (h_grad_tables == nullptr && dx == nullptr) return IDRK_E_ARG;
    if (n == 0) return 0;
    const int budget = 8192;                    // floats of CTA-local accumulators (32 KB)
    const size_t align = (g.n_feat >= 4) ? 16 : 4 * (size_t)g.n_feat;
    for (int l = 0; l < g.n_levels && h_grad_tables != nullptr; ++l) {
        if (h_grad_tables[l] == nullptr) return IDRK_E_ARG;
        if (reinterpret_cast<uintptr_t>(h_grad_tables[l]) % align) return IDRK_E_ALIGN;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int mode = h_grid->frac_mode;
    const bool ordered = (flags & IDRK_HASH_BWD_ORDERED) != 0 || perm != nullptr;      // a permutation is a Z-order by contract
    if (perm != nullptr && n >= (1LL << 31)) return IDRK_E_ARG;
    auto run = [&](const GridDev& g, int l0) -> int {
        GradDev gd;
        gd.small_total = 0;
        gd.any_grad = h_grad_tables != nullptr ? 1 : 0;
        for (int l = 0; l < IDRK_MAX_LEVELS; ++l) { gd.grad[l] = nullptr; gd.small_off[l] = -1; }
        for (int l = 0; l < g.n_levels && h_grad_tables != nullptr; ++l) {
            gd.grad[l] = h_grad_tables[l0 + l];
            const long long cnt = (long long)g.rows[l] * g.n_feat;
            if (cnt <= 2048 && gd.small_total + cnt <= budget) { gd.small_off[l] = gd.small_total; gd.small_total += (int)cnt; }
        }
#define CALL(F, M) launch_bwd<F, M, 4>(g, gd, x, n, ldx, dy, ld_dy, dx, ordered, perm, st)
        IDRK_DISPATCH_F_MODE(CALL)
#undef CALL
    };
    // level windows (tables beyond L2) for the table-gradient pass; with dL/dx the single launch stays (dx is written once)
    int win[IDRK_MAX_LEVELS + 1];
    const int nw = (dx != nullptr || h_grad_tables == nullptr) ? 1 : level_groups(g, n, 0, win);
    if (nw <= 1) return run(g, 0);
    for (int i = 0; i < nw; ++i) {
        rc = run(level_window(g, win[i], win[i + 1]), win[i]);
        if (rc) return rc;
    }
    return 0;
}

// -- deterministic table gradients ---------------------------------------------------------
static int det_layout(const GridDev& g, int mode, long long n, long long& m, int& row_bits, int& key_bits) {
    const int G = mode == IDRK_HASH_REFERENCE ? 1 : 8;
    unsigned long long max_rows = 1;
    for (int l = 0; l < g.n_levels; ++l) max_rows = g.rows[l] > max_rows ? g.rows[l] : max_rows;
    row_bits = ceil_log2(max_rows);
    if (row_bits < 1) row_bits = 1;
    key_bits = row_bits + ceil_log2((unsigned long long)g.n_levels);
    m = n * g.n_levels * G;
    if (key_bits > 32 || m >= (1LL << 31)) return IDRK_E_UNSUP;
    return 0;
}

static long long up256(long long v) { return (v + 255) / 256 * 256; }

extern "C" int idrk_hash_encode_bwd_det_workspace(const idrk_hashgrid_t* h_grid, int64_t n, int64_t* out_bytes) {
    GridDev g;
    int rc = fill_grid(h_grid, g);
    if (rc) return rc;
    if (!out_bytes || n < 0 || g.n_levels == 0) return IDRK_E_ARG;
    long long m; int rb, kb;
    rc = det_layout(g, h_grid->frac_mode, n, m, rb, kb);
    if (rc) return rc;
    const long long sc = radix_sort_scratch_bytes(m);
    if (sc < 0) return IDRK_E_UNSUP;
    *out_bytes = 4 * up256(m * 4) + sc;
    return 0;
}

extern "C" int idrk_hash_encode_bwd_det(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx, const float* dy,
                                        int32_t ld_dy, float* const* h_grad_tables, void* workspace, int64_t workspace_bytes,
                                        void* stream) {
    GridDev g;
    int rc = fill_grid(h_grid, g);
    if (rc) return rc;
    if (n < 0 || ldx < 3 || !x || !dy || ld_dy < g.width || !h_grad_tables || !workspace || g.n_levels == 0) return IDRK_E_ARG;
    if (n == 0) return 0;
    long long m; int row_bits, key_bits;
    rc = det_layout(g, h_grid->frac_mode, n, m, row_bits, key_bits);
    if (rc) return rc;
    int64_t need = 0;
    rc = idrk_hash_encode_bwd_det_workspace(h_grid, n, &need);
    if (rc) return rc;
    if (workspace_bytes < need) return IDRK_E_ARG;
    if (!aligned16(workspace)) return IDRK_E_ALIGN;
    DetGrad gd;
    for (int l = 0; l < IDRK_MAX_LEVELS; ++l) gd.grad[l] = nullptr;
    for (int l = 0; l < g.n_levels; ++l) { if (!h_grad_tables[l]) return IDRK_E_ARG; gd.grad[l] = h_grad_tables[l]; }
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace;
    uint32_t* keys[2] = {(uint32_t*)w, (uint32_t*)(w + up256(m * 4))};
    uint32_t* vals[2] = {(uint32_t*)(w + 2 * up256(m * 4)), (uint32_t*)(w + 3 * up256(m * 4))};
    void* scratch = w + 4 * up256(m * 4);
    const int mode = h_grid->frac_mode;
    long long kb = (n * g.n_levels + 255) / 256;
    if (kb > 8LL * sm_count()) kb = 8LL * sm_count();
    if (mode == IDRK_HASH_REFERENCE) IDRK_CUDA_TRY(launch_k(det_keys_kernel<IDRK_HASH_REFERENCE>, dim3((unsigned)kb), dim3(256), 0, st, g, x, (long long)n, (int)ldx, row_bits, keys[0], vals[0]));
    else if (mode == IDRK_HASH_TRILINEAR) IDRK_CUDA_TRY(launch_k(det_keys_kernel<IDRK_HASH_TRILINEAR>, dim3((unsigned)kb), dim3(256), 0, st, g, x, (long long)n, (int)ldx, row_bits, keys[0], vals[0]));
    else IDRK_CUDA_TRY(launch_k(det_keys_kernel<IDRK_HASH_NGP>, dim3((unsigned)kb), dim3(256), 0, st, g, x, (long long)n, (int)ldx, row_bits, keys[0], vals[0]));
    int cur = 0;
    rc = radix_sort_pairs(keys, vals, nullptr, m, key_bits, scratch, st, &cur);
    if (rc) return rc;
    long long rb = (m + 255) / 256;
    if (rb > 16LL * sm_count()) rb = 16LL * sm_count();
#define DET(F, M) launch_k(det_reduce_kernel<F, M>, dim3((unsigned)rb), dim3(256), 0, st, g, gd, x, (int)ldx, dy, (int)ld_dy, m, row_bits, \
                           (const uint32_t*)keys[cur], (const uint32_t*)vals[cur])
    cudaError_t e = cudaSuccess;
    switch (g.n_feat * 4 + mode) {
        case 1 * 4 + 0: e = DET(1, IDRK_HASH_REFERENCE); break;
        case 2 * 4 + 0: e = DET(2, IDRK_HASH_REFERENCE); break;
        case 4 * 4 + 0: e = DET(4, IDRK_HASH_REFERENCE); break;
        case 8 * 4 + 0: e = DET(8, IDRK_HASH_REFERENCE); break;
        case 1 * 4 + 1: e = DET(1, IDRK_HASH_TRILINEAR); break;
        case 2 * 4 + 1: e = DET(2, IDRK_HASH_TRILINEAR); break;
        case 4 * 4 + 1: e = DET(4, IDRK_HASH_TRILINEAR); break;
        case 8 * 4 + 1: e = DET(8, IDRK_HASH_TRILINEAR); break;
        case 2 * 4 + 2: e = DET(2, IDRK_HASH_NGP); break;
        default: return IDRK_E_UNSUP;
    }
#undef DET
    if (e != cudaSuccess) return (int)e;
    IDRK_LAUNCH_CHECK();
    return 0;
}
