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
// Design (HBM/L2-bound gather):
//   * one thread owns one point and walks all levels, so L (or 8L) independent gathers are in flight
//     per thread and the [n, width] output row is produced once, in its final layout;
//   * rows are staged in shared memory (odd row stride -> conflict free) and written back with
//     fully coalesced 128-byte stores; the input is read once, the output written once;
//   * tables are read through the read-only path (ld.global.nc): coarse levels live in L1, the
//     rest in the 126 MB L2;  persistent grid = (resident CTAs per SM) x (SM count);
//   * backward: dL/dy tile staged the same way; table gradients go out as vector reductions
//     (red.global.add.v2.f32).  Levels whose whole table fits a shared-memory budget are first
//     accumulated per CTA in shared memory and flushed once (kills the contention on tiny tables
//     such as the reference configs' T = 32).
#include "common.cuh"

namespace idrk {

struct GridDev {
    int n_levels, n_feat, n_fourier, width;
    float res[IDRK_MAX_LEVELS];
    uint32_t rows[IDRK_MAX_LEVELS];
    uint32_t pow2mask[IDRK_MAX_LEVELS];     // rows-1 when rows is a power of two, else 0
    unsigned long long magic[IDRK_MAX_LEVELS];   // 2^64 / rows + 1 (Lemire fastmod), used when rows is not a power of two
    const float* tables[IDRK_MAX_LEVELS];
    const float* B;                         // [3, C]
};

struct GradDev {
    float* grad[IDRK_MAX_LEVELS];
    int small_off[IDRK_MAX_LEVELS];         // offset (floats) into the CTA's shared accumulator, -1 = global
    int small_total;                        // floats
};

__device__ __forceinline__ uint32_t hash3(uint32_t c0, uint32_t c1, uint32_t c2) {
    return c0 ^ (c1 * 3u) ^ (c2 * 2654435761u);
}
// h mod rows.  Power-of-two tables mask; others use the exact 64-bit fastmod  (M*h mod 2^64) * rows >> 64.
__device__ __forceinline__ uint32_t wrap(uint32_t h, uint32_t rows, uint32_t mask, unsigned long long magic) {
    return mask ? (h & mask) : (uint32_t)__umul64hi(magic * (unsigned long long)h, (unsigned long long)rows);
}
// .long() of an fp32 value: truncate toward zero to int64, keep the low 32 bits.  Inside the int32 range the
// 32-bit conversion has the same low bits (two's complement); only huge magnitudes need the 64-bit one.
__device__ __forceinline__ uint32_t trunc_u32(float v) {
    return fabsf(v) < 2147483520.f ? (uint32_t)__float2int_rz(v) : (uint32_t)(unsigned long long)__float2ll_rz(v);
}

// sin/cos of a moderate fp32 argument: two-constant Cody-Waite reduction by 2*pi, then the SFU approximations
// on [-pi, pi] (abs error < 1e-6, inside the 4e-6 parity tolerance); huge arguments take the accurate libm path.
__device__ __forceinline__ void sincos_fast(float x, float* sn, float* cs) {
    if (fabsf(x) > 8192.f) { sincosf(x, sn, cs); return; }
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, x);          // 2*pi rounded to fp32
    r = fmaf(-k, -1.7484555314695172e-07f, r);           // 2*pi - fp32(2*pi)
    *sn = __sinf(r);
    *cs = __cosf(r);
}

template <int F> struct Feat;
template <> struct Feat<1> { using T = float;  };
template <> struct Feat<2> { using T = float2; };
template <> struct Feat<4> { using T = float4; };
template <> struct Feat<8> { using T = float4; };   // two float4

template <int F>
__device__ __forceinline__ void gather(const float* table, uint32_t idx, float (&v)[F]) {
    if constexpr (F == 1) {
        v[0] = __ldg(table + idx);
    } else if constexpr (F == 2) {
        float2 t = __ldg(reinterpret_cast<const float2*>(table) + idx);
        v[0] = t.x; v[1] = t.y;
    } else if constexpr (F == 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(table) + idx);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        float4 a = __ldg(reinterpret_cast<const float4*>(table) + 2 * (size_t)idx);
        float4 b = __ldg(reinterpret_cast<const float4*>(table) + 2 * (size_t)idx + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
}

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
// forward
// ------------------------------------------------------------------------------------------
// Reference mode (one gather per level): whole-row block staging, one float4 write-back per row.
template <int F, int MODE, int ROWS>
__global__ void __launch_bounds__(ROWS)
hash_encode_fwd_block_kernel(const GridDev g, const float* __restrict__ x, long long n, int ldx,
                       float* __restrict__ out, int ld_out, uint32_t* __restrict__ idx_dbg,
                       const int* __restrict__ m_count) {
    extern __shared__ float smem[];
    if (m_count != nullptr) { const long long c = *m_count; n = c < n ? c : n; }
    const int lds = ld_out | 1;                 // odd stride: per-thread row writes hit distinct banks
    float* s_rows = smem;                       // [ROWS][lds]
    float* s_B = smem + ROWS * lds;             // [3][C]
    const int C = g.n_fourier, L = g.n_levels;
    for (int i = threadIdx.x; i < 3 * C; i += ROWS) s_B[i] = g.B[i];
    __syncthreads();

    const long long n_tiles = (n + ROWS - 1) / ROWS;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p = tile * ROWS + threadIdx.x;
        float* row = s_rows + threadIdx.x * lds;
        if (p < n) {
            const float x0 = x[p * ldx + 0], x1 = x[p * ldx + 1], x2 = x[p * ldx + 2];
            int col = 0;
            if (C > 0) {
                row[0] = x0; row[1] = x1; row[2] = x2;
                const float t0 = __fmul_rn(x0, 6.283185307179586f);
                const float t1 = __fmul_rn(x1, 6.283185307179586f);
                const float t2 = __fmul_rn(x2, 6.283185307179586f);
                for (int c = 0; c < C; ++c) {
                    float xp = __fmul_rn(t0, s_B[c]);
                    xp = __fmaf_rn(t1, s_B[C + c], xp);
                    xp = __fmaf_rn(t2, s_B[2 * C + c], xp);
                    float sn, cs;
                    sincos_fast(xp, &sn, &cs);
                    row[3 + c] = sn;
                    row[3 + C + c] = cs;
                }
                col = 3 + 2 * C;
            }
            if constexpr (MODE == IDRK_HASH_REFERENCE) {
#pragma unroll 4
                for (int l = 0; l < L; ++l) {
                    const float r = g.res[l];
                    const uint32_t h = hash3(trunc_u32(__fmul_rn(x0, r)), trunc_u32(__fmul_rn(x1, r)),
                                             trunc_u32(__fmul_rn(x2, r)));
                    float v[F];
                    gather<F>(g.tables[l], wrap(h, g.rows[l], g.pow2mask[l], g.magic[l]), v);
#pragma unroll
                    for (int f = 0; f < F; ++f) row[col + l * F + f] = v[f];
                }
            } else {
#pragma unroll 2
                for (int l = 0; l < L; ++l) {
                    const float r = g.res[l];
                    const float s0 = __fmul_rn(x0, r), s1 = __fmul_rn(x1, r), s2 = __fmul_rn(x2, r);
                    const float f0 = floorf(s0), f1 = floorf(s1), f2 = floorf(s2);
                    const float w0 = s0 - f0, w1 = s1 - f1, w2 = s2 - f2;
                    const uint32_t c0 = trunc_u32(f0), c1 = trunc_u32(f1), c2 = trunc_u32(f2);
                    const uint32_t rows = g.rows[l], mask = g.pow2mask[l];
                    const unsigned long long magic = g.magic[l];
                    const float* tab = g.tables[l];
                    float v[8][F];
                    // hash3 is an xor of three per-dimension terms: form the 2 x 3 terms once, xor per corner
                    const uint32_t hx0 = c0, hx1 = c0 + 1u, hy0 = c1 * 3u, hy1 = hy0 + 3u;
                    const uint32_t hz0 = c2 * 2654435761u, hz1 = hz0 + 2654435761u;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        gather<F>(tab, wrap(((k & 1) ? hx1 : hx0) ^ ((k & 2) ? hy1 : hy0) ^ ((k & 4) ? hz1 : hz0), rows, mask, magic), v[k]);
                    float acc[F];
#pragma unroll
                    for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float wk = ((k & 1) ? w0 : 1.f - w0) * ((k & 2) ? w1 : 1.f - w1) * ((k & 4) ? w2 : 1.f - w2);
#pragma unroll
                        for (int f = 0; f < F; ++f) acc[f] = fmaf(wk, v[k][f], acc[f]);
                    }
#pragma unroll
                    for (int f = 0; f < F; ++f) row[col + l * F + f] = acc[f];
                }
            }
            for (int c = g.width; c < ld_out; ++c) row[c] = 0.f;

            if (idx_dbg != nullptr) {           // debug / parity output: table row of all 8 corners
                for (int l = 0; l < L; ++l) {
                    const float r = g.res[l];
                    const float s0 = __fmul_rn(x0, r), s1 = __fmul_rn(x1, r), s2 = __fmul_rn(x2, r);
                    uint32_t c0, c1, c2;
                    if constexpr (MODE == IDRK_HASH_REFERENCE) { c0 = trunc_u32(s0); c1 = trunc_u32(s1); c2 = trunc_u32(s2); }
                    else { c0 = trunc_u32(floorf(s0)); c1 = trunc_u32(floorf(s1)); c2 = trunc_u32(floorf(s2)); }
                    for (int k = 0; k < 8; ++k)
                        idx_dbg[(p * L + l) * 8 + k] =
                            wrap(hash3(c0 + (k & 1), c1 + ((k >> 1) & 1), c2 + ((k >> 2) & 1)), g.rows[l], g.pow2mask[l], g.magic[l]);
                }
            }
        }
        __syncthreads();
        // coalesced write-back: the tile's rows are contiguous in global memory (ld_out floats each)
        const long long rows_here = min((long long)ROWS, n - tile * ROWS);
        float* gdst = out + tile * ROWS * (long long)ld_out;
        // the tile's rows are contiguous in global memory: lanes walk consecutive floats (conflict-free smem
        // reads with the odd row stride, 128-byte coalesced stores)
        if ((ld_out & 3) == 0) {
            const int ld4 = ld_out >> 2;
            for (int r = threadIdx.x >> 5; r < rows_here; r += ROWS / 32) {
                const float* src = s_rows + r * lds;
                float4* dst = reinterpret_cast<float4*>(gdst + (long long)r * ld_out);
                for (int c4 = threadIdx.x & 31; c4 < ld4; c4 += 32)
                    st_stream4(dst + c4, make_float4(src[4 * c4], src[4 * c4 + 1], src[4 * c4 + 2], src[4 * c4 + 3]));
            }
        } else {
            for (int r = threadIdx.x >> 5; r < rows_here; r += ROWS / 32) {
                const float* src = s_rows + r * lds;
                float* dst = gdst + (long long)r * ld_out;
                for (int c = threadIdx.x & 31; c < ld_out; c += 32) dst[c] = src[c];
            }
        }
        __syncthreads();
    }
}


// Writes columns [coloff, coloff + ncols) of `rows_here` consecutive output rows from a warp-private tile.
// The 16-byte aligned middle of each row segment goes out as float4 (8 lanes per row, 4 rows per instruction),
// the unaligned head / tail columns as scalars.
__device__ __forceinline__ void flush_tile(const float* tile, int tws, float* __restrict__ out, long long row0, int rows_here,
                                           int ld_out, int coloff, int ncols, int lane, bool vec_ok) {
    int head = 0, nv = 0;
    if (vec_ok) { head = (4 - (coloff & 3)) & 3; if (head > ncols) head = ncols; nv = (ncols - head) >> 2; }
    const int tail0 = head + 4 * nv;
    if (nv > 0) {
        int lanes_per_row = 1;
        while (lanes_per_row < nv) lanes_per_row <<= 1;              // 8 for a 32..35-column pass
        if (lanes_per_row > 32) lanes_per_row = 32;
        const int rows_per_it = 32 / lanes_per_row;
        const int lr = lane / lanes_per_row, lc = lane % lanes_per_row;
        for (int r = lr; r < rows_here; r += rows_per_it) {
            const float* src = tile + r * tws + head;
            float* dst = out + (row0 + r) * (long long)ld_out + coloff + head;
            for (int c4 = lc; c4 < nv; c4 += lanes_per_row)
                st_stream4(reinterpret_cast<float4*>(dst) + c4,
                           make_float4(src[4 * c4], src[4 * c4 + 1], src[4 * c4 + 2], src[4 * c4 + 3]));
        }
    }
    for (int c = 0; c < head; ++c)
        for (int r = lane; r < rows_here; r += 32) out[(row0 + r) * (long long)ld_out + coloff + c] = tile[r * tws + c];
    for (int c = tail0; c < ncols; ++c)
        for (int r = lane; r < rows_here; r += 32) out[(row0 + r) * (long long)ld_out + coloff + c] = tile[r * tws + c];
}

// Warp-private staging: each warp owns 32 points and a [32][TW+1] shared tile (TW = max(prefix width, level width)),
// used twice per tile of points - once for the Fourier prefix columns, once for the level columns - so the shared
// footprint is ~4.7 KB per warp (48+ resident warps per SM instead of 24 with whole-row staging), there is no
// block-wide barrier in the point loop, and global traffic stays coalesced (row segments of 128 bytes).
template <int F, int MODE, int ROWS>
__global__ void __launch_bounds__(ROWS)
hash_encode_fwd_kernel(const GridDev g, const float* __restrict__ x, long long n, int ldx,
                       float* __restrict__ out, int ld_out, uint32_t* __restrict__ idx_dbg,
                       const int* __restrict__ m_count, int tw) {
    extern __shared__ float smem[];
    if (m_count != nullptr) { const long long c = *m_count; n = c < n ? c : n; }
    const int C = g.n_fourier, L = g.n_levels;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tws = tw + 1;                              // odd-ish stride: row writes and column reads conflict free
    float* s_B = smem;                                   // [3][C]
    float* tile = smem + 3 * C + warp * (32 * tws);      // [32][tws], private to this warp
    for (int i = threadIdx.x; i < 3 * C; i += ROWS) s_B[i] = g.B[i];
    __syncthreads();
    const int pre = C > 0 ? 3 + 2 * C : 0;
    const int lev = ld_out - pre;                        // level columns + zero padding up to ld_out
    float* my = tile + lane * tws;
    const bool vec_ok = (ld_out & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;

    const long long n_wtiles = (n + 31) / 32;
    const long long wstride = (long long)gridDim.x * (ROWS / 32);
    for (long long wt = (long long)blockIdx.x * (ROWS / 32) + warp; wt < n_wtiles; wt += wstride) {
        const long long p0 = wt * 32, p = p0 + lane;
        const bool valid = p < n;
        const int rows_here = (int)min(32LL, n - p0);
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;
        if (valid) { x0 = x[p * ldx + 0]; x1 = x[p * ldx + 1]; x2 = x[p * ldx + 2]; }
        if (C > 0) {
            my[0] = x0; my[1] = x1; my[2] = x2;
            const float t0 = __fmul_rn(x0, 6.283185307179586f);
            const float t1 = __fmul_rn(x1, 6.283185307179586f);
            const float t2 = __fmul_rn(x2, 6.283185307179586f);
            for (int c = 0; c < C; ++c) {
                float xp = __fmul_rn(t0, s_B[c]);
                xp = __fmaf_rn(t1, s_B[C + c], xp);
                xp = __fmaf_rn(t2, s_B[2 * C + c], xp);
                float sn, cs;
                sincos_fast(xp, &sn, &cs);
                my[3 + c] = sn;
                my[3 + C + c] = cs;
            }
            __syncwarp();
            flush_tile(tile, tws, out, p0, rows_here, ld_out, 0, pre, lane, vec_ok);
            __syncwarp();
        }
        if constexpr (MODE == IDRK_HASH_REFERENCE) {
#pragma unroll 4
            for (int l = 0; l < L; ++l) {
                const float r = g.res[l];
                const uint32_t h = hash3(trunc_u32(__fmul_rn(x0, r)), trunc_u32(__fmul_rn(x1, r)),
                                         trunc_u32(__fmul_rn(x2, r)));
                float v[F];
                gather<F>(g.tables[l], wrap(h, g.rows[l], g.pow2mask[l], g.magic[l]), v);
#pragma unroll
                for (int f = 0; f < F; ++f) my[l * F + f] = v[f];
            }
        } else {
#pragma unroll 2
            for (int l = 0; l < L; ++l) {
                const float r = g.res[l];
                const float s0 = __fmul_rn(x0, r), s1 = __fmul_rn(x1, r), s2 = __fmul_rn(x2, r);
                const float f0 = floorf(s0), f1 = floorf(s1), f2 = floorf(s2);
                const float w0 = s0 - f0, w1 = s1 - f1, w2 = s2 - f2;
                const uint32_t c0 = trunc_u32(f0), c1 = trunc_u32(f1), c2 = trunc_u32(f2);
                const uint32_t rows = g.rows[l], mask = g.pow2mask[l];
                const unsigned long long magic = g.magic[l];
                const float* tab = g.tables[l];
                float v[8][F];
                // hash3 is an xor of three per-dimension terms: form the 2 x 3 terms once, xor per corner
                const uint32_t hx0 = c0, hx1 = c0 + 1u, hy0 = c1 * 3u, hy1 = hy0 + 3u;
                const uint32_t hz0 = c2 * 2654435761u, hz1 = hz0 + 2654435761u;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    gather<F>(tab, wrap(((k & 1) ? hx1 : hx0) ^ ((k & 2) ? hy1 : hy0) ^ ((k & 4) ? hz1 : hz0), rows, mask, magic), v[k]);
                float acc[F];
#pragma unroll
                for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float wk = ((k & 1) ? w0 : 1.f - w0) * ((k & 2) ? w1 : 1.f - w1) * ((k & 4) ? w2 : 1.f - w2);
#pragma unroll
                    for (int f = 0; f < F; ++f) acc[f] = fmaf(wk, v[k][f], acc[f]);
                }
#pragma unroll
                for (int f = 0; f < F; ++f) my[l * F + f] = acc[f];
            }
        }
        for (int c = L * F; c < lev; ++c) my[c] = 0.f;
        __syncwarp();
        flush_tile(tile, tws, out, p0, rows_here, ld_out, pre, lev, lane, vec_ok);
        __syncwarp();

        if (idx_dbg != nullptr && valid) {      // debug / parity output: table row of all 8 corners
            for (int l = 0; l < L; ++l) {
                const float r = g.res[l];
                const float s0 = __fmul_rn(x0, r), s1 = __fmul_rn(x1, r), s2 = __fmul_rn(x2, r);
                uint32_t c0, c1, c2;
                if constexpr (MODE == IDRK_HASH_REFERENCE) { c0 = trunc_u32(s0); c1 = trunc_u32(s1); c2 = trunc_u32(s2); }
                else { c0 = trunc_u32(floorf(s0)); c1 = trunc_u32(floorf(s1)); c2 = trunc_u32(floorf(s2)); }
                for (int k = 0; k < 8; ++k)
                    idx_dbg[(p * L + l) * 8 + k] =
                        wrap(hash3(c0 + (k & 1), c1 + ((k >> 1) & 1), c2 + ((k >> 2) & 1)), g.rows[l], g.pow2mask[l], g.magic[l]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
template <int F, int MODE, int ROWS>
__global__ void __launch_bounds__(ROWS)
hash_encode_bwd_kernel(const GridDev g, const GradDev gd, const float* __restrict__ x, long long n, int ldx,
                       const float* __restrict__ dy, int ld_dy, float* __restrict__ dx) {
    extern __shared__ float smem[];
    const int lds = ld_dy | 1;
    float* s_rows = smem;                        // [ROWS][lds]
    float* s_B = smem + ROWS * lds;              // [3][C]
    float* s_acc = s_B + 3 * g.n_fourier;        // [small_total]
    const int C = g.n_fourier, L = g.n_levels;
    for (int i = threadIdx.x; i < 3 * C; i += ROWS) s_B[i] = g.B[i];
    for (int i = threadIdx.x; i < gd.small_total; i += ROWS) s_acc[i] = 0.f;
    __syncthreads();

    const int col0 = (C > 0) ? 3 + 2 * C : 0;
    const long long n_tiles = (n + ROWS - 1) / ROWS;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long rows_here = min((long long)ROWS, n - tile * ROWS);
        const float* gsrc = dy + tile * ROWS * (long long)ld_dy;
        if ((ld_dy & 3) == 0 && ((reinterpret_cast<uintptr_t>(dy) & 15u) == 0)) {
            const int ld4 = ld_dy >> 2;
            for (int r = threadIdx.x >> 5; r < rows_here; r += ROWS / 32) {
                float* dst = s_rows + r * lds;
                for (int c4 = threadIdx.x & 31; c4 < ld4; c4 += 32) {
                    float4 v = ld_stream4(reinterpret_cast<const float4*>(gsrc + (long long)r * ld_dy) + c4);
                    dst[4 * c4] = v.x; dst[4 * c4 + 1] = v.y; dst[4 * c4 + 2] = v.z; dst[4 * c4 + 3] = v.w;
                }
            }
        } else {
            const int total = (int)rows_here * ld_dy;
            for (int i = threadIdx.x; i < total; i += ROWS) {
                const int r = i / ld_dy, c = i - r * ld_dy;
                s_rows[r * lds + c] = gsrc[i];
            }
        }
        __syncthreads();
        const long long p = tile * ROWS + threadIdx.x;
        if (p < n) {
            const float* row = s_rows + threadIdx.x * lds;
            const float x0 = x[p * ldx + 0], x1 = x[p * ldx + 1], x2 = x[p * ldx + 2];
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
            if (C > 0 && dx != nullptr) {
                g0 = row[0]; g1 = row[1]; g2 = row[2];
                const float t0 = __fmul_rn(x0, 6.283185307179586f);
                const float t1 = __fmul_rn(x1, 6.283185307179586f);
                const float t2 = __fmul_rn(x2, 6.283185307179586f);
                for (int c = 0; c < C; ++c) {
                    float xp = __fmul_rn(t0, s_B[c]);
                    xp = __fmaf_rn(t1, s_B[C + c], xp);
                    xp = __fmaf_rn(t2, s_B[2 * C + c], xp);
                    float sn, cs;
                    sincos_fast(xp, &sn, &cs);
                    const float dxp = (row[3 + c] * cs - row[3 + C + c] * sn) * 6.283185307179586f;
                    g0 = fmaf(dxp, s_B[c], g0);
                    g1 = fmaf(dxp, s_B[C + c], g1);
                    g2 = fmaf(dxp, s_B[2 * C + c], g2);
                }
            }
#pragma unroll 2
            for (int l = 0; l < L; ++l) {
                const float r = g.res[l];
                const uint32_t rows = g.rows[l], mask = g.pow2mask[l];
                const unsigned long long magic = g.magic[l];
                float gy[F];
#pragma unroll
                for (int f = 0; f < F; ++f) gy[f] = row[col0 + l * F + f];
                const int soff = gd.small_off[l];
                const bool do_scatter = gd.grad[l] != nullptr;
                if constexpr (MODE == IDRK_HASH_REFERENCE) {
                    if (!do_scatter) continue;
                    const uint32_t idx = wrap(hash3(trunc_u32(__fmul_rn(x0, r)), trunc_u32(__fmul_rn(x1, r)),
                                                    trunc_u32(__fmul_rn(x2, r))), rows, mask, magic);
                    if (soff >= 0) {
#pragma unroll
                        for (int f = 0; f < F; ++f) atomicAdd(s_acc + soff + idx * F + f, gy[f]);
                    } else {
                        scatter_add<F>(gd.grad[l], idx, gy);
                    }
                } else {
                    const float s0 = __fmul_rn(x0, r), s1 = __fmul_rn(x1, r), s2 = __fmul_rn(x2, r);
                    const float f0 = floorf(s0), f1 = floorf(s1), f2 = floorf(s2);
                    const float w0 = s0 - f0, w1 = s1 - f1, w2 = s2 - f2;
                    const uint32_t c0 = trunc_u32(f0), c1 = trunc_u32(f1), c2 = trunc_u32(f2);
                    const float* tab = g.tables[l];
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t idx = wrap(hash3(c0 + (k & 1), c1 + ((k >> 1) & 1), c2 + ((k >> 2) & 1)), rows, mask, magic);
                        const float a0 = (k & 1) ? w0 : 1.f - w0, a1 = (k & 2) ? w1 : 1.f - w1, a2 = (k & 4) ? w2 : 1.f - w2;
                        const float wk = a0 * a1 * a2;
                        float v[F];
#pragma unroll
                        for (int f = 0; f < F; ++f) v[f] = wk * gy[f];
                        if (!do_scatter) {
                        } else if (soff >= 0) {
#pragma unroll
                            for (int f = 0; f < F; ++f) atomicAdd(s_acc + soff + idx * F + f, v[f]);
                        } else {
                            scatter_add<F>(gd.grad[l], idx, v);
                        }
                        if (dx != nullptr) {
                            float t[F];
                            gather<F>(tab, idx, t);
                            float dot = 0.f;
#pragma unroll
                            for (int f = 0; f < F; ++f) dot = fmaf(t[f], gy[f], dot);
                            d0 = fmaf(((k & 1) ? 1.f : -1.f) * a1 * a2, dot, d0);
                            d1 = fmaf(((k & 2) ? 1.f : -1.f) * a0 * a2, dot, d1);
                            d2 = fmaf(((k & 4) ? 1.f : -1.f) * a0 * a1, dot, d2);
                        }
                    }
                    g0 = fmaf(d0, r, g0); g1 = fmaf(d1, r, g1); g2 = fmaf(d2, r, g2);
                }
            }
            if (dx != nullptr) { dx[p * 3 + 0] = g0; dx[p * 3 + 1] = g1; dx[p * 3 + 2] = g2; }
        }
        __syncthreads();
    }
    // flush the CTA-local accumulators of the small tables
    if (gd.small_total > 0) {
        for (int l = 0; l < L; ++l) {
            const int soff = gd.small_off[l];
            if (soff < 0) continue;
            const int cnt = (int)g.rows[l] * F;
            for (int i = threadIdx.x; i < cnt; i += ROWS) {
                const float v = s_acc[soff + i];
                if (v != 0.f) atomicAdd(gd.grad[l] + i, v);
            }
        }
    }
}

static int fill_grid(const idrk_hashgrid_t* h, GridDev& g) {
    if (h == nullptr) return IDRK_E_ARG;
    if (h->n_levels < 0 || h->n_levels > IDRK_MAX_LEVELS) return IDRK_E_ARG;
    if (h->n_levels == 0 && h->n_fourier == 0) return IDRK_E_ARG;
    if (h->n_feat != 1 && h->n_feat != 2 && h->n_feat != 4 && h->n_feat != 8) return IDRK_E_UNSUP;
    if (h->n_fourier < 0 || h->n_fourier > 64) return IDRK_E_ARG;
    if (h->n_fourier > 0 && h->fourier_B == nullptr) return IDRK_E_ARG;
    if (h->frac_mode != IDRK_HASH_REFERENCE && h->frac_mode != IDRK_HASH_TRILINEAR) return IDRK_E_ARG;
    g.n_levels = h->n_levels; g.n_feat = h->n_feat; g.n_fourier = h->n_fourier;
    g.width = (h->n_fourier > 0 ? 3 + 2 * h->n_fourier : 0) + h->n_levels * h->n_feat;
    g.B = h->fourier_B;
    const size_t align = (h->n_feat >= 4) ? 16 : 4 * (size_t)h->n_feat;
    for (int l = 0; l < h->n_levels; ++l) {
        if (h->tables[l] == nullptr || h->rows[l] == 0) return IDRK_E_ARG;
        if (reinterpret_cast<uintptr_t>(h->tables[l]) % align) return IDRK_E_ALIGN;
        g.res[l] = h->res[l]; g.rows[l] = h->rows[l]; g.tables[l] = h->tables[l];
        g.pow2mask[l] = ((h->rows[l] & (h->rows[l] - 1)) == 0) ? h->rows[l] - 1 : 0;
        g.magic[l] = ~0ull / h->rows[l] + 1ull;
        if (h->rows[l] == 1) { g.pow2mask[l] = 0; g.magic[l] = 0; }
    }
    for (int l = h->n_levels; l < IDRK_MAX_LEVELS; ++l) { g.res[l] = 0; g.rows[l] = 1; g.tables[l] = nullptr; g.pow2mask[l] = 0; g.magic[l] = 0; }
    return 0;
}

template <typename K>
static int persistent_grid(K kernel, int threads, size_t smem, long long n_tiles) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long g = (long long)per_sm * sm_count();
    if (g > n_tiles) g = n_tiles;
    if (g < 1) g = 1;
    return (int)g;
}

template <int F, int MODE, int ROWS>
static int launch_fwd(const GridDev& g, const float* x, long long n, int ldx, float* out, int ld_out,
                      uint32_t* idx_dbg, const int* m_count, cudaStream_t st) {
    if (MODE == IDRK_HASH_REFERENCE && (size_t)ROWS * (ld_out | 1) * sizeof(float) <= 72 * 1024) {
        // one gather per level: issue-bound, whole-row staging with a single float4 write-back wins
        const size_t smem_b = ((size_t)ROWS * (ld_out | 1) + 3 * g.n_fourier) * sizeof(float);
        auto kb = hash_encode_fwd_block_kernel<F, MODE, ROWS>;
        IDRK_CUDA_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
        const int grid_b = persistent_grid(kb, ROWS, smem_b, (n + ROWS - 1) / ROWS);
        kb<<<grid_b, ROWS, smem_b, st>>>(g, x, n, ldx, out, ld_out, idx_dbg, m_count);
        IDRK_LAUNCH_CHECK();
        return 0;
    }
    // 8 gathers per level: latency-bound, warp-private staging doubles the resident warps
    const int pre = g.n_fourier > 0 ? 3 + 2 * g.n_fourier : 0;
    const int tw = (pre > ld_out - pre ? pre : ld_out - pre) | 1;
    const size_t smem = ((size_t)(ROWS / 32) * 32 * (tw + 1) + 3 * g.n_fourier) * sizeof(float);
    auto kern = hash_encode_fwd_kernel<F, MODE, ROWS>;
    IDRK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = persistent_grid(kern, ROWS, smem, (n + ROWS - 1) / ROWS);
    kern<<<grid, ROWS, smem, st>>>(g, x, n, ldx, out, ld_out, idx_dbg, m_count, tw);
    IDRK_LAUNCH_CHECK();
    return 0;
}

template <int F, int MODE, int ROWS>
static int launch_bwd(const GridDev& g, const GradDev& gd, const float* x, long long n, int ldx, const float* dy,
                      int ld_dy, float* dx, cudaStream_t st) {
    const size_t smem = ((size_t)ROWS * (ld_dy | 1) + 3 * g.n_fourier + gd.small_total) * sizeof(float);
    auto kern = hash_encode_bwd_kernel<F, MODE, ROWS>;
    IDRK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = persistent_grid(kern, ROWS, smem, (n + ROWS - 1) / ROWS);
    kern<<<grid, ROWS, smem, st>>>(g, gd, x, n, ldx, dy, ld_dy, dx);
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
        default: return IDRK_E_UNSUP;                                                          \
    }

}  // namespace idrk

using namespace idrk;

extern "C" int idrk_version(void) { return 1; }

extern "C" int idrk_device_sm_count(int* out_sms) {
    if (!out_sms) return IDRK_E_ARG;
    int dev = 0;
    IDRK_CUDA_TRY(cudaGetDevice(&dev));
    IDRK_CUDA_TRY(cudaDeviceGetAttribute(out_sms, cudaDevAttrMultiProcessorCount, dev));
    return 0;
}

extern "C" int idrk_hash_encode_fwd(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx,
                                    float* out, int32_t ld_out, uint32_t* idx_debug, const int32_t* m_count,
                                    void* stream) {
    GridDev g;
    int rc = fill_grid(h_grid, g);
    if (rc) return rc;
    if (n < 0 || ldx < 3 || x == nullptr || out == nullptr || ld_out < g.width) return IDRK_E_ARG;
    if (n == 0) return 0;
    if ((ld_out & 3) == 0 && !aligned16(out)) return IDRK_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int mode = h_grid->frac_mode;
#define CALL(F, M) launch_fwd<F, M, 128>(g, x, n, ldx, out, ld_out, idx_debug, m_count, st)
    IDRK_DISPATCH_F_MODE(CALL)
#undef CALL
}

extern "C" int idrk_hash_encode_bwd(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx,
                                    const float* dy, int32_t ld_dy, float* const* h_grad_tables, float* dx,
                                    void* stream) {
    GridDev g;
    int rc = fill_grid(h_grid, g);
    if (rc) return rc;
    if (n < 0 || ldx < 3 || x == nullptr || dy == nullptr || ld_dy < g.width) return IDRK_E_ARG;
    if (h_grad_tables == nullptr && dx == nullptr) return IDRK_E_ARG;
    if (n == 0) return 0;
    GradDev gd;
    gd.small_total = 0;
    const int budget = 8192;                    // floats of CTA-local accumulators (32 KB)
    const size_t align = (g.n_feat >= 4) ? 16 : 4 * (size_t)g.n_feat;
    for (int l = 0; l < IDRK_MAX_LEVELS; ++l) { gd.grad[l] = nullptr; gd.small_off[l] = -1; }
    for (int l = 0; l < g.n_levels && h_grad_tables != nullptr; ++l) {
        if (h_grad_tables[l] == nullptr) return IDRK_E_ARG;
        if (reinterpret_cast<uintptr_t>(h_grad_tables[l]) % align) return IDRK_E_ALIGN;
        gd.grad[l] = h_grad_tables[l];
        const long long cnt = (long long)g.rows[l] * g.n_feat;
        if (cnt <= 2048 && gd.small_total + cnt <= budget) { gd.small_off[l] = gd.small_total; gd.small_total += (int)cnt; }
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int mode = h_grid->frac_mode;
    const size_t row_bytes = (size_t)(ld_dy | 1) * sizeof(float);
    if (row_bytes * 128 <= 72 * 1024) {
#define CALL(F, M) launch_bwd<F, M, 128>(g, gd, x, n, ldx, dy, ld_dy, dx, st)
        IDRK_DISPATCH_F_MODE(CALL)
#undef CALL
    } else if (row_bytes * 64 <= 160 * 1024) {
#define CALL(F, M) launch_bwd<F, M, 64>(g, gd, x, n, ldx, dy, ld_dy, dx, st)
        IDRK_DISPATCH_F_MODE(CALL)
#undef CALL
    }
    return IDRK_E_UNSUP;
}
