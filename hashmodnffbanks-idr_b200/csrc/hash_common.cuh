// Device helpers shared by the hash-grid kernels (hash_encode.cu) and the fused filter-bank encoder (nffb.cu).
#pragma once
#include <math.h>
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
    uint32_t ngp_res[IDRK_MAX_LEVELS];      // IDRK_HASH_NGP: grid resolution R_l = ceil(scale_l) + 1
    uint32_t ngp_dense[IDRK_MAX_LEVELS];    // IDRK_HASH_NGP: 1 = dense indexing (R^3 <= rows), 0 = hashed
    int pair_x;                             // 8-corner mode: x-neighbour corners share one 16-byte access when they can
    int agg_runs;                           // 8-corner backward: in-lane aggregation of runs of points in the same cell
    // level-window passes (tables larger than L2 are walked a few levels at a time, see level_groups()): a launch may
    // cover levels [l0, l0 + n_levels) of the full grid only
    int pre_cols;                           // columns of the row in front of this launch's first level (prefix + earlier levels)
    int tail;                               // 1 = the launch contains the LAST level of the row (owns the pad columns)
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


inline int fill_grid(const idrk_hashgrid_t* h, GridDev& g) {
    if (h == nullptr) return IDRK_E_ARG;
    if (h->n_levels < 0 || h->n_levels > IDRK_MAX_LEVELS) return IDRK_E_ARG;
    if (h->n_levels == 0 && h->n_fourier == 0) return IDRK_E_ARG;
    if (h->n_feat != 1 && h->n_feat != 2 && h->n_feat != 4 && h->n_feat != 8) return IDRK_E_UNSUP;
    if (h->n_fourier < 0 || h->n_fourier > 64) return IDRK_E_ARG;
    if (h->n_fourier > 0 && h->fourier_B == nullptr) return IDRK_E_ARG;
    if (h->frac_mode != IDRK_HASH_REFERENCE && h->frac_mode != IDRK_HASH_TRILINEAR && h->frac_mode != IDRK_HASH_NGP) return IDRK_E_ARG;
    if (h->frac_mode == IDRK_HASH_NGP && h->n_feat != 2) return IDRK_E_UNSUP;
    g.n_levels = h->n_levels; g.n_feat = h->n_feat; g.n_fourier = h->n_fourier;
    g.width = (h->n_fourier > 0 ? 3 + 2 * h->n_fourier : 0) + h->n_levels * h->n_feat;
    g.pre_cols = h->n_fourier > 0 ? 3 + 2 * h->n_fourier : 0;
    g.tail = 1;
    g.B = h->fourier_B;
    { static const int pair = [] { const char* e = getenv("IDRK_HASH_PAIR_X"); return (e && e[0] == '0') ? 0 : 1; }(); g.pair_x = pair; }
    { static const int agg = [] { const char* e = getenv("IDRK_HASH_AGG"); return (e && e[0] == '0') ? 0 : 1; }(); g.agg_runs = agg; }
    const size_t align = (h->n_feat >= 4) ? 16 : 4 * (size_t)h->n_feat;
    for (int l = 0; l < h->n_levels; ++l) {
        if (h->tables[l] == nullptr || h->rows[l] == 0) return IDRK_E_ARG;
        if (reinterpret_cast<uintptr_t>(h->tables[l]) % align) return IDRK_E_ALIGN;
        g.res[l] = h->res[l]; g.rows[l] = h->rows[l]; g.tables[l] = h->tables[l];
        g.pow2mask[l] = ((h->rows[l] & (h->rows[l] - 1)) == 0) ? h->rows[l] - 1 : 0;
        g.magic[l] = ~0ull / h->rows[l] + 1ull;
        if (h->rows[l] == 1) { g.pow2mask[l] = 0; g.magic[l] = 0; }
        g.ngp_res[l] = 0; g.ngp_dense[l] = 0;
        if (h->frac_mode == IDRK_HASH_NGP) {
            if (!(h->res[l] > 0.f) || h->res[l] > 1.0e6f) return IDRK_E_ARG;
            const unsigned long long R = (unsigned long long)ceilf(h->res[l]) + 1ull;
            g.ngp_res[l] = (uint32_t)R;
            g.ngp_dense[l] = (R * R * R <= (unsigned long long)h->rows[l]) ? 1u : 0u;
        }
    }
    for (int l = h->n_levels; l < IDRK_MAX_LEVELS; ++l) { g.res[l] = 0; g.rows[l] = 1; g.tables[l] = nullptr; g.pow2mask[l] = 0; g.magic[l] = 0; g.ngp_res[l] = 0; g.ngp_dense[l] = 0; }
    return 0;
}

// Level windows.  Tables that do not fit the 126 MB L2 turn every gather / reduction into a DRAM sector access (measured,
// scripts/probes/atomics_probe.cu: 288 G random 8-byte gathers/s and 193 G reductions/s while the table is L2-resident,
// 73 G/s and 33 G/s at 256 MB).  The encode is then run as several launches over consecutive level windows whose tables
// fit a budget, so each window's tables are fetched from DRAM once and served from L2 for all points; x is re-read per
// window (12 B) and every window writes / reads only its own columns of the row.  Windows hold a power-of-two number of
// levels (the element-per-lane mapping wants L | 32).  Returns the number of windows; win[i] = first level of window i,
// win[n] = L.  One window = the classic single launch.
// Measured on 2^24 points, L = 16, F = 2 (gpurun_out/hash_mb_r02d_*): at T = 2^22 (314 MB) windows lift the 8-corner table-
// gradient pass 0.87 -> 1.09 Gpts/s and the reference-mode one 3.8 -> 4.1, but do nothing for the forward (2.20 -> 2.30;
// reference mode 5.1 -> 2.9: its single gather per level cannot pay for the extra x reads and partial-sector row writes);
// at T = 2^24 (1.2 GB, a level alone exceeds L2) the 8-corner forward gains 1.02 -> 1.74.  Hence `min_total`: the caller
// passes the table size from which windows pay for its pass.
inline int level_groups(const GridDev& g, long long n_points, long long min_total, int* win) {
    static const long long budget = [] {
        const char* e = getenv("IDRK_HASH_GROUP_MB");
        const long long mb = e ? atoll(e) : 80;
        return (mb <= 0 ? (1LL << 60) : mb << 20);
    }();
    long long total = 0;
    for (int l = 0; l < g.n_levels; ++l) total += (long long)g.rows[l] * g.n_feat * 4;
    int n = 0;
    win[0] = 0;
    // a small batch touches a fraction of the tables anyway: extra launches and x re-reads would only cost
    if (total <= budget || total <= min_total || g.n_levels <= 1 || n_points < (1LL << 20)) { win[1] = g.n_levels; return 1; }
    int l = 0;
    while (l < g.n_levels) {
        long long bytes = 0;
        int cnt = 0;
        while (l + cnt < g.n_levels && (cnt == 0 || bytes + (long long)g.rows[l + cnt] * g.n_feat * 4 <= budget)) {
            bytes += (long long)g.rows[l + cnt] * g.n_feat * 4;
            ++cnt;
        }
        int p2 = 1;
        while (p2 * 2 <= cnt) p2 *= 2;                    // power-of-two window
        l += p2;
        win[++n] = l;
    }
    return n;
}

// GridDev of the window [l0, l1) of `g` (prefix columns only with the first window)
inline GridDev level_window(const GridDev& g, int l0, int l1) {
    GridDev w = g;
    w.n_levels = l1 - l0;
    for (int l = 0; l < IDRK_MAX_LEVELS; ++l) {
        const int s = l0 + l;
        const bool in = s < l1;
        w.res[l] = in ? g.res[s] : 0.f; w.rows[l] = in ? g.rows[s] : 1; w.pow2mask[l] = in ? g.pow2mask[s] : 0;
        w.magic[l] = in ? g.magic[s] : 0; w.tables[l] = in ? g.tables[s] : nullptr;
        w.ngp_res[l] = in ? g.ngp_res[s] : 0; w.ngp_dense[l] = in ? g.ngp_dense[s] : 0;
    }
    if (l0 > 0) w.n_fourier = 0;
    w.pre_cols = g.pre_cols + l0 * g.n_feat;
    w.tail = (l1 == g.n_levels) ? 1 : 0;
    return w;
}

}  // namespace idrk
