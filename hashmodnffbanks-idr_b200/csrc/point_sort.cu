// Z-order (Morton) permutation of a point batch - the pre-pass that makes unordered batches spatially ordered for the
// hash-grid kernels (csrc/hash_encode.cu: run aggregation in the table-gradient pass, table-row locality in both passes;
// measured on 2^24 uniform points, L = 16, F = 2: forward 3.4 -> 4.5 Gpts/s at T = 2^19 and 2.2 -> 4.4 at T = 2^22,
// table-gradient pass 1.3 -> 2.2 and 0.87 -> 1.6).  The reference has no counterpart: its points arrive in ray order
// (model/ray_tracing.py) and its encoder is order-oblivious (model/embeddings/hashGridEmbedding.py:81-102).
//
//   key  = bit-interleaved lattice coordinates of (x - lo) / (hi - lo) on a (2^bits)^3 lattice, bits <= 10
//   sort = least-significant-digit radix sort of (key, index) pairs, 8 bits per pass, ceil(3 bits / 8) passes:
//          per pass   hist     one 4096-key tile per CTA, digit histogram in shared memory -> counts[digit][tile]
//                     scan     exclusive prefix sum of counts in digit-major order (tile scan + one CTA over tile sums)
//                     scatter  the tile again: stable in-tile rank by warp match (no atomics), keys and indices written
//                              to their final place of the pass
//   Every pass moves 20 B per point; nothing is allocated here - the caller passes the workspace.
#include "sort_common.cuh"

namespace idrk {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;     // 4096 keys per CTA
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int SCAN_TILE = 4096;                    // ints per CTA of the prefix-sum kernels (1024 threads x 4)

__device__ __forceinline__ uint32_t spread3(uint32_t v) {        // abcdefghij -> a00b00c00d00e00f00g00h00i00j
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// order-preserving float <-> uint32 map (for atomicMin / atomicMax on floats of either sign)
__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t b = __float_as_uint(f); return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu)); }

__global__ void bbox_init_kernel(uint32_t* __restrict__ box) {
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x < 3) box[threadIdx.x] = 0xFFFFFFFFu; else if (threadIdx.x < 6) box[threadIdx.x] = 0u;
}

// bounding box of the batch, kept on the device (no host read): box[0..2] = min, box[3..5] = max in f2ord encoding
__global__ void __launch_bounds__(256)
bbox_kernel(const float* __restrict__ x, long long n, int ldx, uint32_t* __restrict__ box) {
    pdl_wait();
    pdl_trigger();
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { const float v = x[i * ldx + d]; lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (lo[d] <= hi[d]) { atomicMin(box + d, f2ord(lo[d])); atomicMax(box + 3 + d, f2ord(hi[d])); }
        }
    }
}

__global__ void morton_key_kernel(const float* __restrict__ x, long long n, int ldx, float lo0, float lo1, float lo2,
                                  float s0, float s1, float s2, float qmax, const uint32_t* __restrict__ box,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    pdl_wait();
    pdl_trigger();
    if (box != nullptr) {                                    // box measured on the device by bbox_kernel
        lo0 = ord2f(box[0]); lo1 = ord2f(box[1]); lo2 = ord2f(box[2]);
        const float e0 = ord2f(box[3]) - lo0, e1 = ord2f(box[4]) - lo1, e2 = ord2f(box[5]) - lo2;
        s0 = e0 > 0.f ? (qmax + 1.f) / e0 : 0.f; s1 = e1 > 0.f ? (qmax + 1.f) / e1 : 0.f; s2 = e2 > 0.f ? (qmax + 1.f) / e2 : 0.f;
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = fminf(fmaxf((x[i * ldx] - lo0) * s0, 0.f), qmax);
        const float b = fminf(fmaxf((x[i * ldx + 1] - lo1) * s1, 0.f), qmax);
        const float c = fminf(fmaxf((x[i * ldx + 2] - lo2) * s2, 0.f), qmax);
        keys[i] = spread3((uint32_t)a) | (spread3((uint32_t)b) << 1) | (spread3((uint32_t)c) << 2);
        vals[i] = (uint32_t)i;
    }
}

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint32_t* __restrict__ keys, long long n, int shift, int n_tiles, uint32_t* __restrict__ counts) {
    pdl_wait();
    pdl_trigger();
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const long long t0 = (long long)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const long long i = t0 + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&sh[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * n_tiles + blockIdx.x] = sh[threadIdx.x];
}

// exclusive scan of one SCAN_TILE of `data` in place; the tile's total goes to sums[blockIdx.x]
__global__ void __launch_bounds__(1024)
scan_tiles_kernel(uint32_t* __restrict__ data, long long m, uint32_t* __restrict__ sums) {
    pdl_wait();
    pdl_trigger();
    __shared__ uint32_t s_w[32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (base + k < m) ? data[base + k] : 0u;
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t inc = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_w[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        s_w[lane] = winc - w;                                 // exclusive warp offsets
        if (lane == 31) sums[blockIdx.x] = winc;
    }
    __syncthreads();
    uint32_t run = s_w[warp] + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (base + k < m) data[base + k] = run; run += v[k]; }
}

// exclusive scan of up to 4096 tile sums by one CTA, in place
__global__ void __launch_bounds__(1024)
scan_sums_kernel(uint32_t* __restrict__ sums, int m) {
    pdl_wait();
    pdl_trigger();
    __shared__ uint32_t s_w[32];
    const int base = threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (base + k < m) ? sums[base + k] : 0u;
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t inc = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_w[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        s_w[lane] = winc - w;
    }
    __syncthreads();
    uint32_t run = s_w[warp] + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (base + k < m) sums[base + k] = run; run += v[k]; }
}

// Stable scatter of one tile.  In-tile order = (warp, round, lane) = input order, so equal digits keep their order.
// The tile is first sorted by digit INSIDE shared memory (local position = exclusive digit prefix of the tile + the
// item's stable rank within its digit) and then written out in that order: neighbouring threads write neighbouring
// elements of a digit run, so a run (16 items on average) leaves as whole sectors.  Writing each item straight from the
// lane that ranked it cost one 32-byte sector per 4-byte store (ncu: 7.9 M write sectors per 4 M pairs, 7.6 x the payload)
// and made this kernel 90 % of the sort.
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, long long n, int shift, int n_tiles,
                  const uint32_t* __restrict__ counts_scanned, const uint32_t* __restrict__ tile_offsets,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    pdl_wait();
    pdl_trigger();
    __shared__ uint32_t cnt[RS_WARPS][257];                  // digit 256 = out-of-range lanes of the last tile
    __shared__ uint32_t s_key[RS_TILE], s_val[RS_TILE];
    __shared__ uint32_t lbase[256];                          // exclusive digit prefix inside the tile
    __shared__ uint32_t gbase[256];                          // global position of the tile's first item of each digit
    __shared__ uint32_t s_wsum[RS_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 257; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const long long t0 = (long long)blockIdx.x * RS_TILE;
    const long long w0 = t0 + warp * (RS_ITEMS * 32);
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rk[RS_ITEMS];      // rk = (in-warp rank << 9) | digit
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const long long i = w0 + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys[i] : 0u;
        val[r] = ok ? vals[i] : 0u;
        const uint32_t d = ok ? ((key[r] >> shift) & 255u) : 256u;
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(m) - 1;
        uint32_t old = 0;
        if (lane == leader) { old = cnt[warp][d]; cnt[warp][d] = old + __popc(m); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rk[r] = ((old + __popc(m & ((1u << lane) - 1u))) << 9) | d;
        __syncwarp();
    }
    __syncthreads();
    {   // per digit (one thread each): exclusive prefix over the warps, tile total, global base; then the digit prefix of the tile
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { const uint32_t c = cnt[w][d]; cnt[w][d] = run; run += c; }
        const size_t ci = (size_t)d * n_tiles + blockIdx.x;
        gbase[d] = counts_scanned[ci] + tile_offsets[ci / SCAN_TILE];
        uint32_t inc = run;                                   // inclusive scan of the tile's digit counts over 256 threads
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) woff += (w < warp) ? s_wsum[w] : 0u;
        lbase[d] = woff + inc - run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t d = rk[r] & 511u;
        if (d < 256u) {
            const uint32_t lp = lbase[d] + cnt[warp][d] + (rk[r] >> 9);
            s_key[lp] = key[r];
            s_val[lp] = val[r];
        }
    }
    __syncthreads();
    const long long rem = n - t0;
    const int tile_n = rem < RS_TILE ? (int)rem : RS_TILE;
    for (int i = threadIdx.x; i < tile_n; i += RS_THREADS) {
        const uint32_t k = s_key[i];
        const uint32_t d = (k >> shift) & 255u;
        const uint32_t pos = gbase[d] + ((uint32_t)i - lbase[d]);
        keys_out[pos] = k;
        vals_out[pos] = s_val[i];
    }
}

}  // namespace idrk

using namespace idrk;

static long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

namespace idrk {

long long radix_sort_scratch_bytes(long long n) {
    const long long n_tiles = (n + RS_TILE - 1) / RS_TILE;
    const long long m = 256 * n_tiles;
    const long long scan_tiles = (m + SCAN_TILE - 1) / SCAN_TILE;
    if (scan_tiles > SCAN_TILE) return -1;
    return align_up(m * 4, 256) + align_up(scan_tiles * 4, 256);
}

// Stable LSD radix sort of n (key, value) pairs on the low `bits` key bits.  keys[0] / vals[0] hold the input, [1] are the
// ping-pong buffers; the sorted values land in `vals_final` when given (else in vals[*keys_in]), the sorted keys in
// keys[*keys_in].  `scratch` >= radix_sort_scratch_bytes(n).
int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], uint32_t* vals_final, long long n, int bits, void* scratch,
                     cudaStream_t st, int* keys_in) {
    const long long n_tiles = (n + RS_TILE - 1) / RS_TILE;
    const long long m = 256 * n_tiles;
    const long long scan_tiles = (m + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t* counts = (uint32_t*)scratch;
    uint32_t* sums = (uint32_t*)((char*)scratch + align_up(m * 4, 256));
    const int passes = (bits + 7) / 8;
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        IDRK_CUDA_TRY(launch_k(rs_hist_kernel, dim3((unsigned)n_tiles), dim3(RS_THREADS), 0, st, (const uint32_t*)keys[cur],
                               (long long)n, shift, (int)n_tiles, counts));
        IDRK_CUDA_TRY(launch_k(scan_tiles_kernel, dim3((unsigned)scan_tiles), dim3(1024), 0, st, counts, (long long)m, sums));
        IDRK_CUDA_TRY(launch_k(scan_sums_kernel, dim3(1), dim3(1024), 0, st, sums, (int)scan_tiles));
        uint32_t* vout = (p == passes - 1 && vals_final != nullptr) ? vals_final : vals[cur ^ 1];
        IDRK_CUDA_TRY(launch_k(rs_scatter_kernel, dim3((unsigned)n_tiles), dim3(RS_THREADS), 0, st, (const uint32_t*)keys[cur],
                               (const uint32_t*)vals[cur], (long long)n, shift, (int)n_tiles, (const uint32_t*)counts,
                               (const uint32_t*)sums, keys[cur ^ 1], vout));
        cur ^= 1;
    }
    if (keys_in) *keys_in = cur;
    return 0;
}

}  // namespace idrk

extern "C" int idrk_morton_sort_workspace(int64_t n, int64_t* out_bytes) {
    if (!out_bytes || n < 0 || n >= (1LL << 31)) return IDRK_E_ARG;
    const long long sc = radix_sort_scratch_bytes(n);
    if (sc < 0) return IDRK_E_UNSUP;
    *out_bytes = 4 * align_up(n * 4, 256) + sc + 256;
    return 0;
}

extern "C" int idrk_morton_sort(const float* x, int64_t n, int32_t ldx, const float* h_lo, const float* h_hi, int32_t bits_per_dim,
                                int32_t* perm, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!x || !perm || !workspace || n < 0 || ldx < 3 || bits_per_dim < 1 || bits_per_dim > 10) return IDRK_E_ARG;
    if ((h_lo == nullptr) != (h_hi == nullptr)) return IDRK_E_ARG;
    int64_t need = 0;
    int rc = idrk_morton_sort_workspace(n, &need);
    if (rc) return rc;
    if (workspace_bytes < need) return IDRK_E_ARG;
    if (!aligned16(workspace)) return IDRK_E_ALIGN;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace;
    uint32_t* keys[2] = {(uint32_t*)w, (uint32_t*)(w + align_up(n * 4, 256))};
    uint32_t* vals[2] = {(uint32_t*)(w + 2 * align_up(n * 4, 256)), (uint32_t*)(w + 3 * align_up(n * 4, 256))};
    char* scratch = w + 4 * align_up(n * 4, 256);
    uint32_t* box = (uint32_t*)(scratch + radix_sort_scratch_bytes(n));
    const float qmax = (float)((1 << bits_per_dim) - 1);
    float s[3] = {0.f, 0.f, 0.f}, lo[3] = {0.f, 0.f, 0.f};
    long long kb = (n + 255) / 256;
    if (kb > 16LL * sm_count()) kb = 16LL * sm_count();
    if (h_lo != nullptr) {
        for (int d = 0; d < 3; ++d) { const float e = h_hi[d] - h_lo[d]; s[d] = e > 0.f ? (qmax + 1.f) / e : 0.f; lo[d] = h_lo[d]; }
        box = nullptr;
    } else {                                                 // no box given: measure it on the device
        IDRK_CUDA_TRY(launch_k(bbox_init_kernel, dim3(1), dim3(32), 0, st, box));
        long long bb = kb > 4LL * sm_count() ? 4LL * sm_count() : kb;
        IDRK_CUDA_TRY(launch_k(bbox_kernel, dim3((unsigned)bb), dim3(256), 0, st, x, (long long)n, (int)ldx, box));
    }
    IDRK_CUDA_TRY(launch_k(morton_key_kernel, dim3((unsigned)kb), dim3(256), 0, st, x, (long long)n, (int)ldx, lo[0], lo[1], lo[2],
                           s[0], s[1], s[2], qmax, (const uint32_t*)box, keys[0], vals[0]));
    rc = radix_sort_pairs(keys, vals, (uint32_t*)perm, n, 3 * bits_per_dim, scratch, st, nullptr);     // the last pass writes the permutation
    if (rc) return rc;
    IDRK_LAUNCH_CHECK();
    return 0;
}
