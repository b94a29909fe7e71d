// K4: dense contraction tiles for the SDF / rendering MLPs on the 5th-generation tensor cores.
//
//   C[M,N] = epilogue( A . B ),  fp32 in HBM, TF32 (1 pass) or 3xTF32 (hi/lo split, fp32-accurate) MMA,
//   accumulators in TMEM, operands staged by TMA (SWIZZLE_128B) through an mbarrier ring.
//
// Replaces the cuBLAS SGEMMs + separate activation kernels of the reference's eager path:
//   ImplicitNetwork.forward   implicit_differentiable_renderer.py:96-112 (Linear + Softplus(beta=100))
//   RenderingNetwork.forward  implicit_differentiable_renderer.py:215-223 (Linear + ReLU / tanh)
//   and their autograd (grad wrt input = NN form, grad wrt weight = TN form).
//
// Persistent grid (one CTA per SM) walking 128 x BN output tiles (x K splits).  Warp roles (320 threads):
//   warp 0  : TMA producer (one elected lane)        warp 1 : TMEM alloc + tcgen05.mma issuer (one lane)
//   warps 2-9: epilogue - tcgen05.ld the accumulator, bias / activation / hi-lo split, global stores.
// Two TMEM accumulator stages: the epilogue of tile i overlaps the mainloop of tile i+1.
// Layouts (row-major storage):   NT: A[M,K] B[N,K]    NN: A[M,K] B[K,N]    TN: A[K,M] B[K,N]
// K-major operands use the canonical K-major SW128 smem layout, MN-major operands SW128 with a 32-byte base
// (cute/atom/mma_traits_sm100.hpp make_umma_desc documents both), so no transposes are ever materialised.
#include "gemm_common.cuh"

namespace idrk {


template <bool A_MN, bool B_MN, int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    return (1u << 4)                 // D format f32
         | (2u << 7) | (2u << 10)    // A, B format tf32
         | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16)
         | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int BN, int TERMS>
struct SmemPlan {
    static constexpr int A_BYTES = BM * BK * 4;
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = (TERMS == 3 ? 2 : 1) * (A_BYTES + B_BYTES);
    static constexpr int BUDGET = 200 * 1024;
    static constexpr int STAGES = (BUDGET / STAGE_BYTES) > 8 ? 8 : (BUDGET / STAGE_BYTES);
    static constexpr int SCRATCH_BYTES = 0;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + SCRATCH_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};


// two consecutive columns of one row -> every enabled output array
__device__ __forceinline__ void epi_store2(const EpiParams& e, long long row, int col, int N, float v0, float v1, float s0, float s1) {
    const long long o = row * e.ldc + col;
    const bool both = col + 1 < N;
    if (e.accumulate) {
        atomicAdd(e.C + o, v0);
        if (both) atomicAdd(e.C + o + 1, v1);
        return;
    }
    if (both) {
        if (e.C) *reinterpret_cast<float2*>(e.C + o) = make_float2(v0, v1);
        if (e.C_hi) {
            const float h0 = tf32_rn(v0), h1 = tf32_rn(v1);
            *reinterpret_cast<float2*>(e.C_hi + o) = make_float2(h0, h1);
            *reinterpret_cast<float2*>(e.C_lo + o) = make_float2(tf32_rn(v0 - h0), tf32_rn(v1 - h1));
        }
        if (e.S) *reinterpret_cast<float2*>(e.S + row * e.lds + col) = make_float2(s0, s1);
    } else {
        if (e.C) e.C[o] = v0;
        if (e.C_hi) { const float h0 = tf32_rn(v0); e.C_hi[o] = h0; e.C_lo[o] = tf32_rn(v0 - h0); }
        if (e.S) e.S[row * e.lds + col] = s0;
    }
}

// NB 8-column blocks of a 16-row half (rows row0 + g and row0 + g + 8, g = lane / 4)
template <int MODE, int NB>
__device__ __forceinline__ void epi_frag(const EpiParams& e, const uint32_t* r, int lane, long long row0, long long m_eff,
                                         int col0, int N) {
    const int t = lane & 3, g = lane >> 2;
    const long long ra = row0 + g, rb = ra + 8;
    const bool va = ra < m_eff, vb = rb < m_eff;
    const float inv_act = MODE == IDRK_EPI_SOFTPLUS ? 1.f / e.act : 0.f;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int col = col0 + 8 * i + 2 * t;
        if (col >= N) continue;
        float b0 = 0.f, b1 = 0.f;
        if (e.bias != nullptr) { b0 = __ldg(e.bias + col); b1 = (col + 1 < N) ? __ldg(e.bias + col + 1) : 0.f; }
        float v[4], sd[4];
        if constexpr (MODE >= 0) {
            v[0] = epi_fast<MODE>(__uint_as_float(r[4 * i + 0]) + b0, e.act, inv_act, e.scale, sd[0]);
            v[1] = epi_fast<MODE>(__uint_as_float(r[4 * i + 1]) + b1, e.act, inv_act, e.scale, sd[1]);
            v[2] = epi_fast<MODE>(__uint_as_float(r[4 * i + 2]) + b0, e.act, inv_act, e.scale, sd[2]);
            v[3] = epi_fast<MODE>(__uint_as_float(r[4 * i + 3]) + b1, e.act, inv_act, e.scale, sd[3]);
        } else {
            const int c1 = col + 1 < N ? col + 1 : col;
            const long long qa = va ? ra : row0, qb = vb ? rb : row0;
            v[0] = epi_value(e, __uint_as_float(r[4 * i + 0]) + b0, qa, col, sd[0]);
            v[1] = epi_value(e, __uint_as_float(r[4 * i + 1]) + b1, qa, c1, sd[1]);
            v[2] = epi_value(e, __uint_as_float(r[4 * i + 2]) + b0, qb, col, sd[2]);
            v[3] = epi_value(e, __uint_as_float(r[4 * i + 3]) + b1, qb, c1, sd[3]);
        }
        if (va) epi_store2(e, ra, col, N, v[0], v[1], sd[0], sd[1]);
        if (vb) epi_store2(e, rb, col, N, v[2], v[3], sd[2], sd[3]);
    }
}

template <int NB>
__device__ __forceinline__ void epi_frag_dispatch(const EpiParams& e, const uint32_t* r, int lane, long long row0, long long m_eff,
                                                  int col0, int N) {
    switch (e.mode) {
        case IDRK_EPI_SOFTPLUS: epi_frag<IDRK_EPI_SOFTPLUS, NB>(e, r, lane, row0, m_eff, col0, N); break;
        case IDRK_EPI_NONE: epi_frag<IDRK_EPI_NONE, NB>(e, r, lane, row0, m_eff, col0, N); break;
        case IDRK_EPI_RELU: epi_frag<IDRK_EPI_RELU, NB>(e, r, lane, row0, m_eff, col0, N); break;
        default: epi_frag<-1, NB>(e, r, lane, row0, m_eff, col0, N); break;
    }
}

// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+ TMEM alloc),
// warps 2..9 = epilogue (two warps per TMEM lane quarter, each owning half of the BN columns).
template <bool A_MN, bool B_MN, int BN, int TERMS>
__global__ void __launch_bounds__(GEMM_THREADS_V2, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                 long long M, int N, int K, EpiParams e, const int* __restrict__ m_count, int kb_per_split, int splits) {
    pdl_trigger();
    if (m_count != nullptr) {            // device-side row count of zero (gated tracer queries): leave before any set-up
        pdl_wait();
        if (*m_count <= 0) return;
    }
    using P = SmemPlan<BN, TERMS>;
    const int n_tiles = (N + BN - 1) / BN;
    const int kb_total = (K + BK - 1) / BK;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::STAGES * P::STAGE_BYTES + P::SCRATCH_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P::STAGES + 2 * ACC_STAGES);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (P::STAGES + s); };
    auto acc_full_bar = [&](int a) { return bar_base + 8u * (2 * P::STAGES + a); };
    auto acc_empty_bar = [&](int a) { return bar_base + 8u * (2 * P::STAGES + ACC_STAGES + a); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 32) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmAlo); tma_prefetch_desc(&tmBlo); }
    if (threadIdx.x == 0) {
        for (int s = 0; s < P::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(acc_full_bar(a), 1); mbar_init(acc_empty_bar(a), EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)(ACC_STAGES * BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                 // barriers, TMEM and descriptors above overlap the producer's tail; its data is read below
    long long m_eff = M;
    if (m_count != nullptr) { const long long c = *m_count; m_eff = c < M ? c : M; }
    const int m_tiles = (int)((m_eff + BM - 1) / BM);
    const int items = m_tiles * n_tiles * splits;          // CTAs beyond the device-side row count fall through to the teardown

    // work item -> (m tile, n tile, k range); n fastest so concurrently running CTAs share A rows in L2
    auto decode = [&](int item, long long& m0, int& n0, int& kb0, int& kb1) {
        const int z = item % splits;
        const int t = item / splits;
        n0 = (t % n_tiles) * BN;
        m0 = (long long)(t / n_tiles) * BM;
        kb0 = z * kb_per_split;
        kb1 = min(kb_total, kb0 + kb_per_split);
    };

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                long long m0; int n0, kb0, kb1;
                decode(item, m0, n0, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % P::STAGES;
                    const uint32_t ph = (it / P::STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(full_bar(s), P::STAGE_BYTES);
                    const uint32_t st = smem_base + s * P::STAGE_BYTES;
                    const int k0 = kb * BK;
#pragma unroll
                    for (int t = 0; t < (TERMS == 3 ? 2 : 1); ++t) {
                        const uint32_t sA = st + t * (P::A_BYTES + P::B_BYTES);
                        const uint32_t sB = sA + P::A_BYTES;
                        const CUtensorMap* ta = t ? &tmAlo : &tmA;
                        const CUtensorMap* tb = t ? &tmBlo : &tmB;
                        if constexpr (!A_MN) tma_load_2d(sA, ta, full_bar(s), k0, (int)m0);
                        else {
#pragma unroll
                            for (int c = 0; c < BM / 32; ++c) tma_load_2d(sA + c * (BK * 128), ta, full_bar(s), (int)m0 + 32 * c, k0);
                        }
                        if constexpr (!B_MN) tma_load_2d(sB, tb, full_bar(s), k0, n0);
                        else {
#pragma unroll
                            for (int c = 0; c < BN / 32; ++c) tma_load_2d(sB + c * (BK * 128), tb, full_bar(s), n0 + 32 * c, k0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc<A_MN, B_MN, BN>();
            constexpr uint32_t a_lbo = A_MN ? BK * 128 : 16, b_lbo = B_MN ? BK * 128 : 16;
            constexpr uint32_t a_step = A_MN ? (1024 >> 4) : (32 >> 4), b_step = B_MN ? (1024 >> 4) : (32 >> 4);
            constexpr uint32_t a_sbo = A_MN ? 512 : 1024, b_sbo = B_MN ? 512 : 1024;
            constexpr uint32_t a_lt = A_MN ? 1 : 2, b_lt = B_MN ? 1 : 2;
            int it = 0, ti = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++ti) {
                long long m0; int n0, kb0, kb1;
                decode(item, m0, n0, kb0, kb1);
                const int a = ti & 1;
                mbar_wait(acc_empty_bar(a), ((ti >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % P::STAGES;
                    const uint32_t ph = (it / P::STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t st = smem_base + s * P::STAGE_BYTES;
                    const uint64_t a_hi = umma_desc(st, a_lbo, a_sbo, a_lt);
                    const uint64_t b_hi = umma_desc(st + P::A_BYTES, b_lbo, b_sbo, b_lt);
                    const uint64_t a_lo = umma_desc(st + P::A_BYTES + P::B_BYTES, a_lbo, a_sbo, a_lt);
                    const uint64_t b_lo = umma_desc(st + 2 * P::A_BYTES + P::B_BYTES, b_lbo, b_sbo, b_lt);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
                        if constexpr (TERMS == 3) {
                            tc_mma_tf32(d_tmem, a_lo + k * a_step, b_hi + k * b_step, idesc, acc);
                            tc_mma_tf32(d_tmem, a_hi + k * a_step, b_lo + k * b_step, idesc, 1u);
                            tc_mma_tf32(d_tmem, a_hi + k * a_step, b_hi + k * b_step, idesc, 1u);
                        } else {
                            tc_mma_tf32(d_tmem, a_hi + k * a_step, b_hi + k * b_step, idesc, acc);
                        }
                    }
                    tc_commit(empty_bar(s));
                }
                tc_commit(acc_full_bar(a));
            }
        }
    } else {
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;               // which quarter of the BN columns
        constexpr int COLS = BN / 4;                    // 32 or 16 columns per epilogue warp
        constexpr int NB = COLS / 8;
        int ti = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++ti) {
            long long m0; int n0, kb0, kb1;
            decode(item, m0, n0, kb0, kb1);
            const int a = ti & 1;
            mbar_wait(acc_full_bar(a), (ti >> 1) & 1);
            tc_fence_after();
            uint32_t ra[4 * NB], rb[4 * NB];            // rows +0..15 and +16..31 of this warp's quarter
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + half * COLS);
            if constexpr (NB == 4) { tc_ld_16x256b_x4(taddr, ra); tc_ld_16x256b_x4(taddr + (16u << 16), rb); }
            else { tc_ld_16x256b_x2(taddr, ra); tc_ld_16x256b_x2(taddr + (16u << 16), rb); }
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty_bar(a));       // accumulator drained: MMA may reuse it
            const long long row_base = m0 + q * 32;
            const int col0 = n0 + half * COLS;
            if (col0 < N && row_base < m_eff) {
                epi_frag_dispatch<NB>(e, ra, lane, row_base, m_eff, col0, N);
                epi_frag_dispatch<NB>(e, rb, lane, row_base + 16, m_eff, col0, N);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)(ACC_STAGES * BN)) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one 256 x 256 output tile per cluster of two CTAs.  Each CTA stages its own
// 128 rows of A and HALF of the B tile (128 of the 256 columns); the leader CTA issues tcgen05.mma.cta_group::2,
// which reads both CTAs' shared memory, so every SM ingests half the B bytes per FLOP of the single-CTA kernel
// (3xTF32 with fp32 hi/lo operands is bound by L2->SM operand traffic, not by the tensor pipe).  Each CTA keeps
// its 128 x 256 accumulator half in its own TMEM and runs the same epilogue.  NT layout (forward / inference).
// ------------------------------------------------------------------------------------------
constexpr int BN2 = 256;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(dst), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(bar_cluster) : "memory");
}

template <int TERMS>
struct SmemPlan2 {
    static constexpr int A_BYTES = BM * BK * 4;                 // this CTA's 128 rows
    static constexpr int B_BYTES = (BN2 / 2) * BK * 4;          // this CTA's half of the B tile
    static constexpr int STAGE_BYTES = (TERMS == 3 ? 2 : 1) * (A_BYTES + B_BYTES);
    static constexpr int STAGES = TERMS == 3 ? 3 : 6;
    static constexpr int SCRATCH_BYTES = 0;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + SCRATCH_BYTES + 1024 + 256;
};

template <int TERMS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS_V2, 1)
gemm_tf32_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                      const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                      long long M, int N, int K, EpiParams e, const int* __restrict__ m_count) {
    pdl_trigger();
    if (m_count != nullptr) {            // device-side row count of zero (gated tracer queries): leave before any set-up
        pdl_wait();
        if (*m_count <= 0) return;
    }
    using P = SmemPlan2<TERMS>;
    const int n_tiles = (N + BN2 - 1) / BN2;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int kb_total = (K + BK - 1) / BK;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::STAGES * P::STAGE_BYTES + P::SCRATCH_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P::STAGES + 2 * ACC_STAGES);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (P::STAGES + s); };
    auto acc_full_bar = [&](int a) { return bar_base + 8u * (2 * P::STAGES + a); };
    auto acc_empty_bar = [&](int a) { return bar_base + 8u * (2 * P::STAGES + ACC_STAGES + a); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 32) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmAlo); tma_prefetch_desc(&tmBlo); }
    if (threadIdx.x == 0) {
        for (int s = 0; s < P::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(acc_full_bar(a), 1); mbar_init(acc_empty_bar(a), 2 * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)(ACC_STAGES * BN2)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                          // peer barriers initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                 // barriers, TMEM and descriptors above overlap the producer's tail; its data is read below
    long long m_eff = M;
    if (m_count != nullptr) { const long long c = *m_count; m_eff = c < M ? c : M; }
    const int m_tiles = (int)((m_eff + 2 * BM - 1) / (2 * BM));
    const int items = m_tiles * n_tiles;

    auto decode = [&](int item, long long& m0, int& n0) {
        n0 = (item % n_tiles) * BN2;
        m0 = (long long)(item / n_tiles) * (2 * BM) + (long long)rank * BM;     // this CTA's 128 rows
    };

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int item = pair; item < items; item += n_pairs) {
                long long m0; int n0;
                decode(item, m0, n0);
                const int nb = n0 + (int)rank * (BN2 / 2);       // this CTA's half of the B columns
                for (int kb = 0; kb < kb_total; ++kb, ++it) {
                    const int s = it % P::STAGES;
                    const uint32_t ph = (it / P::STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    if (leader) mbar_expect_tx(full_bar(s), 2 * P::STAGE_BYTES);    // bytes of both CTAs
                    const uint32_t fb = mapa_cluster(full_bar(s), 0);               // the leader's barrier
                    const uint32_t st = smem_base + s * P::STAGE_BYTES;
                    const int k0 = kb * BK;
#pragma unroll
                    for (int t = 0; t < (TERMS == 3 ? 2 : 1); ++t) {
                        const uint32_t sA = st + t * (P::A_BYTES + P::B_BYTES);
                        const uint32_t sB = sA + P::A_BYTES;
                        tma_load_2d_2sm(sA, t ? &tmAlo : &tmA, fb, k0, (int)m0);
                        tma_load_2d_2sm(sB, t ? &tmBlo : &tmB, fb, k0, nb);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            // instruction shape M = 256 (both CTAs), N = 256, K = 8, K-major tf32 operands, f32 accumulate
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
            int it = 0, ti = 0;
            for (int item = pair; item < items; item += n_pairs, ++ti) {
                const int a = ti & 1;
                mbar_wait(acc_empty_bar(a), ((ti >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN2);
                for (int kb = 0; kb < kb_total; ++kb, ++it) {
                    const int s = it % P::STAGES;
                    const uint32_t ph = (it / P::STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t st = smem_base + s * P::STAGE_BYTES;
                    const uint64_t a_hi = umma_desc(st, 16, 1024, 2);
                    const uint64_t b_hi = umma_desc(st + P::A_BYTES, 16, 1024, 2);
                    const uint64_t a_lo = umma_desc(st + P::A_BYTES + P::B_BYTES, 16, 1024, 2);
                    const uint64_t b_lo = umma_desc(st + 2 * P::A_BYTES + P::B_BYTES, 16, 1024, 2);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                        if constexpr (TERMS == 3) {
                            tc_mma_tf32_2sm(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, acc);
                            tc_mma_tf32_2sm(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
                            tc_mma_tf32_2sm(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
                        } else {
                            tc_mma_tf32_2sm(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
                        }
                    }
                    tc_commit_2sm(empty_bar(s));                 // frees the stage in BOTH CTAs
                }
                tc_commit_2sm(acc_full_bar(a));                  // accumulator ready in BOTH CTAs
            }
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int COLS = BN2 / 4;                            // 64 columns per epilogue warp
        int ti = 0;
        for (int item = pair; item < items; item += n_pairs, ++ti) {
            long long m0; int n0;
            decode(item, m0, n0);
            const int a = ti & 1;
            mbar_wait(acc_full_bar(a), (ti >> 1) & 1);
            tc_fence_after();
            const long long row_base = m0 + q * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN2 + half * COLS);
            const int col0 = n0 + half * COLS;
            const bool live = col0 < N && row_base < m_eff;
#pragma unroll 1
            for (int hrow = 0; hrow < 2; ++hrow) {               // rows +0..15, then +16..31 (keeps registers < 96)
                uint32_t r[32];
                tc_ld_16x256b_x8(taddr + ((uint32_t)(16 * hrow) << 16), r);
                tc_wait_ld();
                if (hrow == 1) {                                 // last TMEM read of this tile: release the accumulator
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_cluster(acc_empty_bar(a), 0));
                }
                if (live) epi_frag_dispatch<8>(e, r, lane, row_base + 16 * hrow, m_eff, col0, N);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                          // the peer's smem / barriers stay alive until both are done
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)(ACC_STAGES * BN2)) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// FP16-pair variant for the no-grad SDF path ("fp16x2"): every operand is stored as two halves
//     x ~= h + l * 2^-11,   h = fp16(x),  l = fp16((x - h) * 2^11)        (relative error ~2^-22, like 3xTF32)
// so it is 4 bytes per element instead of the 8 of the TF32 hi/lo pair, and the three product terms run on the
// kind::f16 pipe at twice the TF32 rate:   D0 += Ah.Bh,   D1 += Ah.Bl + Al.Bh,   result = D0 + 2^-11 * D1
// (two TMEM accumulators per tile; the scaled low halves stay in fp16's normal range).  Valid where magnitudes are
// benign (|x| < 65504: weights, embeddings and softplus activations of the SDF network), i.e. inference only.
// NT layout, 128 x 128 tiles, BK = 64 halves (128-byte swizzle rows), same warp roles as gemm_tf32_kernel.
// ------------------------------------------------------------------------------------------


struct EpiParamsH {
    float* C; __half* C_h; __half* C_l;
    const float* bias;
    int ldc, ldh;
    int mode; float act; float scale;
    int vec16;                      // C_h / C_l rows 16-byte aligned: quad-transposed 16-byte stores
    const float* dot_w; float* dot_out; int ld_dot;     // fused row-dot (SDF head), see idrk_epilogue_f16_t
};

template <int BN_, int STAGES_>
struct SmemPlanHT {
    static constexpr int A_BYTES = BM * BKH * 2;
    static constexpr int B_BYTES = BN_ * BKH * 2;
    static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
    static constexpr int STAGES = STAGES_;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
};



// One 16-row x 32-column accumulator fragment pair (D0, D1) of a warp -> activation -> fp32 and / or fp16-pair stores.
// bias2[i] = bias of columns col0 + 8 i + 2 t + {0, 1} (prefetched by the caller before the accumulator wait).
// FULL: all 32 columns are inside N (no per-column bounds checks); with e.vec16 the fp16 pair goes out as 16-byte
// stores after a quad transpose (4 B stores write quarter sectors: 4 x the L1 -> L2 requests, which is what bounded
// the epilogue - 36 of 60 us at M = 32700 in the ablation).
template <int MODE, bool FULL>
__device__ __forceinline__ void epi_frag_h(const EpiParamsH& e, const SoftplusC& c, const uint32_t* r0, const uint32_t* r1,
                                           const float2 (&bias2)[4], int lane, long long row0, long long m_eff, int col0, int N) {
    constexpr int NB = 4;
    const int t = lane & 3, g = lane >> 2;
    const long long ra = row0 + g, rb = ra + 8;
    const bool va = ra < m_eff, vb = rb < m_eff;
    if (e.dot_w != nullptr) {
        // fused row-dot: this lane holds 8 columns of rows ra and rb; products are summed per lane in column order,
        // then over the quad's 4 lanes - a fixed order, so the result does not depend on scheduling
        float pa = 0.f, pb = 0.f;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int col = col0 + 8 * i + 2 * t;
            float w0 = 0.f, w1 = 0.f;
            if (FULL || col < N) w0 = __ldg(e.dot_w + col);
            if (FULL || col + 1 < N) w1 = __ldg(e.dot_w + col + 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float z = fmaf(__uint_as_float(r1[4 * i + k]), 1.f / F16S_SCALE, __uint_as_float(r0[4 * i + k])) + ((k & 1) ? bias2[i].y : bias2[i].x);
                const float v = epi_act_h<MODE>(z, c);
                if (k < 2) pa = fmaf(v, (k & 1) ? w1 : w0, pa); else pb = fmaf(v, (k & 1) ? w1 : w0, pb);
            }
        }
        pa += __shfl_xor_sync(0xffffffffu, pa, 1); pb += __shfl_xor_sync(0xffffffffu, pb, 1);
        pa += __shfl_xor_sync(0xffffffffu, pa, 2); pb += __shfl_xor_sync(0xffffffffu, pb, 2);
        if (t == 0) {
            const int j = col0 >> 5;
            if (va) e.dot_out[ra * e.ld_dot + j] = pa;
            if (vb) e.dot_out[rb * e.ld_dot + j] = pb;
        }
        return;
    }
    if (FULL && e.vec16 && e.C == nullptr) {
        uint32_t ha[4], la[4], hb[4], lb[4];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float z = fmaf(__uint_as_float(r1[4 * i + k]), 1.f / F16S_SCALE, __uint_as_float(r0[4 * i + k])) + ((k & 1) ? bias2[i].y : bias2[i].x);
                v[k] = epi_act_h<MODE>(z, c);
            }
            const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
            const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
            ha[i] = h2_bits(h0); hb[i] = h2_bits(h1);
            la[i] = h2_bits(__floats2half2_rn((v[0] - f0.x) * F16S_SCALE, (v[1] - f0.y) * F16S_SCALE));
            lb[i] = h2_bits(__floats2half2_rn((v[2] - f1.x) * F16S_SCALE, (v[3] - f1.y) * F16S_SCALE));
        }
        quad_transpose(ha, t); quad_transpose(la, t); quad_transpose(hb, t); quad_transpose(lb, t);
        const int col = col0 + 8 * t;
        if (va) {
            const long long o = ra * e.ldh + col;
            *reinterpret_cast<uint4*>(e.C_h + o) = make_uint4(ha[0], ha[1], ha[2], ha[3]);
            *reinterpret_cast<uint4*>(e.C_l + o) = make_uint4(la[0], la[1], la[2], la[3]);
        }
        if (vb) {
            const long long o = rb * e.ldh + col;
            *reinterpret_cast<uint4*>(e.C_h + o) = make_uint4(hb[0], hb[1], hb[2], hb[3]);
            *reinterpret_cast<uint4*>(e.C_l + o) = make_uint4(lb[0], lb[1], lb[2], lb[3]);
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int col = col0 + 8 * i + 2 * t;
        if (!FULL && col >= N) continue;
        const bool both = FULL || col + 1 < N;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float z = fmaf(__uint_as_float(r1[4 * i + k]), 1.f / F16S_SCALE, __uint_as_float(r0[4 * i + k])) + ((k & 1) ? bias2[i].y : bias2[i].x);
            v[k] = epi_act_h<MODE>(z, c);
        }
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const long long row = hrow ? rb : ra;
            if (!(hrow ? vb : va)) continue;
            const float x0 = v[2 * hrow], x1 = v[2 * hrow + 1];
            if (e.C) {
                if (both) *reinterpret_cast<float2*>(e.C + row * e.ldc + col) = make_float2(x0, x1);
                else e.C[row * e.ldc + col] = x0;
            }
            if (e.C_h) {
                const __half2 h = __floats2half2_rn(x0, x1);
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn((x0 - hf.x) * F16S_SCALE, (x1 - hf.y) * F16S_SCALE);
                const long long o = row * e.ldh + col;
                if (both) {
                    *reinterpret_cast<__half2*>(e.C_h + o) = h;
                    *reinterpret_cast<__half2*>(e.C_l + o) = l;
                } else { e.C_h[o] = __low2half(h); e.C_l[o] = __low2half(l); }
            }
        }
    }
}

// bias of this lane's 4 column pairs of the warp's 32-column group (zero outside N / without a bias vector)
__device__ __forceinline__ void load_bias2(const float* __restrict__ bias, int col0, int lane, int N, float2 (&b)[4]) {
    const int t = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int col = col0 + 8 * i + 2 * t;
        b[i] = make_float2(0.f, 0.f);
        if (bias != nullptr && col < N) { b[i].x = __ldg(bias + col); if (col + 1 < N) b[i].y = __ldg(bias + col + 1); }
    }
}

// Template parameters: BN_ output columns per tile, STAGES_ operand stages, ACCS TMEM accumulator stages, EPIW epilogue
// warps (4 lane quarters x BN_/32 column groups), MINB resident CTAs per SM.
//   <128, 3, 2, 16, 1> is the shipped configuration: persistent, 192 KB of operand stages, epilogue(i) overlaps
//   mainloop(i+1).  Measured alternative for one-wave launches (M = 4096): <64, 2, 1, 8, 2>, two CTAs per SM so that one
//   CTA's epilogue and the next launch's prologue overlap the other's mainloop - 11.0 us vs 9.7 us, slower (more
//   shared-memory traffic per MMA cycle at BN = 64), not instantiated.
template <int BN_, int STAGES_, int ACCS, int EPIW, int MINB>
__global__ void __launch_bounds__(64 + 32 * EPIW, MINB)
gemm_f16s_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAl,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBl,
                 long long M, int N, int K, EpiParamsH e, const int* __restrict__ m_count) {
    pdl_trigger();
    if (m_count != nullptr) {            // device-side row count of zero (gated tracer queries): leave before any set-up
        pdl_wait();
        if (*m_count <= 0) return;
    }
    using P = SmemPlanHT<BN_, STAGES_>;
    const int n_tiles = (N + BN_ - 1) / BN_;
    const int kb_total = (K + BKH - 1) / BKH;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::STAGES * P::STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P::STAGES + 2 * ACCS);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (P::STAGES + s); };
    auto acc_full_bar = [&](int a) { return bar_base + 8u * (2 * P::STAGES + a); };
    auto acc_empty_bar = [&](int a) { return bar_base + 8u * (2 * P::STAGES + ACCS + a); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 32) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmBl); }
    if (threadIdx.x == 0) {
        for (int s = 0; s < P::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < ACCS; ++a) { mbar_init(acc_full_bar(a), 1); mbar_init(acc_empty_bar(a), EPIW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr uint32_t TMEM_COLS = ACCS * 2 * BN_;         // two accumulators (D0, D1) per stage
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                 // barriers, TMEM and descriptors above overlap the producer's tail; its data is read below
    long long m_eff = M;
    if (m_count != nullptr) { const long long c = *m_count; m_eff = c < M ? c : M; }
    const int m_tiles = (int)((m_eff + BM - 1) / BM);
    const int items = m_tiles * n_tiles;

    if (warp == 0) {
        {
            const uint32_t leader = elect_one();
            int it = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int n0 = (item % n_tiles) * BN_;
                const int m0 = (item / n_tiles) * BM;
                for (int kb = 0; kb < kb_total; ++kb, ++it) {
                    const int s = it % P::STAGES;
                    const uint32_t ph = (it / P::STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx_p(full_bar(s), P::STAGE_BYTES, leader);
                    const uint32_t st = smem_base + s * P::STAGE_BYTES;
                    const int k0 = kb * BKH;
                    tma_load_2d_p(st, &tmA, full_bar(s), k0, m0, leader);
                    tma_load_2d_p(st + 2 * P::A_BYTES, &tmB, full_bar(s), k0, n0, leader);
                    tma_load_2d_p(st + P::A_BYTES, &tmAl, full_bar(s), k0, m0, leader);
                    tma_load_2d_p(st + 2 * P::A_BYTES + P::B_BYTES, &tmBl, full_bar(s), k0, n0, leader);
                }
            }
        }
    } else if (warp == 1) {
        {
            const uint32_t leader = elect_one();
            // kind::f16: A, B fp16 (format 0), f32 accumulate, K-major, M = 128, N = 128, K = 16 per instruction
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN_ >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * BN_) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int it = 0, ti = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++ti) {
                const int a = ti % ACCS;
                mbar_wait(acc_empty_bar(a), ((ti / ACCS) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(a * 2 * BN_);
                const uint32_t d1 = d0 + BN_;
                for (int kb = 0; kb < kb_total; ++kb, ++it) {
                    const int s = it % P::STAGES;
                    const uint32_t ph = (it / P::STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t st = smem_base + s * P::STAGE_BYTES;
                    // stage = [A_h | A_l | B_h | B_l]: B_h and B_l are adjacent, so [B_h ; B_l] is one 256-row K-major operand
                    // and [D0 | D1] one 256-column accumulator - Ah.Bh and Ah.Bl issue as ONE N = 256 instruction that reads
                    // A_h from shared memory once (shared-memory bandwidth, MMA operand reads + TMA fills, bounds this kernel)
                    const uint64_t a_h = umma_desc(st, 16, 1024, 2);
                    const uint64_t a_l = umma_desc(st + P::A_BYTES, 16, 1024, 2);
                    const uint64_t b_h = umma_desc(st + 2 * P::A_BYTES, 16, 1024, 2);
#pragma unroll
                    for (int k = 0; k < BKH / 16; ++k) {
                        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                        tc_mma_f16_p(d0, a_h + 2 * k, b_h + 2 * k, idesc2, acc, leader);     // [D0 | D1] (+)= Ah . [Bh ; Bl]^T
                        tc_mma_f16_p(d1, a_l + 2 * k, b_h + 2 * k, idesc, 1u, leader);       //  D1      +=  Al . Bh^T
                    }
                    tc_commit_p(empty_bar(s), leader);
                }
                tc_commit_p(acc_full_bar(a), leader);
            }
        }
    } else {
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;                        // 32-column group of this warp
        const SoftplusC spc = {e.act * 1.4426950408889634f, e.mode == IDRK_EPI_SOFTPLUS ? 0.6931471805599453f / e.act * e.scale : 0.f, e.scale};
        int ti = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++ti) {
            const int n0 = (item % n_tiles) * BN_;
            const long long m0 = (long long)(item / n_tiles) * BM;
            const int a = ti % ACCS;
            const int col0 = n0 + grp * 32;
            float2 bias2[4];
            load_bias2(e.bias, col0, lane, N, bias2);            // in flight while the accumulator is still being produced
            mbar_wait(acc_full_bar(a), (ti / ACCS) & 1);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 2 * BN_ + grp * 32);
            uint32_t a0[16], a1[16], b0[16], b1[16];            // D0 / D1 fragments of rows +0..15 and +16..31
            tc_ld_16x256b_x4(t0, a0);
            tc_ld_16x256b_x4(t0 + BN_, a1);
            tc_ld_16x256b_x4(t0 + (16u << 16), b0);
            tc_ld_16x256b_x4(t0 + (16u << 16) + BN_, b1);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty_bar(a));
            const long long row_base = m0 + q * 32;
            if (col0 >= N || row_base >= m_eff) continue;
            const bool full = col0 + 32 <= N;
            if (e.mode == IDRK_EPI_SOFTPLUS) {
                if (full) {
                    epi_frag_h<IDRK_EPI_SOFTPLUS, true>(e, spc, a0, a1, bias2, lane, row_base, m_eff, col0, N);
                    epi_frag_h<IDRK_EPI_SOFTPLUS, true>(e, spc, b0, b1, bias2, lane, row_base + 16, m_eff, col0, N);
                } else {
                    epi_frag_h<IDRK_EPI_SOFTPLUS, false>(e, spc, a0, a1, bias2, lane, row_base, m_eff, col0, N);
                    epi_frag_h<IDRK_EPI_SOFTPLUS, false>(e, spc, b0, b1, bias2, lane, row_base + 16, m_eff, col0, N);
                }
            } else {
                epi_frag_h<IDRK_EPI_NONE, false>(e, spc, a0, a1, bias2, lane, row_base, m_eff, col0, N);
                epi_frag_h<IDRK_EPI_NONE, false>(e, spc, b0, b1, bias2, lane, row_base + 16, m_eff, col0, N);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// x -> fp16 pair (h, l); optionally a second pair (h2, l2) of scale2 * x in another buffer in the same pass (the skip
// connection's copy of the embedding inside the layer-4 operand).
__global__ void split_f16_kernel(const float* __restrict__ x, long long rows, int cols, int ldx, float scale,
                                 __half* __restrict__ h, __half* __restrict__ l, int ldo, int pad_cols,
                                 __half* __restrict__ h2, __half* __restrict__ l2, int ldo2, int pad_cols2, float scale2,
                                 const int* __restrict__ m_count) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    long long r_eff = rows;
    if (m_count) { const long long c = *m_count; r_eff = c < rows ? c : rows; }
    const int pmax = (h2 && pad_cols2 > pad_cols) ? pad_cols2 : pad_cols;
    const int w = cols + pmax;
    const long long total = r_eff * (long long)w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / w;
        const int c = (int)(i - r * w);
        const float xv = c < cols ? x[r * ldx + c] : 0.f;
        if (c < cols + pad_cols) {
            const float v = fminf(fmaxf(xv * scale, -65504.f), 65504.f);
            const __half hv = __float2half_rn(v);
            h[r * ldo + c] = hv;
            l[r * ldo + c] = __float2half_rn((v - __half2float(hv)) * F16S_SCALE);
        }
        if (h2 != nullptr && c < cols + pad_cols2) {
            const float v = fminf(fmaxf(xv * scale2, -65504.f), 65504.f);
            const __half hv = __float2half_rn(v);
            h2[r * ldo2 + c] = hv;
            l2[r * ldo2 + c] = __float2half_rn((v - __half2float(hv)) * F16S_SCALE);
        }
    }
}

// ------------------------------------------------------------------------------------------
// plain fp32 FFMA tiles: exact-fp32 mode (parity debugging, tiny shapes) - same epilogue
// ------------------------------------------------------------------------------------------
template <int TM, int TN>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long sa_m, long long sa_k, const float* __restrict__ B, long long sb_n,
                long long sb_k, long long M, int N, int K, EpiParams e, const int* __restrict__ m_count, int k_per_split) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    constexpr int SBM = 16 * TM, SBN = 16 * TN, SBK = 16;
    __shared__ float sA[SBK][SBM + 1];
    __shared__ float sB[SBK][SBN + 1];
    long long m_eff = M;
    if (m_count != nullptr) { const long long c = *m_count; m_eff = c < M ? c : M; }
    const long long m0 = (long long)blockIdx.x * SBM;
    if (m0 >= m_eff) return;
    const int n0 = blockIdx.y * SBN;
    const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int k0 = k_begin; k0 < k_end; k0 += SBK) {
        for (int i = threadIdx.x; i < SBM * SBK; i += 256) {
            int mm, kk;
            if (sa_k == 1) { kk = i % SBK; mm = i / SBK; } else { mm = i % SBM; kk = i / SBM; }
            const long long gm = m0 + mm; const int gk = k0 + kk;
            sA[kk][mm] = (gm < M && gk < k_end) ? A[gm * sa_m + gk * sa_k] : 0.f;
        }
        for (int i = threadIdx.x; i < SBN * SBK; i += 256) {
            int nn, kk;
            if (sb_k == 1) { kk = i % SBK; nn = i / SBK; } else { nn = i % SBN; kk = i / SBN; }
            const int gn = n0 + nn; const int gk = k0 + kk;
            sB[kk][nn] = (gn < N && gk < k_end) ? B[gn * sb_n + gk * sb_k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SBK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = sA[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = sB[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long row = m0 + ty + 16 * i;
        if (row >= m_eff) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = n0 + tx + 16 * j;
            if (col >= N) continue;
            float s;
            const float z = acc[i][j] + (e.bias ? e.bias[col] : 0.f);
            const float v = epi_value(e, z, row, col, s);
            epi_store(e, row, col, v, s);
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------

// 2-D fp32 tensor map: dim0 = contiguous extent, dim1 = rows, row pitch ld floats; box (32, box1), SW128
static int make_tmap(CUtensorMap* tm, const float* base, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box1,
                     bool mn_major = false) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return IDRK_E_DRIVER;
    if (!aligned16(base) || (ld & 3)) return IDRK_E_ALIGN;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {32, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : IDRK_E_ARG;
}

template <bool A_MN, bool B_MN, int BN, int TERMS>
static int launch_tc(const CUtensorMap& tA, const CUtensorMap& tAl, const CUtensorMap& tB, const CUtensorMap& tBl,
                     long long M, int N, int K, const EpiParams& e, const int* m_count, int splits, cudaStream_t st) {
    using P = SmemPlan<BN, TERMS>;
    auto kern = gemm_tf32_kernel<A_MN, B_MN, BN, TERMS>;
    static bool attr_done = false;
    if (!attr_done) {
        IDRK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::TOTAL));
        attr_done = true;
    }
    const int kb_total = (K + BK - 1) / BK;
    int kbps = (kb_total + splits - 1) / splits;
    splits = (kb_total + kbps - 1) / kbps;
    const long long items = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * splits;
    const long long grid = items < sm_count() ? items : sm_count();      // persistent: one CTA per SM
    IDRK_CUDA_TRY(launch_k(kern, dim3((unsigned)grid), dim3(GEMM_THREADS_V2), P::TOTAL, st, tA, tAl, tB, tBl, M, N, K, e, m_count, kbps, splits));
    IDRK_LAUNCH_CHECK();
    return 0;
}

static int make_tmap_h(CUtensorMap* tm, const void* base, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box1) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return IDRK_E_DRIVER;
    if (!aligned16(base) || (ld & 7)) return IDRK_E_ALIGN;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : IDRK_E_ARG;
}

template <int TERMS>
static int launch_2cta(const CUtensorMap& tA, const CUtensorMap& tAl, const CUtensorMap& tB, const CUtensorMap& tBl,
                       long long M, int N, int K, const EpiParams& e, const int* m_count, cudaStream_t st) {
    using P = SmemPlan2<TERMS>;
    auto kern = gemm_tf32_2cta_kernel<TERMS>;
    static bool attr_done = false;
    if (!attr_done) {
        IDRK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::TOTAL));
        attr_done = true;
    }
    const long long items = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN2 - 1) / BN2);
    long long pairs = sm_count() / 2;
    if (pairs > items) pairs = items;
    IDRK_CUDA_TRY(launch_k(kern, dim3((unsigned)(2 * pairs)), dim3(GEMM_THREADS_V2), P::TOTAL, st, tA, tAl, tB, tBl, M, N, K, e, m_count));
    IDRK_LAUNCH_CHECK();
    return 0;
}

static bool use_2cta() {
    static int v = -1;
    if (v < 0) { const char* s = getenv("IDRK_GEMM_2CTA"); v = (s == nullptr || s[0] != '0') ? 1 : 0; }
    return v == 1;
}

}  // namespace idrk

using namespace idrk;

extern "C" int idrk_gemm(int32_t layout, int32_t precision, int64_t M, int32_t N, int32_t K,
                         const float* A, const float* A_lo, int32_t lda, const float* B, const float* B_lo, int32_t ldb,
                         const idrk_epilogue_t* h_epi, const int32_t* m_count, int32_t split_k, void* stream) {
    if (!A || !B || !h_epi || M < 0 || N < 1 || K < 1 || lda < 1 || ldb < 1) return IDRK_E_ARG;
    if (layout < IDRK_GEMM_NT || layout > IDRK_GEMM_TN) return IDRK_E_ARG;
    if (precision != IDRK_PREC_FP32 && precision != IDRK_PREC_TF32 && precision != IDRK_PREC_3XTF32) return IDRK_E_ARG;
    if (M == 0) return 0;
    EpiParams e;
    e.C = h_epi->C; e.C_hi = h_epi->C_hi; e.C_lo = h_epi->C_lo; e.S = h_epi->S;
    e.bias = h_epi->bias; e.aux = h_epi->aux; e.ldc = h_epi->ldc; e.lds = h_epi->lds; e.ldaux = h_epi->ldaux;
    e.mode = h_epi->mode; e.act = h_epi->act_param; e.scale = h_epi->scale; e.accumulate = h_epi->accumulate;
    if (split_k < 1) split_k = 1;
    if (split_k > 1) e.accumulate = 1;
    if (e.accumulate && (!e.C || e.C_hi || e.S || e.mode != IDRK_EPI_NONE)) return IDRK_E_ARG;
    if (!e.C && !e.C_hi) return IDRK_E_ARG;
    if ((e.C_hi == nullptr) != (e.C_lo == nullptr)) return IDRK_E_ARG;
    if (e.mode == IDRK_EPI_MUL_AUX && !e.aux) return IDRK_E_ARG;
    if (e.ldc < N) return IDRK_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool a_mn = layout == IDRK_GEMM_TN, b_mn = layout != IDRK_GEMM_NT;

    if (precision == IDRK_PREC_FP32) {
        const long long sa_m = a_mn ? 1 : lda, sa_k = a_mn ? lda : 1;
        const long long sb_n = b_mn ? 1 : ldb, sb_k = b_mn ? ldb : 1;
        int kps = (K + split_k - 1) / split_k;
        kps = (kps + 15) / 16 * 16;
        const int splits = (K + kps - 1) / kps;
        dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64), (unsigned)splits);
        IDRK_CUDA_TRY(launch_k(gemm_f32_kernel<4, 4>, dim3(grid), dim3(256), 0, st, A, sa_m, sa_k, B, sb_n, sb_k, M, N, K, e, m_count, kps));
        IDRK_LAUNCH_CHECK();
        return 0;
    }

    const int terms = precision == IDRK_PREC_3XTF32 ? 3 : 1;
    if (terms == 3 && (!A_lo || !B_lo)) return IDRK_E_ARG;
    if (layout == IDRK_GEMM_NT && M >= 16384 && N >= 256 && split_k == 1 && use_2cta()) {
        // large forward / inference batches: CTA-pair kernel (256 x 256 tiles, half the B ingress per SM).  Measured
        // (3xTF32, N = K = 512, back-to-back launches): M = 8192 31.7 vs 27.6 us for the single-CTA kernel, 16384 48.0 vs
        // 48.7, 32768 77.6 vs 83.8 - it pays from ~16 K rows
        CUtensorMap tA, tAl, tB, tBl;
        int rc;
        if ((rc = make_tmap(&tA, A, (uint64_t)K, (uint64_t)M, lda, BM))) return rc;
        if ((rc = make_tmap(&tB, B, (uint64_t)K, (uint64_t)N, ldb, BN2 / 2))) return rc;
        if (terms == 3) {
            if ((rc = make_tmap(&tAl, A_lo, (uint64_t)K, (uint64_t)M, lda, BM))) return rc;
            if ((rc = make_tmap(&tBl, B_lo, (uint64_t)K, (uint64_t)N, ldb, BN2 / 2))) return rc;
            return launch_2cta<3>(tA, tAl, tB, tBl, M, N, K, e, m_count, st);
        }
        return launch_2cta<1>(tA, tA, tB, tB, M, N, K, e, m_count, st);
    }
    // Operand ingest (L2 -> shared memory), not the tensor pipe, bounds these launches: a 128 x 64 tile moves 48 KB per
    // 32-wide k block against 384 MMA cycles, a 128 x 128 tile 64 KB against 768.  Measured (graph of back-to-back
    // launches, 3xTF32, N = K = 512): M = 2048 -> 13.5 us (BN 64) vs 16.3 us (BN 128, 64 tiles leave SMs idle);
    // M = 3072 -> 21.5 vs 16.6 us; M = 4096 -> 21.9 vs 16.7 us.
    const int bn = (N <= 64 || M <= 2048) ? 64 : 128;
    CUtensorMap tA, tAl, tB, tBl;
    int rc;
    // K-major: dims (K, rows) box (32, rows_per_tile).  MN-major: dims (rows_mn, K) box (32, BK)
    if ((rc = a_mn ? make_tmap(&tA, A, (uint64_t)M, (uint64_t)K, lda, BK, true) : make_tmap(&tA, A, (uint64_t)K, (uint64_t)M, lda, BM))) return rc;
    if ((rc = b_mn ? make_tmap(&tB, B, (uint64_t)N, (uint64_t)K, ldb, BK, true) : make_tmap(&tB, B, (uint64_t)K, (uint64_t)N, ldb, bn))) return rc;
    if (terms == 3) {
        if ((rc = a_mn ? make_tmap(&tAl, A_lo, (uint64_t)M, (uint64_t)K, lda, BK, true) : make_tmap(&tAl, A_lo, (uint64_t)K, (uint64_t)M, lda, BM))) return rc;
        if ((rc = b_mn ? make_tmap(&tBl, B_lo, (uint64_t)N, (uint64_t)K, ldb, BK, true) : make_tmap(&tBl, B_lo, (uint64_t)K, (uint64_t)N, ldb, bn))) return rc;
    } else { tAl = tA; tBl = tB; }

#define IDRK_TC(AM, BMN)                                                                                   \
    do {                                                                                                    \
        if (bn == 64) return terms == 3 ? launch_tc<AM, BMN, 64, 3>(tA, tAl, tB, tBl, M, N, K, e, m_count, split_k, st)  \
                                        : launch_tc<AM, BMN, 64, 1>(tA, tAl, tB, tBl, M, N, K, e, m_count, split_k, st); \
        return terms == 3 ? launch_tc<AM, BMN, 128, 3>(tA, tAl, tB, tBl, M, N, K, e, m_count, split_k, st)               \
                          : launch_tc<AM, BMN, 128, 1>(tA, tAl, tB, tBl, M, N, K, e, m_count, split_k, st);              \
    } while (0)
    if (layout == IDRK_GEMM_NT) IDRK_TC(false, false);
    if (layout == IDRK_GEMM_NN) IDRK_TC(false, true);
    IDRK_TC(true, true);
#undef IDRK_TC
}

extern "C" int idrk_gemm_f16s(int64_t M, int32_t N, int32_t K, const void* A_h, const void* A_l, int32_t lda,
                              const void* B_h, const void* B_l, int32_t ldb, const idrk_epilogue_f16_t* h_epi,
                              const int32_t* m_count, void* stream) {
    if (!A_h || !A_l || !B_h || !B_l || !h_epi || M < 0 || N < 1 || K < 1) return IDRK_E_ARG;
    if (M == 0) return 0;
    EpiParamsH e;
    e.C = h_epi->C; e.C_h = (__half*)h_epi->C_h; e.C_l = (__half*)h_epi->C_l; e.bias = h_epi->bias;
    e.ldc = h_epi->ldc; e.ldh = h_epi->ldh; e.mode = h_epi->mode; e.act = h_epi->act_param; e.scale = h_epi->scale;
    e.dot_w = h_epi->dot_w; e.dot_out = h_epi->dot_out; e.ld_dot = h_epi->ld_dot;
    if (e.dot_w) {
        if (e.C || e.C_h || h_epi->C_l || !e.dot_out || e.ld_dot < (N + 31) / 32) return IDRK_E_ARG;
    } else if (!e.C && !e.C_h) return IDRK_E_ARG;
    e.vec16 = e.C_h && e.C_l && (e.ldh % 8) == 0 && aligned16(e.C_h) && aligned16(e.C_l);
    if ((e.C_h == nullptr) != (e.C_l == nullptr)) return IDRK_E_ARG;
    if (e.mode != IDRK_EPI_NONE && e.mode != IDRK_EPI_SOFTPLUS) return IDRK_E_UNSUP;
    if ((e.C && (e.ldc < N || (e.ldc & 1))) || (e.C_h && (e.ldh < N || (e.ldh & 1)))) return IDRK_E_ARG;
    CUtensorMap tA, tAl, tB, tBl;
    int rc;
    if ((rc = make_tmap_h(&tA, A_h, (uint64_t)K, (uint64_t)M, lda, BM))) return rc;
    if ((rc = make_tmap_h(&tAl, A_l, (uint64_t)K, (uint64_t)M, lda, BM))) return rc;
    const long long m_tiles = (M + BM - 1) / BM;
    auto kern = gemm_f16s_kernel<BNH, 3, 2, 16, 1>;
    using P = SmemPlanHT<BNH, 3>;
    if ((rc = make_tmap_h(&tB, B_h, (uint64_t)K, (uint64_t)N, ldb, BNH))) return rc;
    if ((rc = make_tmap_h(&tBl, B_l, (uint64_t)K, (uint64_t)N, ldb, BNH))) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        IDRK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::TOTAL));
        attr_done = true;
    }
    const long long items = m_tiles * ((N + BNH - 1) / BNH);
    const long long grid = items < sm_count() ? items : sm_count();
    IDRK_CUDA_TRY(launch_k(kern, dim3((unsigned)grid), dim3(GEMM_THREADS_V2), P::TOTAL, (cudaStream_t)stream, tA, tAl, tB, tBl, M, N, K, e, m_count));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_split_f16(const float* x, int64_t rows, int32_t cols, int32_t ldx, float scale, void* h, void* l,
                              int32_t ld_out, int32_t pad_cols, void* h2, void* l2, int32_t ld_out2, int32_t pad_cols2,
                              float scale2, const int32_t* m_count, void* stream) {
    if (!x || !h || !l || rows < 0 || cols < 1 || ldx < cols || pad_cols < 0 || ld_out < cols + pad_cols) return IDRK_E_ARG;
    if ((h2 == nullptr) != (l2 == nullptr)) return IDRK_E_ARG;
    if (h2 && (pad_cols2 < 0 || ld_out2 < cols + pad_cols2)) return IDRK_E_ARG;
    if (rows == 0) return 0;
    const int pmax = (h2 && pad_cols2 > pad_cols) ? pad_cols2 : pad_cols;
    long long total = rows * (long long)(cols + pmax);
    long long b = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    IDRK_CUDA_TRY(launch_k(split_f16_kernel, dim3((int)b), dim3(256), 0, (cudaStream_t)stream, x, rows, cols, ldx, scale,
                           (__half*)h, (__half*)l, ld_out, pad_cols, (__half*)h2, (__half*)l2, ld_out2, pad_cols2, scale2, m_count));
    IDRK_LAUNCH_CHECK();
    return 0;
}
