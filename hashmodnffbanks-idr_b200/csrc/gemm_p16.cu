// K4p: 16-bit-pair contraction tiles for the DIFFERENTIABLE path of the SDF / rendering MLPs (training step).
//
//   C[M,N] = epilogue( A . B ),   every operand a pair of 16-bit floats  x ~= h + l * 2^-11
//       fp16 pair  (h = fp16(x), l = fp16((x - h) 2^11)): ~22 significant bits, |x| < 65504 - weights, activations
//       bf16 pair  (h = bf16(x), l = bf16((x - h) 2^11)): ~17 significant bits, the full fp32 range - cotangents
//   three tcgen05.mma.kind::f16 per product term set:  D0 += Ah.Bh,  D1 += Ah.Bl + Al.Bh,  result = D0 + 2^-11 D1
//   (A and B of one launch share the format: a kind::f16 instruction with an fp16 and a bf16 operand faults on
//   sm_100a - measured, cudaErrorIllegalInstruction - so mixed pairs are rejected with IDRK_E_UNSUP.)
//
// Replaces the 3xTF32 path (8 bytes / element, tf32 rate) for the launches of the recorded forward / backward /
// double backward of implicit_differentiable_renderer.py:96-128,215-223: those launches are one wave of 128 x 128
// tiles and run at the L2 -> SM ingest limit (128 CTAs x 1 MB of operands = 10.8 TB/s of the ~12 TB/s L2 cap), so
// halving the operand bytes - not more tensor throughput - is what makes them faster.
//
// Layouts (row-major storage):   NT: A[M,K] B[N,K]    NN: A[M,K] B[K,N]    TN: A[K,M] B[K,N]
// K-major operands: canonical SWIZZLE_128B K-major tiles (64 halves = 128-byte rows); MN-major operands:
// SWIZZLE_128B MN-major ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)), one TMA box of 64 (MN) x 64 (K) per 8 KB chunk -
// no transposes are ever materialised.  Warp roles as in gemm.cu: warp 0 TMA producer, warp 1 MMA issuer,
// 16 epilogue warps; two TMEM accumulator stages.
#include "gemm_common.cuh"

namespace idrk {

struct EpiParamsP {
    float* C; float* S; void* C_h; void* C_l;
    const float* bias; const float* aux;
    int ldc, lds, ldh, ldaux;
    int c_fmt;                      // format of the (C_h, C_l) pair: 0 fp16, 1 bf16
    int mode; float act; float scale; int accumulate;
};

template <int BN_>
struct SmemPlanP {
    static constexpr int A_BYTES = BM * BKH * 2;          // 16 KB: one of (A_h, A_l)
    static constexpr int B_BYTES = BN_ * BKH * 2;
    static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
    static constexpr int STAGES = (196 * 1024) / STAGE_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ uint32_t pack_pair16(float x0, float x1, int fmt) {
    if (fmt == 0) { const __half2 h = __floats2half2_rn(x0, x1); return *reinterpret_cast<const uint32_t*>(&h); }
    const __nv_bfloat162 b = __floats2bfloat162_rn(x0, x1);
    return *reinterpret_cast<const uint32_t*>(&b);
}
__device__ __forceinline__ float2 unpack_pair16(uint32_t u, int fmt) {
    if (fmt == 0) return __half22float2(*reinterpret_cast<const __half2*>(&u));
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
// (h, l) words of two neighbouring columns; fp16 values are clamped to the format's range first
__device__ __forceinline__ void split_pair16(float x0, float x1, int fmt, uint32_t& h, uint32_t& l) {
    if (fmt == 0) { x0 = fminf(fmaxf(x0, -65504.f), 65504.f); x1 = fminf(fmaxf(x1, -65504.f), 65504.f); }
    h = pack_pair16(x0, x1, fmt);
    const float2 hf = unpack_pair16(h, fmt);
    l = pack_pair16((x0 - hf.x) * F16S_SCALE, (x1 - hf.y) * F16S_SCALE, fmt);
}

// One 16-row x (8 NB)-column accumulator fragment pair (D0, D1) of a warp.  Fragment layout as in gemm.cu
// (tcgen05.ld 16x256b): lane (g = lane / 4, t = lane % 4) holds, per 8-column block i, columns 8 i + 2 t + {0, 1}
// of rows g (regs 4 i + 0..1) and g + 8 (regs 4 i + 2..3).
template <int MODE, int NB>
__device__ __forceinline__ void epi_frag_p(const EpiParamsP& e, const uint32_t* r0, const uint32_t* r1, int lane, long long row0,
                                           long long m_eff, int col0, int N) {
    const int t = lane & 3, g = lane >> 2;
    const long long ra = row0 + g, rb = ra + 8;
    const bool va = ra < m_eff, vb = rb < m_eff;
    const float inv_act = MODE == IDRK_EPI_SOFTPLUS ? 1.f / e.act : 0.f;
    const bool full = col0 + 8 * NB <= N;
    uint32_t ha[NB], la[NB], hb[NB], lb[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int col = col0 + 8 * i + 2 * t;
        ha[i] = la[i] = hb[i] = lb[i] = 0u;
        if (col >= N) continue;
        const bool both = col + 1 < N;
        float b0 = 0.f, b1 = 0.f;
        if (e.bias != nullptr) { b0 = __ldg(e.bias + col); b1 = both ? __ldg(e.bias + col + 1) : 0.f; }
        float z[4], v[4], sd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            z[k] = fmaf(__uint_as_float(r1[4 * i + k]), 1.f / F16S_SCALE, __uint_as_float(r0[4 * i + k])) + ((k & 1) ? b1 : b0);
        if constexpr (MODE >= 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = epi_fast<MODE>(z[k], e.act, inv_act, e.scale, sd[k]);
        } else {
            EpiParams g4;                              // the generic (switch) path of gemm.cu for sine / tanh / x aux
            g4.aux = e.aux; g4.ldaux = e.ldaux; g4.mode = e.mode; g4.act = e.act; g4.scale = e.scale;
            const int c1 = both ? col + 1 : col;
            const long long qa = va ? ra : row0, qb = vb ? rb : row0;
            v[0] = epi_value(g4, z[0], qa, col, sd[0]);
            v[1] = epi_value(g4, z[1], qa, c1, sd[1]);
            v[2] = epi_value(g4, z[2], qb, col, sd[2]);
            v[3] = epi_value(g4, z[3], qb, c1, sd[3]);
        }
        if (e.accumulate) {
            if (va) { atomicAdd(e.C + ra * e.ldc + col, v[0]); if (both) atomicAdd(e.C + ra * e.ldc + col + 1, v[1]); }
            if (vb) { atomicAdd(e.C + rb * e.ldc + col, v[2]); if (both) atomicAdd(e.C + rb * e.ldc + col + 1, v[3]); }
            continue;
        }
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            if (!(hrow ? vb : va)) continue;
            const long long row = hrow ? rb : ra;
            const float x0 = v[2 * hrow], x1 = v[2 * hrow + 1];
            if (e.C) {
                if (both) *reinterpret_cast<float2*>(e.C + row * e.ldc + col) = make_float2(x0, x1);
                else e.C[row * e.ldc + col] = x0;
            }
            if (e.S) {
                if (both) *reinterpret_cast<float2*>(e.S + row * e.lds + col) = make_float2(sd[2 * hrow], sd[2 * hrow + 1]);
                else e.S[row * e.lds + col] = sd[2 * hrow];
            }
        }
        if (e.C_h) {
            split_pair16(v[0], both ? v[1] : 0.f, e.c_fmt, ha[i], la[i]);
            split_pair16(v[2], both ? v[3] : 0.f, e.c_fmt, hb[i], lb[i]);
            if (!full) {                               // ragged column group: 4-byte (2-byte at the edge) stores
                uint16_t* Ch = reinterpret_cast<uint16_t*>(e.C_h);
                uint16_t* Cl = reinterpret_cast<uint16_t*>(e.C_l);
                if (va) {
                    const long long o = ra * e.ldh + col;
                    if (both) { *reinterpret_cast<uint32_t*>(Ch + o) = ha[i]; *reinterpret_cast<uint32_t*>(Cl + o) = la[i]; }
                    else { Ch[o] = (uint16_t)ha[i]; Cl[o] = (uint16_t)la[i]; }
                }
                if (vb) {
                    const long long o = rb * e.ldh + col;
                    if (both) { *reinterpret_cast<uint32_t*>(Ch + o) = hb[i]; *reinterpret_cast<uint32_t*>(Cl + o) = lb[i]; }
                    else { Ch[o] = (uint16_t)hb[i]; Cl[o] = (uint16_t)lb[i]; }
                }
            }
        }
    }
    if (e.C_h && full && !e.accumulate) {
        // quad transpose: lane t ends up with the 8 consecutive columns 8 t .. 8 t + 7 of a 32-column group (NB = 4) or,
        // for NB = 2, lanes 0-1 with the two 8-column blocks -> 16-byte stores, full sectors (ldh % 8 == 0, col0 % 16 == 0)
        uint16_t* Ch = reinterpret_cast<uint16_t*>(e.C_h);
        uint16_t* Cl = reinterpret_cast<uint16_t*>(e.C_l);
        if constexpr (NB == 4) {
            quad_transpose(ha, t); quad_transpose(la, t); quad_transpose(hb, t); quad_transpose(lb, t);
            const int col = col0 + 8 * t;
            if (va) {
                const long long o = ra * e.ldh + col;
                *reinterpret_cast<uint4*>(Ch + o) = make_uint4(ha[0], ha[1], ha[2], ha[3]);
                *reinterpret_cast<uint4*>(Cl + o) = make_uint4(la[0], la[1], la[2], la[3]);
            }
            if (vb) {
                const long long o = rb * e.ldh + col;
                *reinterpret_cast<uint4*>(Ch + o) = make_uint4(hb[0], hb[1], hb[2], hb[3]);
                *reinterpret_cast<uint4*>(Cl + o) = make_uint4(lb[0], lb[1], lb[2], lb[3]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int col = col0 + 8 * i + 2 * t;
                if (va) { const long long o = ra * e.ldh + col; *reinterpret_cast<uint32_t*>(Ch + o) = ha[i]; *reinterpret_cast<uint32_t*>(Cl + o) = la[i]; }
                if (vb) { const long long o = rb * e.ldh + col; *reinterpret_cast<uint32_t*>(Ch + o) = hb[i]; *reinterpret_cast<uint32_t*>(Cl + o) = lb[i]; }
            }
        }
    }
}

template <int NB>
__device__ __forceinline__ void epi_frag_p_dispatch(const EpiParamsP& e, const uint32_t* r0, const uint32_t* r1, int lane,
                                                    long long row0, long long m_eff, int col0, int N) {
    switch (e.mode) {
        case IDRK_EPI_SOFTPLUS: epi_frag_p<IDRK_EPI_SOFTPLUS, NB>(e, r0, r1, lane, row0, m_eff, col0, N); break;
        case IDRK_EPI_NONE: epi_frag_p<IDRK_EPI_NONE, NB>(e, r0, r1, lane, row0, m_eff, col0, N); break;
        case IDRK_EPI_RELU: epi_frag_p<IDRK_EPI_RELU, NB>(e, r0, r1, lane, row0, m_eff, col0, N); break;
        default: epi_frag_p<-1, NB>(e, r0, r1, lane, row0, m_eff, col0, N); break;
    }
}

template <bool A_MN, bool B_MN, int BN_>
__global__ void __launch_bounds__(GEMM_THREADS_V2, 1)
gemm_p16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBl,
                long long M, int N, int K, EpiParamsP e, const int* __restrict__ m_count, int kb_per_split, int splits,
                uint32_t fmt_bits) {
    pdl_trigger();
    if (m_count != nullptr) {
        pdl_wait();
        if (*m_count <= 0) return;
    }
    using P = SmemPlanP<BN_>;
    const int n_tiles = (N + BN_ - 1) / BN_;
    const int kb_total = (K + BKH - 1) / BKH;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::STAGES * P::STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P::STAGES + 2 * ACC_STAGES);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (P::STAGES + s); };
    auto acc_full_bar = [&](int a) { return bar_base + 8u * (2 * P::STAGES + a); };
    auto acc_empty_bar = [&](int a) { return bar_base + 8u * (2 * P::STAGES + ACC_STAGES + a); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 32) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmBl); }
    if (threadIdx.x == 0) {
        for (int s = 0; s < P::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(acc_full_bar(a), 1); mbar_init(acc_empty_bar(a), EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr uint32_t TMEM_COLS = ACC_STAGES * 2 * BN_;          // (D0 | D1) per accumulator stage
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    long long m_eff = M;
    if (m_count != nullptr) { const long long c = *m_count; m_eff = c < M ? c : M; }
    const int m_tiles = (int)((m_eff + BM - 1) / BM);
    const int items = m_tiles * n_tiles * splits;

    auto decode = [&](int item, long long& m0, int& n0, int& kb0, int& kb1) {
        const int z = item % splits;
        const int t = item / splits;
        n0 = (t % n_tiles) * BN_;
        m0 = (long long)(t / n_tiles) * BM;
        kb0 = z * kb_per_split;
        kb1 = min(kb_total, kb0 + kb_per_split);
    };

    if (warp == 0) {
        const uint32_t leader = elect_one();
        int it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            long long m0; int n0, kb0, kb1;
            decode(item, m0, n0, kb0, kb1);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % P::STAGES;
                const uint32_t ph = (it / P::STAGES) & 1;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_expect_tx_p(full_bar(s), P::STAGE_BYTES, leader);
                const uint32_t st = smem_base + s * P::STAGE_BYTES;
                const int k0 = kb * BKH;
#pragma unroll
                for (int t = 0; t < 2; ++t) {                     // t = 0: high halves, 1: scaled low halves
                    const uint32_t sA = st + t * P::A_BYTES;
                    const uint32_t sB = st + 2 * P::A_BYTES + t * P::B_BYTES;
                    const CUtensorMap* ta = t ? &tmAl : &tmA;
                    const CUtensorMap* tb = t ? &tmBl : &tmB;
                    if constexpr (!A_MN) tma_load_2d_p(sA, ta, full_bar(s), k0, (int)m0, leader);
                    else {
#pragma unroll
                        for (int c = 0; c < BM / 64; ++c) tma_load_2d_p(sA + c * (BKH * 128), ta, full_bar(s), (int)m0 + 64 * c, k0, leader);
                    }
                    if constexpr (!B_MN) tma_load_2d_p(sB, tb, full_bar(s), k0, n0, leader);
                    else {
#pragma unroll
                        for (int c = 0; c < BN_ / 64; ++c) tma_load_2d_p(sB + c * (BKH * 128), tb, full_bar(s), n0 + 64 * c, k0, leader);
                    }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t leader = elect_one();
        // kind::f16, f32 accumulate; A / B formats (fp16 = 0, bf16 = 1) in bits 7-9 / 10-12 (fmt_bits), majors in bits 15 / 16
        const uint32_t base = (1u << 4) | fmt_bits | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t idesc = base | ((uint32_t)(BN_ >> 3) << 17);
        const uint32_t idesc2 = base | ((uint32_t)((2 * BN_) >> 3) << 17);
        constexpr uint32_t a_lbo = A_MN ? BKH * 128 : 16, b_lbo = B_MN ? BKH * 128 : 16;
        constexpr uint32_t a_step = A_MN ? (2048 >> 4) : (32 >> 4), b_step = B_MN ? (2048 >> 4) : (32 >> 4);
        int it = 0, ti = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++ti) {
            long long m0; int n0, kb0, kb1;
            decode(item, m0, n0, kb0, kb1);
            const int a = ti & 1;
            mbar_wait(acc_empty_bar(a), ((ti >> 1) & 1) ^ 1u);
            tc_fence_after();
            const uint32_t d0 = tmem_base + (uint32_t)(a * 2 * BN_);
            const uint32_t d1 = d0 + BN_;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % P::STAGES;
                const uint32_t ph = (it / P::STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t st = smem_base + s * P::STAGE_BYTES;
                // stage = [A_h | A_l | B_h | B_l]; [B_h ; B_l] is ONE operand of 2 BN rows (K-major: row groups continue at
                // the 1024-byte stride; MN-major: the 64-column chunks continue at the 8 KB stride), [D0 | D1] one accumulator
                const uint64_t a_h = umma_desc(st, a_lbo, 1024, 2);
                const uint64_t a_l = umma_desc(st + P::A_BYTES, a_lbo, 1024, 2);
                const uint64_t b_h = umma_desc(st + 2 * P::A_BYTES, b_lbo, 1024, 2);
#pragma unroll
                for (int k = 0; k < BKH / 16; ++k) {
                    const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
                    tc_mma_f16_p(d0, a_h + k * a_step, b_h + k * b_step, idesc2, acc, leader);      // [D0 | D1] (+)= Ah . [Bh ; Bl]
                    tc_mma_f16_p(d1, a_l + k * a_step, b_h + k * b_step, idesc, 1u, leader);        //  D1      +=  Al . Bh
                }
                tc_commit_p(empty_bar(s), leader);
            }
            tc_commit_p(acc_full_bar(a), leader);
        }
    } else {
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        constexpr int COLS = BN_ / 4;                   // 32 or 16 columns per epilogue warp
        constexpr int NB = COLS / 8;
        int ti = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++ti) {
            long long m0; int n0, kb0, kb1;
            decode(item, m0, n0, kb0, kb1);
            const int a = ti & 1;
            mbar_wait(acc_full_bar(a), (ti >> 1) & 1);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 2 * BN_ + grp * COLS);
            uint32_t a0[4 * NB], a1[4 * NB], b0[4 * NB], b1[4 * NB];      // D0 / D1 fragments of rows +0..15 and +16..31
            if constexpr (NB == 4) {
                tc_ld_16x256b_x4(t0, a0); tc_ld_16x256b_x4(t0 + BN_, a1);
                tc_ld_16x256b_x4(t0 + (16u << 16), b0); tc_ld_16x256b_x4(t0 + (16u << 16) + BN_, b1);
            } else {
                tc_ld_16x256b_x2(t0, a0); tc_ld_16x256b_x2(t0 + BN_, a1);
                tc_ld_16x256b_x2(t0 + (16u << 16), b0); tc_ld_16x256b_x2(t0 + (16u << 16) + BN_, b1);
            }
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty_bar(a));
            const long long row_base = m0 + q * 32;
            const int col0 = n0 + grp * COLS;
            if (col0 < N && row_base < m_eff) {
                epi_frag_p_dispatch<NB>(e, a0, a1, lane, row_base, m_eff, col0, N);
                epi_frag_p_dispatch<NB>(e, b0, b1, lane, row_base + 16, m_eff, col0, N);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// x -> (h, l) pair in `fmt`; pad columns are zero-filled
__global__ void split_p16_kernel(const float* __restrict__ x, long long rows, int cols, int ldx, float scale,
                                 uint16_t* __restrict__ h, uint16_t* __restrict__ l, int ldo, int pad_cols, int fmt,
                                 const int* __restrict__ m_count) {
    pdl_wait();
    pdl_trigger();
    long long r_eff = rows;
    if (m_count) { const long long c = *m_count; r_eff = c < rows ? c : rows; }
    const int w2 = (cols + pad_cols + 1) >> 1;                     // column pairs (ldo is even)
    const long long total = r_eff * (long long)w2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / w2;
        const int c = (int)(i - r * w2) << 1;
        const float x0 = c < cols ? x[r * ldx + c] * scale : 0.f;
        const float x1 = c + 1 < cols ? x[r * ldx + c + 1] * scale : 0.f;
        uint32_t hh, ll;
        split_pair16(x0, x1, fmt, hh, ll);
        if (c + 1 < cols + pad_cols) {
            *reinterpret_cast<uint32_t*>(h + r * ldo + c) = hh;
            *reinterpret_cast<uint32_t*>(l + r * ldo + c) = ll;
        } else { h[r * ldo + c] = (uint16_t)hh; l[r * ldo + c] = (uint16_t)ll; }
    }
}

// 2-D 16-bit tensor map.  K-major operand: dims (K, rows), box (64, box_rows).  MN-major: dims (MN, K), box (64, 64).
static int make_tmap_p16(CUtensorMap* tm, const void* base, int fmt, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box1) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return IDRK_E_DRIVER;
    if (!aligned16(base) || (ld & 7)) return IDRK_E_ALIGN;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : IDRK_E_ARG;
}

template <bool A_MN, bool B_MN, int BN_>
static int launch_p16(const CUtensorMap& tA, const CUtensorMap& tAl, const CUtensorMap& tB, const CUtensorMap& tBl,
                      long long M, int N, int K, const EpiParamsP& e, const int* m_count, int splits, uint32_t fmt_bits, cudaStream_t st) {
    using P = SmemPlanP<BN_>;
    auto kern = gemm_p16_kernel<A_MN, B_MN, BN_>;
    static bool attr_done = false;
    if (!attr_done) {
        IDRK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::TOTAL));
        attr_done = true;
    }
    const int kb_total = (K + BKH - 1) / BKH;
    int kbps = (kb_total + splits - 1) / splits;
    splits = (kb_total + kbps - 1) / kbps;
    const long long items = ((M + BM - 1) / BM) * ((N + BN_ - 1) / BN_) * splits;
    const long long grid = items < sm_count() ? items : sm_count();
    IDRK_CUDA_TRY(launch_k(kern, dim3((unsigned)grid), dim3(GEMM_THREADS_V2), P::TOTAL, st, tA, tAl, tB, tBl, M, N, K, e, m_count,
                           kbps, splits, fmt_bits));
    IDRK_LAUNCH_CHECK();
    return 0;
}

}  // namespace idrk

using namespace idrk;

extern "C" int idrk_gemm_p16(int32_t layout, int64_t M, int32_t N, int32_t K, const void* A_h, const void* A_l, int32_t a_fmt,
                             int32_t lda, const void* B_h, const void* B_l, int32_t b_fmt, int32_t ldb,
                             const idrk_epilogue_p16_t* h_epi, const int32_t* m_count, int32_t split_k, void* stream) {
    if (!A_h || !A_l || !B_h || !B_l || !h_epi || M < 0 || N < 1 || K < 1 || lda < 1 || ldb < 1) return IDRK_E_ARG;
    if (layout < IDRK_GEMM_NT || layout > IDRK_GEMM_TN) return IDRK_E_ARG;
    if ((a_fmt | b_fmt | h_epi->c_fmt) & ~1) return IDRK_E_ARG;
    if (a_fmt != b_fmt) return IDRK_E_UNSUP;            // the hardware rejects mixed fp16 x bf16 kind::f16 operands
    if (M == 0) return 0;
    EpiParamsP e;
    e.C = h_epi->C; e.S = h_epi->S; e.C_h = h_epi->C_h; e.C_l = h_epi->C_l; e.bias = h_epi->bias; e.aux = h_epi->aux;
    e.ldc = h_epi->ldc; e.lds = h_epi->lds; e.ldh = h_epi->ldh; e.ldaux = h_epi->ldaux; e.c_fmt = h_epi->c_fmt;
    e.mode = h_epi->mode; e.act = h_epi->act_param; e.scale = h_epi->scale; e.accumulate = h_epi->accumulate;
    if (split_k < 1) split_k = 1;
    if (split_k > 1) e.accumulate = 1;
    if (e.accumulate && (!e.C || e.C_h || e.S || e.mode != IDRK_EPI_NONE)) return IDRK_E_ARG;
    if (!e.C && !e.C_h) return IDRK_E_ARG;
    if ((e.C_h == nullptr) != (e.C_l == nullptr)) return IDRK_E_ARG;
    if (e.mode == IDRK_EPI_MUL_AUX && !e.aux) return IDRK_E_ARG;
    if ((e.C && (e.ldc < N || (e.ldc & 1))) || (e.S && (e.lds < N || (e.lds & 1)))) return IDRK_E_ARG;
    if (e.C_h && (e.ldh < N || (e.ldh & 7) || !aligned16(e.C_h) || !aligned16(e.C_l))) return IDRK_E_ALIGN;
    if ((e.C && (reinterpret_cast<uintptr_t>(e.C) & 7)) || (e.S && (reinterpret_cast<uintptr_t>(e.S) & 7))) return IDRK_E_ALIGN;
    const bool a_mn = layout == IDRK_GEMM_TN, b_mn = layout != IDRK_GEMM_NT;
    // one wave of 128 x 128 tiles when that fills the SMs, else 128 x 64 (twice the tiles)
    const long long tiles128 = ((M + BM - 1) / BM) * ((N + 127) / 128) * split_k;
    const int bn = (N <= 64 || tiles128 <= sm_count() / 2) ? 64 : 128;
    CUtensorMap tA, tAl, tB, tBl;
    int rc;
    if ((rc = a_mn ? make_tmap_p16(&tA, A_h, a_fmt, (uint64_t)M, (uint64_t)K, lda, BKH) : make_tmap_p16(&tA, A_h, a_fmt, (uint64_t)K, (uint64_t)M, lda, BM))) return rc;
    if ((rc = a_mn ? make_tmap_p16(&tAl, A_l, a_fmt, (uint64_t)M, (uint64_t)K, lda, BKH) : make_tmap_p16(&tAl, A_l, a_fmt, (uint64_t)K, (uint64_t)M, lda, BM))) return rc;
    if ((rc = b_mn ? make_tmap_p16(&tB, B_h, b_fmt, (uint64_t)N, (uint64_t)K, ldb, BKH) : make_tmap_p16(&tB, B_h, b_fmt, (uint64_t)K, (uint64_t)N, ldb, bn))) return rc;
    if ((rc = b_mn ? make_tmap_p16(&tBl, B_l, b_fmt, (uint64_t)N, (uint64_t)K, ldb, BKH) : make_tmap_p16(&tBl, B_l, b_fmt, (uint64_t)K, (uint64_t)N, ldb, bn))) return rc;
    const uint32_t fmt_bits = ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10);
    cudaStream_t st = (cudaStream_t)stream;
#define IDRK_P16(AM, BMN)                                                                                               \
    do {                                                                                                                 \
        if (bn == 64) return launch_p16<AM, BMN, 64>(tA, tAl, tB, tBl, M, N, K, e, m_count, split_k, fmt_bits, st);      \
        return launch_p16<AM, BMN, 128>(tA, tAl, tB, tBl, M, N, K, e, m_count, split_k, fmt_bits, st);                   \
    } while (0)
    if (layout == IDRK_GEMM_NT) IDRK_P16(false, false);
    if (layout == IDRK_GEMM_NN) IDRK_P16(false, true);
    IDRK_P16(true, true);
#undef IDRK_P16
}

extern "C" int idrk_split_p16(const float* x, int64_t rows, int32_t cols, int32_t ldx, float scale, void* h, void* l,
                              int32_t ld_out, int32_t pad_cols, int32_t fmt, const int32_t* m_count, void* stream) {
    if (!x || !h || !l || rows < 0 || cols < 1 || ldx < cols || pad_cols < 0 || ld_out < cols + pad_cols || (ld_out & 1) || (fmt & ~1))
        return IDRK_E_ARG;
    if ((reinterpret_cast<uintptr_t>(h) & 3) || (reinterpret_cast<uintptr_t>(l) & 3)) return IDRK_E_ALIGN;
    if (rows == 0) return 0;
    const long long total = rows * (long long)((cols + pad_cols + 1) / 2);
    long long b = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    IDRK_CUDA_TRY(launch_k(split_p16_kernel, dim3((int)b), dim3(256), 0, (cudaStream_t)stream, x, rows, cols, ldx, scale,
                           (uint16_t*)h, (uint16_t*)l, ld_out, pad_cols, fmt, m_count));
    IDRK_LAUNCH_CHECK();
    return 0;
}
