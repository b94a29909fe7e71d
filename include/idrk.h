/*
 * idrk.h - C ABI of libidrk.so: the B200 (sm_100a) kernels behind the IDR hash-grid
 * rendering hot path of ArtoriasAbyssslayer/HashModNFFBanks-IDR.
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the parameter
 *     name starts with h_ or the struct is documented as host-side;
 *   - the library never allocates or frees caller-visible memory; workspaces are passed in;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it;
 *   - return value: 0 = ok, < 0 = argument error (IDRK_E_*), > 0 = cudaError_t;
 *   - fp32 row-major matrices with an explicit leading dimension (`ld`, in floats).
 *
 * Each function names the reference interface it replaces (file:line under
 * /root/reference/code).  INTEGRATION.md shows the ctypes binding a reference maintainer adds.
 */
#ifndef IDRK_H_
#define IDRK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDRK_MAX_LEVELS 32

#define IDRK_E_ARG      (-1)   /* bad argument (null pointer, size out of range)        */
#define IDRK_E_ALIGN    (-2)   /* pointer / leading dimension not 16-byte aligned         */
#define IDRK_E_UNSUP    (-3)   /* unsupported configuration (e.g. features per level)    */
#define IDRK_E_DRIVER   (-4)   /* CUDA driver entry point (tensor-map encode) unavailable */

/* frac_mode of the hash grid */
#define IDRK_HASH_REFERENCE 0  /* reference semantics: xf == 0, floor-corner row, trunc toward zero
                                  (model/embeddings/hashGridEmbedding.py:84-102)                */
#define IDRK_HASH_TRILINEAR 1  /* floor/frac with 8 weighted corners (documented extension)     */
#define IDRK_HASH_NGP       2  /* instant-ngp / tiny-cuda-nn grid semantics, what tcnn.Encoding("Grid", "Hash", "Linear")
                                  computes for the reference's HashGridTcnn / FFBTcnn selector entries
                                  (model/embeddings/tcnn_src/hashGridEncoderTcnn.py:63-80; tiny-cuda-nn itself is an unpinned,
                                  un-vendored dependency, so this follows its published algorithm): inputs in [0, 1],
                                  pos = fma(x, scale_l, 0.5) with scale_l = res[l] = base * per_level_scale^l - 1,
                                  grid resolution R_l = ceil(scale_l) + 1, 8 corners weighted by pos - floor(pos);
                                  table row = x + y R + z R^2 when R^3 <= rows[l] (dense level), else
                                  (x * 1) ^ (y * 2654435761) ^ (z * 805459861); both taken mod rows[l].  F = 2 only. */

/* Host-side description of one MultiResHashGridMLP (model/embeddings/hashGridEmbedding.py:105-155).
 * Output row layout: [x(3) | sin(C) | cos(C) | level_0(F) ... level_{L-1}(F)], C = n_fourier.
 * With n_fourier == 0 the row is just the L*F level features. */
typedef struct idrk_hashgrid {
    int32_t n_levels;                       /* L                                              */
    int32_t n_feat;                         /* F in {1,2,4,8}                                  */
    int32_t frac_mode;                      /* IDRK_HASH_*                                     */
    int32_t n_fourier;                      /* C: channels of the FourierFeature prefix        */
    float    res[IDRK_MAX_LEVELS];          /* grid resolution per level (float(res_l))        */
    uint32_t rows[IDRK_MAX_LEVELS];         /* table rows per level T_l                        */
    const float* tables[IDRK_MAX_LEVELS];   /* device pointer to [T_l, F] per level            */
    const float* fourier_B;                 /* device [3, C] (freq_encoding.B) or NULL         */
} idrk_hashgrid_t;

/* -- version / capability ------------------------------------------------------------- */
int idrk_version(void);                        /* ABI version, currently 4 (3 -> 4: idrk_posenc_dx_bwd, idrk_rt_iter_tail, idrk_nffb_encode_f16pair, idrk_gemm_p16, idrk_split_p16, idrk_weight_norm_fwd_p16, idrk_act_bwd_p16; 2 -> 3: idrk_hash_encode_fwd gained `perm`, idrk_hash_encode_bwd `flags` and `perm`; idrk_morton_sort; idrk_hash_encode_bwd_det; idrk_sdf_squash_rows; idrk_sumsq_det; idrk_rt_linesearch_points / _resolve; idrk_camera_rays, idrk_idr_loss, idrk_scale3, idrk_fourier_dx_fwd / _bwd; 1 -> 2: idrk_epilogue_f16_t grew dot_w / dot_out / ld_dot; IDRK_HASH_NGP; idrk_hash_encode_f16pair) */
int idrk_device_sm_count(int* out_sms);        /* SM count of the current device */

/* -- K1: hash-grid encode forward -------------------------------------------------------
 * Replaces MultiResHashGridMLP.forward (hashGridEmbedding.py:150-155) =
 * FourierFeature.forward (frequency_enc.py:63-67) ++ _HashGridMLP.forward x L (:81-102)
 * ++ hash_func (:32-40).   x [n, ldx>=3], out [n, ld_out], ld_out >= width; columns
 * width..ld_out-1 are written as zeros.  idx_debug (nullable) receives the uint32 table row
 * of all 8 corners, [n, L, 8].  m_count (nullable, device int32): only the first min(n, *m_count)
 * points are encoded (device-side compaction in the ray tracer).  perm (nullable, device int32 [n], a permutation
 * of 0..n-1, e.g. from idrk_morton_sort): points are PROCESSED in the order perm[0], perm[1], ... - x and out keep the
 * caller's row order, only the walk changes - so an unordered batch gets the table-row locality of a Z-ordered one
 * without being copied.  Tables beyond L2 (> 512 MB, 8-corner modes) are walked in level windows (several launches). */
int idrk_hash_encode_fwd(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx,
                         float* out, int32_t ld_out, uint32_t* idx_debug, const int32_t* m_count,
                         const int32_t* perm, void* stream);

/* -- K2d: deterministic table gradients ----------------------------------------------------------------------
 * Same result as the table-gradient part of idrk_hash_encode_bwd up to fp32 summation order, but with that order FIXED:
 * every (point, level, corner) contribution is keyed by (level, table row), a stable radix sort groups a row's
 * contributions in ascending point order and one thread sums each row's segment - bit-identical gradients run to run
 * (the reference's dense embedding backward is deterministic on CPU; the atomic scatter is not).  ACCUMULATES into
 * h_grad_tables like idrk_hash_encode_bwd; dL/dx comes from idrk_hash_encode_bwd with h_grad_tables = NULL (already
 * order-independent).  workspace: >= idrk_hash_encode_bwd_det_workspace(n) bytes of device scratch, 16-byte aligned;
 * n * L * (8 | 1) contributions must stay below 2^31 and table rows below 2^(32 - ceil(log2 L)). */
int idrk_hash_encode_bwd_det_workspace(const idrk_hashgrid_t* h_grid, int64_t n, int64_t* out_bytes);
int idrk_hash_encode_bwd_det(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx, const float* dy,
                             int32_t ld_dy, float* const* h_grad_tables, void* workspace, int64_t workspace_bytes,
                             void* stream);

/* -- K1p: hash-grid encode straight into the fp16-pair operand of idrk_gemm_f16s ----------------
 * Same values as idrk_hash_encode_fwd (bit-identical fp32 columns) stored as h = fp16(v), l = fp16((v - h) * 2^11)
 * into (h, l) [n, ld_out] halves, columns width .. width + pad_cols - 1 zero-filled; optionally scale2 * v into a
 * second pair (h2, l2) [n, ld_out2] (the skip connection's copy of the embedding).  Replaces
 * MultiResHashGridMLP.forward inside the ray tracer's SDF queries (ray_tracing.py: sdf(points)), F = 2. */
int idrk_hash_encode_f16pair(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx,
                             const int32_t* m_count, void* h, void* l, int32_t ld_out, int32_t pad_cols,
                             void* h2, void* l2, int32_t ld_out2, int32_t pad_cols2, float scale2, void* stream);

/* -- K2: hash-grid encode backward ------------------------------------------------------
 * Replaces autograd through the same functions: embedding_dense_backward scatter-add into
 * every level's table gradient and d/dx of the Fourier prefix.  dy [n, ld_dy] is dL/d(out).
 * h_grad_tables: HOST array of L device pointers, each [T_l, F] (NULL = skip table gradients); ACCUMULATED
 * (caller zero-fills).  dx (nullable) [n, 3] is overwritten.
 * flags: IDRK_HASH_BWD_ORDERED = the caller's hint that consecutive points are spatially close (samples along rays,
 * Z-ordered batches): the 8-corner table-gradient pass then merges runs of points that share a cell of a level in
 * registers before issuing reductions (same result up to fp32 summation order; harmless but slower on unordered input).
 * perm (nullable): as in idrk_hash_encode_fwd (rows of x, dy and dx are addressed through it) and implies the hint.
 * Table-gradient passes over tables larger than ~80 MB run as level windows so each window's gradient rows stay in L2. */
#define IDRK_HASH_BWD_ORDERED 1
int idrk_hash_encode_bwd(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx,
                         const float* dy, int32_t ld_dy, float* const* h_grad_tables,
                         float* dx, int32_t flags, const int32_t* perm, void* stream);

/* -- positional encoding ----------------------------------------------------------------
 * Replaces PositionalEncoding.embed (frequency_enc.py:19-51) and get_embedder (:156-168).
 * x [n, d] -> out [n, ld_out]; row = [x | x | sin(f_0 x) | cos(f_0 x) | ...] when include_input
 * (the input appears twice, as in the reference), h_bands = HOST array of n_bands frequencies.
 * Columns beyond the width are zero-filled.  bwd writes dx [n, d]. */
int idrk_posenc_fwd(const float* x, int64_t n, int32_t d, int32_t ldx, const float* h_bands, int32_t n_bands,
                    int32_t include_input, float* out, int32_t ld_out, void* stream);
int idrk_posenc_bwd(const float* x, int64_t n, int32_t d, int32_t ldx, const float* h_bands, int32_t n_bands,
                    int32_t include_input, const float* dy, int32_t ld_dy, float* dx, int32_t ld_dx, void* stream);
/* Backward of idrk_posenc_bwd's map (x, dy) -> dx, for the RECORDED backward of the encoding (ImplicitNetwork.gradient with
 * create_graph=True through the filter banks' positional encodings, nffb3d.py:170-173): given g = d L / d dx it writes
 * g_dy [n, ld_gdy] (all `width` columns, pads zero) and / or g_x [n, d] (either may be NULL). */
int idrk_posenc_dx_bwd(const float* g, int32_t ld_g, const float* x, int64_t n, int32_t d, int32_t ldx, const float* h_bands,
                       int32_t n_bands, int32_t include_input, const float* dy, int32_t ld_dy, float* g_dy, int32_t ld_gdy,
                       float* g_x, int32_t ld_gx, void* stream);

/* -- K4: MLP contraction tiles (tcgen05 / TMEM / TMA) --------------------------------------
 * Replaces the nn.Linear + activation pairs of ImplicitNetwork.forward
 * (implicit_differentiable_renderer.py:96-112), RenderingNetwork.forward (:215-223), the SIREN layers
 * of FourierFilterBanks.forward (nffb3d.py:163-184) and their autograd (data / weight gradients).
 *   layout NT: A[M,K] B[N,K]  C = A B^T   (forward: activations x weights)
 *   layout NN: A[M,K] B[K,N]  C = A B     (gradient w.r.t. the layer input)
 *   layout TN: A[K,M] B[K,N]  C = A^T B   (gradient w.r.t. the weights; contraction over points)
 * precision: FP32 = plain FFMA tiles; TF32 = one tensor-core pass; 3XTF32 = hi/lo split operands
 * (A_lo / B_lo hold  tf32(x - tf32(x)),  A / B hold tf32(x)), fp32-accurate.
 * Tensor-core modes need 16-byte aligned operands with lda, ldb % 4 == 0.
 * m_count (nullable, device): only rows < min(M, *m_count) are computed (device-side compaction).
 * split_k > 1 accumulates partial tiles with atomics into C (caller zero-fills, epilogue must be NONE). */
#define IDRK_GEMM_NT 0
#define IDRK_GEMM_NN 1
#define IDRK_GEMM_TN 2
#define IDRK_PREC_FP32   0
#define IDRK_PREC_TF32   1
#define IDRK_PREC_3XTF32 3
#define IDRK_EPI_NONE     0   /* v = acc + bias                                              */
#define IDRK_EPI_SOFTPLUS 1   /* v = softplus(acc + bias; beta = act_param, threshold 20), S = sigmoid */
#define IDRK_EPI_RELU     2   /* v = max(acc + bias, 0), S = step                             */
#define IDRK_EPI_MUL_AUX  3   /* v = (acc + bias) * aux[row, col]                             */
#define IDRK_EPI_SINE     4   /* v = sin(act_param * (acc + bias)), S = act_param * cos(...)  */
#define IDRK_EPI_TANH     5   /* v = tanh(acc + bias), S = 1 - v^2                            */

typedef struct idrk_epilogue {
    float* C;            /* [M, ldc] fp32 result (nullable when C_hi/C_lo are given)            */
    float* C_hi;         /* [M, ldc] tf32(v)          } pre-split operands for a following      */
    float* C_lo;         /* [M, ldc] tf32(v - tf32(v))} 3xTF32 contraction (both or neither)    */
    float* S;            /* [M, lds] activation derivative (nullable)                           */
    const float* bias;   /* [N] (nullable)                                                      */
    const float* aux;    /* [M, ldaux] multiplier for IDRK_EPI_MUL_AUX                          */
    int32_t ldc, lds, ldaux;
    int32_t mode;        /* IDRK_EPI_*                                                          */
    float act_param;     /* softplus beta / sine w0                                             */
    float scale;         /* v *= scale after the activation (1/sqrt(2) of the skip connection)  */
    int32_t accumulate;  /* 1: C += v with atomics                                              */
} idrk_epilogue_t;

int idrk_gemm(int32_t layout, int32_t precision, int64_t M, int32_t N, int32_t K,
              const float* A, const float* A_lo, int32_t lda, const float* B, const float* B_lo, int32_t ldb,
              const idrk_epilogue_t* h_epi, const int32_t* m_count, int32_t split_k, void* stream);

/* -- K7: fused Fourier-filter-bank encoder, forward only (no-grad SDF queries of the ray tracer) ----------------
 * FourierFilterBanks.forward of the configuration the reference's selector uses (PositionalEncodingNET + SIREN,
 * has_out = False; nffb3d.py:122-194) with the optional StyleAttention block (styleMod.py:17-44) in ONE launch:
 * out[p] = [u | f / n_levels_div], u = (x + bound) / (2 bound).  Weights are the module's own tensors
 * (ff_lin{j}.weight [width, in] row-major, out_layer, StyleAttentionBlock.linear_transform).  width <= 64. */
typedef struct idrk_nffb {
    idrk_hashgrid_t grid;          /* grid_enc (reference frac mode), evaluated at u                  */
    float bands[32];               /* PositionalEncoding freq_bands                                   */
    int32_t n_bands, include_input;
    int32_t n_lin;                 /* SIREN layers ff_lin0 .. ff_lin{n_lin-1}                          */
    int32_t width, chunk;          /* filter-bank width; grid columns per chunk (2 * features/level)  */
    int32_t style, n_levels_div;   /* style modulation on/off; divisor of the accumulated features    */
    float bound, w0, eps;
    const float* lin_w[16];
    const float* lin_b[16];
    const float* out_w; const float* out_b;
    const float* style_w; const float* style_b;
} idrk_nffb_t;
int idrk_nffb_encode_fwd(const idrk_nffb_t* desc, const float* x, int64_t n, int32_t ldx, float* out, int32_t ld_out,
                         const int32_t* m_count, void* stream);
/* The same encoder inside the ray tracer's SDF queries, written directly as the fp16-pair operand of idrk_gemm_f16s
 * (h = fp16(v), l = fp16((v - h) 2^11); columns [3 + width, 3 + width + pad_cols) zero-filled) and, with out_h2 != NULL,
 * a second pair of scale2 * v (the skip connection's copy of the embedding). */
int idrk_nffb_encode_f16pair(const idrk_nffb_t* desc, const float* x, int64_t n, int32_t ldx, const int32_t* m_count,
                             void* out_h, void* out_l, int32_t ld_out, int32_t pad_cols,
                             void* out_h2, void* out_l2, int32_t ld_out2, int32_t pad_cols2, float scale2, void* stream);

/* -- K4c: fp16-pair contraction for the no-grad SDF path ------------------------------------
 * Operands are pairs of IEEE halves  x ~= h + l * 2^-11  (h = fp16(x), l = fp16((x - h) * 2^11)): 4 bytes per element,
 * three kind::f16 MMAs per product (Ah.Bh | Ah.Bl + Al.Bh) into two TMEM accumulators, result = D0 + 2^-11 D1.
 * Same use as idrk_gemm NT (ImplicitNetwork.forward under no_grad, ray_tracing.py SDF queries); needs |x| < 65504.
 * A [M, K], B [N, K] row-major halves with lda, ldb % 8 == 0.  Outputs: fp32 C and/or the half pair (C_h, C_l). */
typedef struct idrk_epilogue_f16 {
    float* C;            /* [M, ldc] fp32 (nullable)                      */
    void* C_h;           /* [M, ldh] fp16 high halves (nullable, with C_l) */
    void* C_l;           /* [M, ldh] fp16 scaled low halves                */
    const float* bias;   /* [N] (nullable)                                 */
    int32_t ldc, ldh;
    int32_t mode;        /* IDRK_EPI_NONE or IDRK_EPI_SOFTPLUS             */
    float act_param, scale;
    /* Fused row-dot of the activated tile (the SDF head: ImplicitNetwork's last Linear, row 0,
     * implicit_differentiable_renderer.py:96-112 with only column 0 consumed, ray_tracing.py SDF queries).
     * With dot_w != NULL the [M, N] activation is NOT stored (C, C_h, C_l must be NULL); instead
     *   dot_out[r * ld_dot + j] = sum over columns c in [32 j, 32 j + 32) of act(z[r, c]) * dot_w[c],
     * one fp32 partial per 32-column group (ld_dot >= ceil(N / 32)), summed in a fixed order: deterministic.
     * idrk_sdf_head over the partials (K = ceil(N / 32), w = ones) finishes: + bias, Laplace squash. */
    const float* dot_w;  /* [N] (nullable)                                 */
    float* dot_out;      /* [M, ld_dot]                                    */
    int32_t ld_dot;
} idrk_epilogue_f16_t;
int idrk_gemm_f16s(int64_t M, int32_t N, int32_t K, const void* A_h, const void* A_l, int32_t lda,
                   const void* B_h, const void* B_l, int32_t ldb, const idrk_epilogue_f16_t* h_epi,
                   const int32_t* m_count, void* stream);
/* idrk_split_f16: fp32 -> fp16 pair; (h2, l2) optional second destination receiving scale2 * x in the same pass. */
int idrk_split_f16(const float* x, int64_t rows, int32_t cols, int32_t ldx, float scale, void* h, void* l,
                   int32_t ld_out, int32_t pad_cols, void* h2, void* l2, int32_t ld_out2, int32_t pad_cols2,
                   float scale2, const int32_t* m_count, void* stream);

/* -- K4p: 16-bit-pair contraction for the DIFFERENTIABLE MLP path (csrc/gemm_p16.cu) ------------------------------
 * Same operation set as idrk_gemm (layouts NT / NN / TN, bias, activation + derivative outputs, split-K) - i.e. the
 * Linear / Softplus / ReLU / tanh / sine layers of ImplicitNetwork.forward / .gradient and RenderingNetwork.forward and
 * their first- and second-order autograd (implicit_differentiable_renderer.py:96-128,215-223) - on operands stored as
 * pairs of 16-bit floats  x ~= h + l * 2^-11  instead of the 8-byte 3xTF32 pair:
 *   fmt 0 (IDRK_P16_FP16): h = fp16(x), l = fp16((x - h) 2^11): ~22 bits, |x| < 65504 (weights, activations);
 *   fmt 1 (IDRK_P16_BF16): h = bf16(x), l = bf16((x - h) 2^11): ~17 bits, full fp32 range (cotangents).
 * A and B of one call must share the format (mixed fp16 x bf16 instructions fault on sm_100a: IDRK_E_UNSUP); the output pair
 * may use either.  lda, ldb, ldh % 8 == 0 and 16-byte aligned bases (TMA).  Outputs: fp32 C and / or the
 * pair (C_h, C_l) in c_fmt, optionally S = act'(z); accumulate / split_k > 1: atomic adds into a zero-initialised fp32 C. */
#define IDRK_P16_FP16 0
#define IDRK_P16_BF16 1
typedef struct idrk_epilogue_p16 {
    float* C;            /* [M, ldc] fp32 (nullable when the pair is written)   */
    float* S;            /* [M, lds] activation derivative (nullable)           */
    void* C_h;           /* [M, ldh] high halves (nullable, with C_l)           */
    void* C_l;           /* [M, ldh] scaled low halves                          */
    const float* bias;   /* [N] (nullable)                                      */
    const float* aux;    /* [M, ldaux] for IDRK_EPI_MUL_AUX                     */
    int32_t ldc, lds, ldh, ldaux;
    int32_t c_fmt;       /* IDRK_P16_* of (C_h, C_l)                            */
    int32_t mode;        /* IDRK_EPI_*                                          */
    float act_param, scale;
    int32_t accumulate;
} idrk_epilogue_p16_t;
int idrk_gemm_p16(int32_t layout, int64_t M, int32_t N, int32_t K, const void* A_h, const void* A_l, int32_t a_fmt,
                  int32_t lda, const void* B_h, const void* B_l, int32_t b_fmt, int32_t ldb,
                  const idrk_epilogue_p16_t* epi, const int32_t* m_count, int32_t split_k, void* stream);
/* fp32 -> 16-bit pair of scale * x; columns [cols, cols + pad_cols) zero-filled; ld_out even. */
int idrk_split_p16(const float* x, int64_t rows, int32_t cols, int32_t ldx, float scale, void* h, void* l,
                   int32_t ld_out, int32_t pad_cols, int32_t fmt, const int32_t* m_count, void* stream);
/* idrk_weight_norm_fwd / idrk_act_bwd (below) with the operand written as a 16-bit pair instead of the tf32 hi / lo pair */
int idrk_weight_norm_fwd_p16(const float* g, const float* v, int32_t N, int32_t K, int32_t ldv, float* W, int32_t ldw,
                             void* W_h, void* W_l, int32_t ldp, int32_t fmt, void* stream);
int idrk_act_bwd_p16(const float* dH, int32_t ld_dh, const float* dS, int32_t ld_ds, const float* S, int32_t ld_s,
                     const float* H, int32_t ld_h, int64_t rows, int32_t cols, int32_t mode, float act, float scale,
                     float* dZ, int32_t ld_out, void* dZ_h, void* dZ_l, int32_t ld_pair, int32_t fmt, void* stream);

/* -- helpers around the MLP tiles ----------------------------------------------------------
 * idrk_split_tf32: v = scale * x; hi = tf32(v), lo = tf32(v - hi) for 3xTF32 operands (lo nullable -> plain
 *   scaled copy rounded to tf32 is NOT applied: hi then receives v itself); columns [cols, cols+pad_cols)
 *   of the outputs are zero-filled.
 * idrk_weight_norm_fwd/bwd: legacy nn.utils.weight_norm(dim=0) used by every lin{l}
 *   (implicit_differentiable_renderer.py:80-81,195-196): W = g * v / ||v||_row.  g == NULL copies v.
 *   Optional outputs: W, its hi/lo split, and the transposed copies Wt[K, ldwt] (K-major operand of
 *   the input-gradient contraction).
 * idrk_colsum: out[c] += sum_r x[r, c]  (bias gradients).
 * idrk_sdf_head: sdf = tanh(s / (2 + rho(s))), s = <h[p, :K], w> + bias[0]  - the only output column the
 *   ray tracer consumes (implicit_differentiable_renderer.py:112, :257; density_net.py:20-30).
 * idrk_sdf_squash: the same squash on precomputed s (+ optional derivative d out / d s). */
int idrk_split_tf32(const float* x, int64_t rows, int32_t cols, int32_t ldx, float scale, float* hi, float* lo,
                    int32_t ld_out, int32_t pad_cols, const int32_t* m_count, void* stream);
int idrk_weight_norm_fwd(const float* g, const float* v, int32_t N, int32_t K, int32_t ldv,
                         float* W, float* W_hi, float* W_lo, int32_t ldw,
                         float* Wt, float* Wt_hi, float* Wt_lo, int32_t ldwt, void* stream);
int idrk_weight_norm_bwd(const float* g, const float* v, const float* dW, int32_t N, int32_t K, int32_t ldv,
                         int32_t lddw, float* dg, float* dv, int32_t lddv, int32_t accumulate, void* stream);
int idrk_colsum(const float* x, int64_t rows, int32_t cols, int32_t ldx, float* out, void* stream);
int idrk_sdf_head(const float* h, int64_t rows, int32_t K, int32_t ldh, const float* w, const float* bias,
                  float beta, float* out, const int32_t* m_count, void* stream);
int idrk_sdf_squash(const float* s, int64_t n, float beta, float* out, float* dout, void* stream);
/* idrk_sdf_squash_rows: out [rows, ld_out] = x [rows, ldx] with column 0 squashed (the autograd path's last step,
 * implicit_differentiable_renderer.py:108-113), d[r] = d out_0 / d s and d2[r] = d^2 out_0 / d s^2 with rho held constant
 * (nullable); columns cols..ld_out-1 are zero-filled. */
int idrk_sdf_squash_rows(const float* x, int64_t rows, int32_t cols, int32_t ldx, float beta, float* out, int32_t ld_out,
                         float* d, float* d2, void* stream);
/* idrk_act_bwd: backward of the fused activation epilogue H = scale * act(Z), S = act'(Z) (autograd of the
 * Softplus / ReLU / Sine / Tanh layers):  dZ = dH * S * scale + dS * act''(Z)  (dH or dS may be NULL); dZ is also
 * written as a 3xTF32 operand pair when dZ_hi / dZ_lo are given.  Columns [cols, ld_out) are zero-filled. */
int idrk_act_bwd(const float* dH, int32_t ld_dh, const float* dS, int32_t ld_ds, const float* S, int32_t ld_s,
                 const float* H, int32_t ld_h, int64_t rows, int32_t cols, int32_t mode, float act, float scale,
                 float* dZ, float* dZ_hi, float* dZ_lo, int32_t ld_out, void* stream);

/* -- K5: ray-state kernels of RayTracing ----------------------------------------------------
 * Replace model/ray_tracing.py (forward :26-95, sphere_tracing :98-187, ray_sampler :189-249,
 * secant :251-268, minimal_sdf_points :270-298).  The caller owns all per-ray arrays (host struct of
 * device pointers below) and evaluates the SDF on the compacted point lists between calls.
 * `counter` arguments are device int32 append counters (caller zeroes them), `gate` arguments are
 * device int32 flags: the kernel is a no-op when *gate == 0 (the reference's `break`). */
typedef struct idrk_ray_state {
    const float* cam_loc;    /* [B, 3]                                                    */
    const float* ray_dirs;   /* [N, 3], N = B * num_pixels                                */
    int32_t n_rays, num_pixels;
    float *t0, *t1;          /* acc_start_dis, acc_end_dis                                */
    float *cur_s, *cur_e, *nxt_s, *nxt_e;   /* curr / next sdf at both ends               */
    float *ps, *pe;          /* curr_start_points, curr_end_points [N, 3]                 */
    float *min_dis, *max_dis;
    uint8_t *unf_s, *unf_e;  /* unfinished masks                                          */
    int32_t *slot_s, *slot_e;/* list slot of the point submitted for evaluation, -1 none  */
} idrk_ray_state_t;

int idrk_rt_init(const idrk_ray_state_t* h_state, const float* t_sph, const uint8_t* hit, float* pts, int32_t* counter,
                 void* stream);
/* gather_mode: 0 none, 1 nxt = slot >= 0 ? vals[slot] : 0, 2 nxt = vals[slot] where slot >= 0 */
int idrk_rt_top(const idrk_ray_state_t* h_state, const float* vals, int32_t gather_mode, float sdf_threshold,
                int32_t* n_unfinished, void* stream);
int idrk_rt_step(const idrk_ray_state_t* h_state, const int32_t* gate, float* pts, int32_t* counter, void* stream);
int idrk_rt_linesearch(const idrk_ray_state_t* h_state, const int32_t* gate, const float* vals, int32_t gather_mode,
                       float factor, float* pts, int32_t* counter, void* stream);
/* The back-off line search of one sphere-tracing iteration as ONE evaluation (ray_tracing.py:167-183): _points gathers
 * the step's values, lists the n_ls candidate positions t_k = t_{k-1} -+ h_factors[k] * cur_sdf of every ray end whose new
 * sdf is negative (n_ls consecutive slots, counted in *counter; pts needs room for 2 * n_ls * n_rays points), _resolve
 * adopts per ray end the first candidate whose value is not negative (else the last) - the state the reference's
 * sequential re-evaluation leaves.  h_factors: HOST float[n_ls] = (1 - line_search_step) / 2^k; n_ls <= 8. */
int idrk_rt_linesearch_points(const idrk_ray_state_t* h_state, const int32_t* gate, const float* vals, const float* h_factors,
                              int32_t n_ls, float* pts, int32_t* counter, void* stream);
int idrk_rt_linesearch_resolve(const idrk_ray_state_t* h_state, const int32_t* gate, const float* vals, const float* h_factors,
                               int32_t n_ls, void* stream);
int idrk_rt_end(const idrk_ray_state_t* h_state, const int32_t* gate, const float* vals, int32_t gather_mode, void* stream);
/* idrk_rt_linesearch_resolve + idrk_rt_end + the next iteration's idrk_rt_top (gather_mode 0) in ONE launch: the three are
 * per-ray passes over the same state (ray_tracing.py:167-186 and the loop top :131-142 of the following iteration);
 * *n_unfinished (zero-initialised) receives the next gate. */
int idrk_rt_iter_tail(const idrk_ray_state_t* h_state, const int32_t* gate, const float* vals, const float* h_factors,
                      int32_t n_ls, float sdf_threshold, int32_t* n_unfinished, void* stream);
int idrk_rt_select_sampler(const idrk_ray_state_t* h_state, uint8_t* net_mask, int32_t* ray_of_slot, int32_t* counter,
                           void* stream);
/* n_dev (nullable, device int32): when given, the number of valid slots is min(host bound, *n_dev) so the caller
 * needs no host read of the list length; idrk_rt_chunk_counts writes, per chunk c of `per_chunk` slots,
 * out[c] = mult * clamp(*n_dev - c * per_chunk, 0, per_chunk)  (row limits for the SDF kernels). */
int idrk_rt_chunk_counts(const int32_t* n_dev, int32_t per_chunk, int32_t mult, int32_t n_chunks, int32_t* out, void* stream);
int idrk_rt_sampler_points(const idrk_ray_state_t* h_state, const int32_t* ray_of_slot, int32_t slot0, int32_t n_slots,
                           int32_t n_steps, const float* lin, float* pts, const int32_t* n_dev, void* stream);
int idrk_rt_sampler_resolve(const idrk_ray_state_t* h_state, const int32_t* ray_of_slot, int32_t n_slots, int32_t n_steps,
                            const float* lin, const float* vals, const uint8_t* object_mask, int32_t training,
                            uint8_t* net_mask, float* z_lo, float* z_hi, float* s_lo, float* s_hi,
                            int32_t* sec_slots, int32_t* sec_counter, const int32_t* n_dev, void* stream);
/* mode 0: emit first prediction, 1: consume sdf + emit next, 2: consume + write result, 3: write initial prediction */
int idrk_rt_secant(const idrk_ray_state_t* h_state, const int32_t* ray_of_slot, const int32_t* sec_slots, int32_t n_sec,
                   int32_t mode, const float* vals, float* z_lo, float* z_hi, float* s_lo, float* s_hi, float* pts,
                   const int32_t* n_dev, void* stream);
int idrk_rt_select_minsdf(const idrk_ray_state_t* h_state, const uint8_t* net_mask, const uint8_t* object_mask,
                          const uint8_t* hit, const uint8_t* sampler_mask, int32_t* ray_of_slot, int32_t* counter,
                          void* stream);
int idrk_rt_minsdf_points(const idrk_ray_state_t* h_state, const int32_t* ray_of_slot, int32_t slot0, int32_t n_slots,
                          int32_t n_steps, const float* u, float* pts, const int32_t* n_dev, void* stream);
int idrk_rt_minsdf_resolve(const idrk_ray_state_t* h_state, const int32_t* ray_of_slot, int32_t n_slots, int32_t n_steps,
                           const float* u, const float* vals, const int32_t* n_dev, void* stream);

/* -- K6: optimiser step on the flat bucket ------------------------------------------------
 * Replaces clip_grad_norm_(params, max_norm) + torch.optim.Adam.step() of idr_train.py:306-308 for
 * the data-parallel trainer.  idrk_sumsq: *out += sum(g^2) (caller zeroes out).  idrk_clip_adam: g is
 * first scaled by grad_scale (1/world_size after the all-reduce), clipped by
 * min(1, max_norm / (grad_scale * sqrt(*sumsq) + 1e-6)) when max_norm > 0, then Adam (no weight decay,
 * bias correction with `step` >= 1) updates p, m, v in place. */
int idrk_sumsq(const float* g, int64_t n, float* out, void* stream);
/* idrk_sumsq_det: *out = sum(g^2) (overwritten) with a summation order that depends only on (n, n_partials): CTA b leaves
 * its partial in partials[b] (device scratch, n_partials floats), one block adds them in a fixed order.  Data-parallel
 * replicas must derive the SAME clip factor from the all-reduced bucket; the atomic form above differs in the last bits
 * from run to run and lets replicas drift apart. */
int idrk_sumsq_det(const float* g, int64_t n, float* out, float* partials, int32_t n_partials, void* stream);
int idrk_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, int32_t step, float max_norm, const float* sumsq, float grad_scale, void* stream);

/* -- K8: O(rays) ends of the step (csrc/render_glue.cu) --------------------------------------
 * idrk_camera_rays replaces get_camera_params + lift (utils/rend_util.py:48-75, :87-100: pixel -> world ray through a
 * 4x4 camera-to-world pose and a 4x4 intrinsics matrix with skew, normalised) and, when t_sph / hit are given,
 * get_sphere_intersection (utils/rend_util.py:141-162: under = (d.c)^2 - (|c|^2 - r^2), hit = under > 0,
 * t = -+sqrt(under) - d.c clamped at 0, zeros for rays that miss) - one launch, no boolean-mask indexing.
 *   uv [B, N, 2], pose [B, 4, 4], intrinsics [B, 4, 4] -> ray_dirs [B, N, 3], cam_loc [B, 3], t_sph [B, N, 2], hit [B, N] (bytes).
 * idrk_idr_loss replaces IDRLoss.forward (model/loss.py:5-71) and its backward with respect to the three network
 * outputs: out_losses[4] = {loss, rgb_loss, eikonal_loss, mask_loss};
 *   rgb_loss  = sum_{net & obj} |rgb - gt| / N,  eikonal = mean_M (|g| - 1)^2,
 *   mask_loss = (1 / alpha) * sum_{!(net & obj)} BCEwithlogits(-alpha * sdf, obj) / N,
 *   loss = rgb_loss + eikonal_weight * eikonal + mask_weight * mask_loss;
 * d_rgb [N, 3], d_sdf [N], d_grad_theta [M, 3] (any may be NULL) receive d loss / d input.  Masks are bytes (torch.bool
 * storage).  One thread block, fixed reduction order: deterministic.  idrk_scale3: y = (*scale) * x for three buffers
 * (the chain-rule factor of the incoming loss gradient). */
int idrk_camera_rays(const float* uv, const float* pose, const float* intrinsics, int32_t n_images, int32_t n_pixels,
                     float radius, float* ray_dirs, float* cam_loc, float* t_sph, uint8_t* hit, void* stream);
int idrk_idr_loss(const float* rgb_values, int32_t ld_rgb, const float* rgb_gt, const uint8_t* network_object_mask,
                  const uint8_t* object_mask, const float* sdf_output, int32_t ld_sdf, int64_t n_rays,
                  const float* grad_theta, int32_t ld_grad, int64_t n_grad, float eikonal_weight, float mask_weight,
                  float alpha, float* out_losses, float* d_rgb, float* d_sdf, float* d_grad_theta, void* stream);
/* Differentiable d/dx of the FourierFeature prefix [x | sin(2 pi x B) | cos(2 pi x B)] (model/embeddings/frequency_enc.py:63-67)
 * for the RECORDED backward pass of ImplicitNetwork.gradient (create_graph=True, implicit_differentiable_renderer.py:116-128):
 * fwd: dx = dy_x + 2 pi sum_j (dy_sin_j cos xp_j - dy_cos_j sin xp_j) B[:, j];  bwd: given g = d L / d dx, the gradient with
 * respect to dy (full `width` columns, zeros beyond the prefix) and, when g_x != NULL, with respect to x. */
int idrk_fourier_dx_fwd(const float* x, int32_t ldx, const float* dy, int32_t ld_dy, const float* B, int32_t n_fourier,
                        int64_t n, float* dx, void* stream);
int idrk_fourier_dx_bwd(const float* g, int32_t ld_g, const float* x, int32_t ldx, const float* dy, int32_t ld_dy,
                        const float* B, int32_t n_fourier, int64_t n, float* g_dy, int32_t ld_gdy, int32_t width,
                        float* g_x, void* stream);
/* -- Z-order permutation of a point batch (csrc/point_sort.cu) -------------------------------------------------
 * perm[i] = index of the i-th point along the Morton curve of a (2^bits_per_dim)^3 lattice over the box [h_lo, h_hi]
 * (HOST float[3] each; points outside are clamped; both NULL = the batch's own bounding box, measured on the device
 * without a host read), bits_per_dim in 1..10.  Stable LSD radix sort of (key, index) pairs,
 * 8 bits per pass.  The reference has no counterpart (its encoder is order-oblivious, hashGridEmbedding.py:81-102); this
 * is the pre-pass that feeds `perm` of idrk_hash_encode_fwd / _bwd.  workspace: device scratch of at least
 * idrk_morton_sort_workspace(n) bytes, 16-byte aligned; n < 2^31. */
int idrk_morton_sort_workspace(int64_t n, int64_t* out_bytes);
int idrk_morton_sort(const float* x, int64_t n, int32_t ldx, const float* h_lo, const float* h_hi, int32_t bits_per_dim,
                     int32_t* perm, void* workspace, int64_t workspace_bytes, void* stream);
int idrk_scale3(const float* scale, const float* a, float* ya, int64_t na, const float* b, float* yb, int64_t nb,
                const float* c, float* yc, int64_t nc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IDRK_H_ */
