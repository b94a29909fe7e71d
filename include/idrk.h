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
int idrk_version(void);                        /* ABI version, currently 1 */
int idrk_device_sm_count(int* out_sms);        /* SM count of the current device */

/* -- K1: hash-grid encode forward -------------------------------------------------------
 * Replaces MultiResHashGridMLP.forward (hashGridEmbedding.py:150-155) =
 * FourierFeature.forward (frequency_enc.py:63-67) ++ _HashGridMLP.forward x L (:81-102)
 * ++ hash_func (:32-40).   x [n, ldx>=3], out [n, ld_out], ld_out >= width; columns
 * width..ld_out-1 are written as zeros.  idx_debug (nullable) receives the uint32 table row
 * of all 8 corners, [n, L, 8]. */
int idrk_hash_encode_fwd(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx,
                         float* out, int32_t ld_out, uint32_t* idx_debug, void* stream);

/* -- K2: hash-grid encode backward ------------------------------------------------------
 * Replaces autograd through the same functions: embedding_dense_backward scatter-add into
 * every level's table gradient and d/dx of the Fourier prefix.  dy [n, ld_dy] is dL/d(out).
 * h_grad_tables: HOST array of L device pointers, each [T_l, F] (NULL = skip table gradients); ACCUMULATED
 * (caller zero-fills).  dx (nullable) [n, 3] is overwritten. */
int idrk_hash_encode_bwd(const idrk_hashgrid_t* h_grid, const float* x, int64_t n, int32_t ldx,
                         const float* dy, int32_t ld_dy, float* const* h_grad_tables,
                         float* dx, void* stream);

/* -- positional encoding ----------------------------------------------------------------
 * Replaces PositionalEncoding.embed (frequency_enc.py:19-51) and get_embedder (:156-168).
 * x [n, d] -> out [n, ld_out]; row = [x | x | sin(f_0 x) | cos(f_0 x) | ...] when include_input
 * (the input appears twice, as in the reference), h_bands = HOST array of n_bands frequencies.
 * Columns beyond the width are zero-filled.  bwd writes dx [n, d]. */
int idrk_posenc_fwd(const float* x, int64_t n, int32_t d, int32_t ldx, const float* h_bands, int32_t n_bands,
                    int32_t include_input, float* out, int32_t ld_out, void* stream);
int idrk_posenc_bwd(const float* x, int64_t n, int32_t d, int32_t ldx, const float* h_bands, int32_t n_bands,
                    int32_t include_input, const float* dy, int32_t ld_dy, float* dx, int32_t ld_dx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IDRK_H_ */
