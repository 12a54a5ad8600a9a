/*
 * ub200.h — C ABI of the B200-native UNITER encoder hot path (libub200.so).
 *
 * The reference (ChenRocks/UNITER) is pure Python: its "FFI" for this path is the set of
 * torch / apex / horovod library calls made by model/layer.py and model/model.py.  Each entry
 * point below replaces one group of those call sites (cited as <file>:<lines> under the
 * reference root).  Rules of the boundary:
 *   - plain C: raw device pointers, explicit sizes / leading dimensions, POD structs;
 *   - the CALLER owns all memory (outputs, saved tensors, workspaces are pre-allocated);
 *   - no allocation, no host synchronisation, no exceptions across the ABI;
 *   - every launch goes onto the cudaStream_t passed in (never the legacy default stream);
 *   - return 0 on success, a negative UB200_E* code otherwise; ub200_last_error_string()
 *     gives the message for the calling thread.
 *
 * Data layout: activations are PACKED — row t of a [T, H] matrix is the t-th valid token of
 * the batch (sequence b occupies rows cu_seqlens[b] .. cu_seqlens[b+1]); there is no padding.
 * 16-bit storage type is selected by `dtype` (UB200_F16 / UB200_BF16); accumulation, softmax
 * and LayerNorm statistics are fp32.
 */
#ifndef UB200_H_
#define UB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* ub200_stream_t; /* == cudaStream_t */

enum { UB200_F16 = 0, UB200_BF16 = 1 };

enum {
  UB200_OK = 0,
  UB200_EINVAL = -1,      /* bad argument (null pointer, misaligned, negative size) */
  UB200_EUNSUPPORTED = -2, /* shape outside what the kernels implement */
  UB200_ECUDA = -3,       /* CUDA runtime / driver error, see last_error_string */
};

/* library identity ------------------------------------------------------------------------ */
int ub200_version(void);                      /* MAJOR*10000 + MINOR*100 + PATCH */
const char* ub200_last_error_string(void);    /* thread-local, never NULL */
int ub200_device_check(void);                 /* 0 iff the current device is sm_100 (B200) */

/* ------------------------------------------------------------------------------------------
 * GEMM core:  D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )      (tcgen05 + TMA, fp32 accumulate)
 *
 * Replaces every nn.Linear on the path and its autograd mirror:
 *   forward  model/layer.py:76-78 (query/key/value), :112 (attention.output.dense),
 *            :140 (intermediate.dense) + :31-37 (erf GELU), :153 (output.dense);
 *   backward torch autograd of the same nn.Linear modules (dgrad: B is MN-major = the weight
 *            read un-transposed; wgrad: A and B MN-major = activations read un-transposed).
 *
 * Operand storage:
 *   a_major == 0 : A is [M, K] row-major (K contiguous),  lda = row pitch in elements
 *   a_major == 1 : A is [K, M] row-major (M contiguous),  lda = row pitch in elements
 *   b_major == 0 : B is [N, K] row-major (K contiguous)   -- an nn.Linear weight as stored
 *   b_major == 1 : B is [K, N] row-major (N contiguous)
 * Epilogue, applied in this order to v = acc[m,n]:
 *   UB200_EPI_BIAS      v += bias[n]
 *   UB200_EPI_DROPOUT   v = keep(m,n) ? v / (1-p) : 0          (Philox, regenerated in bwd)
 *   UB200_EPI_RESIDUAL  v += residual[m,n]
 *   UB200_EPI_GELU      out2[m,n] = v ; v = gelu_erf(v)         (model/layer.py:31-37)
 *   UB200_EPI_DGELU     v *= gelu_erf'(aux[m,n])
 *   UB200_EPI_ACCUM     v += out[m,n]   (previous contents, e.g. gradient accumulation)
 *   UB200_EPI_OUT_F32   out is fp32 instead of the 16-bit dtype
 *   UB200_EPI_COLSUM    colsum[n] += sum_m v  (fp32 atomics; bias gradients)
 * ------------------------------------------------------------------------------------------ */
enum {
  UB200_EPI_BIAS = 1,
  UB200_EPI_DROPOUT = 2,
  UB200_EPI_RESIDUAL = 4,
  UB200_EPI_GELU = 8,
  UB200_EPI_DGELU = 16,
  UB200_EPI_ACCUM = 32,
  UB200_EPI_OUT_F32 = 64,
  UB200_EPI_COLSUM = 128,
};

typedef struct {
  const void* a;
  const void* b;
  int64_t lda, ldb;
  int32_t a_major, b_major;
  int32_t M, N, K;
  int32_t dtype;
  int32_t epilogue;        /* OR of UB200_EPI_* */
  const void* bias;        /* [N] 16-bit */
  const void* residual;    /* [M, N] 16-bit, pitch ldr */
  const void* aux;         /* [M, N] 16-bit, pitch ldaux (pre-activation for DGELU) */
  void* out;               /* [M, N], pitch ldo */
  void* out2;              /* [M, N] 16-bit, pitch ldo (pre-activation, GELU only) */
  float* colsum;           /* [N] fp32, accumulated with atomics */
  int64_t ldr, ldaux, ldo;
  float dropout_p;         /* 0 <= p < 1 */
  uint64_t rng_seed;       /* Philox key */
  uint64_t rng_stream;     /* distinguishes dropout sites / layers / steps */
  int32_t tile_n;          /* 0 = heuristic, else force 64 / 128 / 256 */
  int32_t max_ctas;        /* 0 = one CTA per SM */
} ub200_gemm_args;

int ub200_gemm(const ub200_gemm_args* args, ub200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UB200_H_ */
