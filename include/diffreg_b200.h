/*
 * diffreg_b200 -- C ABI of the B200 (sm_100a) kernels for Diff-Reg's per-step coarse
 * matching-matrix update.
 *
 * The reference has no FFI for this path: the boundary a maintainer sees is a set of
 * PyTorch modules / functions.  Every entry point below names the
 * reference interface (file:line, relative to the reference checkout) whose arithmetic it
 * replaces.  The Python drop-in modules in diff-reg_b200/ bind these through ctypes
 * (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - tensors are row-major contiguous fp32; masks are 1 byte per element (torch.bool);
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work, they never
 *     synchronise, allocate or free;
 *   - scratch memory comes from the caller: ask drg_*_workspace_bytes() and pass a buffer
 *     of at least that size (256-byte aligned);
 *   - return value 0 = ok, otherwise a DRG_ERR_* code and drg_last_error() describes it
 *     (thread-local string);
 *   - scalars that live on the device in the reference (the learnable dustbin score) are
 *     taken as device pointers so that no host read-back is needed.
 */
#ifndef DIFFREG_B200_H
#define DIFFREG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRG_OK 0
#define DRG_ERR_INVALID 1      /* bad argument (null pointer, non-positive size, ...) */
#define DRG_ERR_UNSUPPORTED 2  /* shape outside what the kernels handle */
#define DRG_ERR_WORKSPACE 3    /* workspace too small */
#define DRG_ERR_CUDA 4         /* a CUDA runtime call failed */

int drg_version(void);
const char* drg_last_error(void);
/* Number of kernels this library has launched since load (all streams); bench.py reports
 * the difference over the timed region as "gpu_launches". */
unsigned long long drg_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Log-domain Sinkhorn with dustbin row/column
 *   replaces log_optimal_transport(scores, alpha, iters, src_mask, tgt_mask)
 *     Diff-Reg-4dmatch/models/matching.py:6-38   (= Diff-Reg-3dmatch/models/matching.py:61-93,
 *     Diff-Reg-2d3d/experiments/<exp>/matching.py:6-38)
 *   plus, through the output modes, the consumers that follow it in the reference:
 *     .exp()[:, :-1, :-1].contiguous()                 matching.py:169-170, pipeline.py:215-216
 *     the DDIM update x_next = f(x0, x_t, noise)       Diff-Reg-4dmatch/models/pipeline.py:180-190,
 *                                                      Diff-Reg-3dmatch/models/pipeline.py:252-256
 *   and the dual-softmax branch                        matching.py:147-157
 * ------------------------------------------------------------------------------------ */

/* what the final pass writes */
#define DRG_OUT_LOG_FULL 0 /* out[B,N+1,M+1] = Z + u + v - norm              (matching.py:34-36)  */
#define DRG_OUT_CONF 1     /* out[B,N,M]     = exp(log-assignment)[:-1,:-1]  (matching.py:169-170) */
#define DRG_OUT_DDIM 2     /* out[B,N,M]     = k_x0*conf + k_xt*x_t + sigma*noise (pipeline.py:190) */
#define DRG_OUT_NONE 3     /* potentials only */

typedef struct drg_sinkhorn_args {
  /* problem */
  const float* scores;     /* [B,N,M]; may hold -inf at padded entries                          */
  const uint8_t* src_mask; /* [B,N] bool                                                        */
  const uint8_t* tgt_mask; /* [B,M] bool                                                        */
  const float* alpha;      /* device scalar: dustbin score (Matching.bin_score)                 */
  const float* shift;      /* device scalar or NULL: scores are read as (scores - *shift), the
                              3DMatch sampler's x - x.min()  (Diff-Reg-3dmatch pipeline.py:239) */
  int B, N, M;
  int iters;               /* skh_iters                                                         */
  int apply_mask;          /* 1: entries with an invalid src row or tgt column are treated as -inf
                              whatever is stored (fuses the masked_fill_ of matching.py:163-165) */
  /* outputs */
  int out_mode;            /* DRG_OUT_*                                                          */
  float* out;              /* see DRG_OUT_*; NULL for DRG_OUT_NONE                               */
  float* u;                /* [B,N+1] row potentials (optional, NULL = keep in workspace)        */
  float* v;                /* [B,M+1] column potentials (optional)                               */
  /* DRG_OUT_DDIM only */
  const float* x_t;        /* [B,N,M] current sampler state (read as x_t - *shift)               */
  const float* noise;      /* [B,N,M] N(0,1) draws or NULL (no noise term)                       */
  float* conf;             /* optional [B,N,M]: also store x0 = conf                             */
  float k_x0, k_xt, sigma; /* x_next = k_x0*conf + k_xt*x_t + sigma*noise                        */
  float* x_min;            /* optional device scalar: min over valid entries of x_next is folded
                              in with atomicMin (caller initialises to +inf)                     */
} drg_sinkhorn_args;

size_t drg_sinkhorn_workspace_bytes(int B, int N, int M);
int drg_sinkhorn(const drg_sinkhorn_args* args, void* workspace, size_t workspace_bytes, void* stream);

/* Dual-softmax confidence: conf = softmax_src(sim/T | src mask) * softmax_tgt(sim/T | tgt mask)
 *   replaces Diff-Reg-4dmatch/models/matching.py:147-157 (sim already divided by nothing:
 *   the temperature is applied here).  out[B,N,M]. */
int drg_dual_softmax(const float* sim, const uint8_t* src_mask, const uint8_t* tgt_mask, int B, int N, int M,
                     float temperature, float* out, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFREG_B200_H */
