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
/* Optional per-kernel timing used by bench.py's roofline leg: when enabled every launch of a slotted kernel is
 * bracketed by CUDA events on its stream.  Slots: 0 sinkhorn iteration, 1 column merge, 2 sinkhorn final/DDIM pass,
 * 3 sinkhorn prep, 4 similarity GEMM, 5 operand prep, 6 row/column best, 7 match rows, 8 top-K collect,
 * 9 Procrustes moments / solve, 10 top-K threshold, 11 Procrustes select.  drg_profile_read synchronises on the recorded events. */
void drg_profile_enable(int on);
void drg_profile_reset(void);
int drg_profile_read(int slot, double* total_ms, long long* count);
int drg_profile_slots(void);
/* sizeof(drg_sinkhorn_args) / sizeof(drg_procrustes_args) as this library was compiled: a binding checks its own mirror of
 * the argument structs against these before the first call (a shorter mirror would make the library read past its end). */
/* Tuning hook: a caller-owned DEVICE buffer of >= 1024 int64 in which CTA 0 of the persistent Sinkhorn (clock64 cycles) and
 * the pose kernel (globaltimer ns, slots 800+) leave stamps of their phases; NULL switches the stamps off (the default).
 * Used by tools/pose_timeline.py and tools/skh_timeline.py; the library itself never allocates. */
int drg_tuning_set_stamp_buffer(long long* device_buffer);
size_t drg_sizeof_sinkhorn_args(void);
size_t drg_sizeof_procrustes_args(void);

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
  const float* x_t;        /* [B,N,M] current sampler state (read as x_t - *xt_shift)            */
  const float* xt_shift;   /* device scalar or NULL: the 3DMatch sampler's x.min() for x_t; `shift`
                              above applies to `scores` only                                     */
  const float* noise;      /* [B,N,M] N(0,1) draws or NULL (no noise term)                       */
  float* conf;             /* optional [B,N,M]: also store x0 = conf                             */
  float k_x0, k_xt, sigma; /* x_next = k_x0*conf + k_xt*x_t + sigma*noise                        */
  float* x_min;            /* optional device scalar: min over valid entries of x_next is folded
                              in with atomicMin (caller initialises to +inf)                     */
  int gen_noise;           /* 1 and noise == NULL: draw the N(0,1) noise in the kernel (Philox4x32-7
                              + Box-Muller), replacing torch.randn_like(x) of pipeline.py:188    */
  unsigned long long noise_seed;   /* Philox key                                                  */
  unsigned long long noise_offset; /* high half of the Philox counter: use a new value per step   */
  const unsigned long long* noise_offset_dev; /* optional device counter added to noise_offset, so a
                              CUDA-graph replay of the same step draws fresh noise (drg_counter_add) */
  unsigned long long* rowbest; /* optional [B,N] (with colbest [B,M]; out_mode CONF or DDIM, M % 4 == 0): the final pass
                              also leaves every row's / column's best confidence and its lowest index as packed keys
                              (order-preserving float bits << 32 | ~index) for drg_match_from_best               */
  unsigned long long* colbest;
  int has_best_floor;      /* 1: only entries with confidence > best_floor are tracked in rowbest / colbest; rows / columns
                              without such an entry keep key 0 ("no best").  Exact for drg_match_from_best with a threshold
                              >= best_floor (Matching.get_match, matching.py:71-88: a match must exceed the threshold, and
                              whatever beats it in its row / column does too) and lets the pass skip the arg-max
                              bookkeeping for the ~97 % of warp-rows that hold no such entry.  0: track everything        */
  float best_floor;
} drg_sinkhorn_args;

size_t drg_sinkhorn_workspace_bytes(int B, int N, int M);
int drg_sinkhorn(const drg_sinkhorn_args* args, void* workspace, size_t workspace_bytes, void* stream);

/* Backward pass of the log-domain Sinkhorn (the training path): what torch's autograd computes for
 *   loss.backward() through log_optimal_transport   Diff-Reg-4dmatch/models/matching.py:6-38 (called at :167)
 * as 2 * iters "weighted exp" mat-vec passes over the scores and one final pass (2 iters + 1 reads, 1 write; deterministic).
 *   scores [B,N,M] as the forward saw them (-inf at masked entries), alpha device scalar, masks [B,N] / [B,M] bool (counts only),
 *   u_all [iters,B,N+1] / v_all [iters,B,M+1]: the potentials after each iteration t = 1..iters (drg_sinkhorn with iters = t,
 *   DRG_OUT_NONE), grad_out [B,N+1,M+1] = dL/d out  ->  grad_scores [B,N,M], grad_alpha [B] (per batch element; sum them). */
size_t drg_sinkhorn_backward_workspace_bytes(int B, int N, int M, int iters);
int drg_sinkhorn_backward(const float* scores, const float* alpha, const uint8_t* src_mask, const uint8_t* tgt_mask, int B, int N, int M,
                          int iters, const float* u_all, const float* v_all, const float* grad_out, float* grad_scores,
                          float* grad_alpha, void* workspace, size_t workspace_bytes, void* stream);

/* ... and of the dual-softmax branch (matching.py:147-157): grad_sim [B,N,M] from grad_conf [B,N,M]; sim is the undivided
 * similarity (the temperature is applied inside, as in drg_dual_softmax). */
size_t drg_dual_softmax_backward_workspace_bytes(int B, int N, int M);
int drg_dual_softmax_backward(const float* sim, const uint8_t* src_mask, const uint8_t* tgt_mask, int B, int N, int M, float temperature,
                              const float* grad_conf, float* grad_sim, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Row-sharded Sinkhorn: ONE matrix whose rows are spread over several GPUs (BASELINE.json configs[4]).
 *   The reference has no counterpart (its matrix always lives on one GPU); the arithmetic is
 *   log_optimal_transport's (Diff-Reg-4dmatch/models/matching.py:6-38) with the column log-sum-exp split as
 *   LSE_i = LSE over ranks of (LSE over the rank's rows).
 *   Every call takes the rank's LOCAL problem in drg_sinkhorn_args (scores [B,Nloc,M], src_mask [B,Nloc], N = Nloc)
 *   and the same workspace (drg_sinkhorn_workspace_bytes(B, Nloc, M)).  Per iteration:
 *     drg_sinkhorn_shard_local   row pass + this rank's column partials -> partial [B, M+1, 2] = (max, sum) in the log2 domain
 *     (caller)                   all-reduce the partials over the ranks: MAX on max, SUM on sum * 2^(max_local - max_global)
 *     drg_sinkhorn_shard_update  v from the reduced partials (+ the dustbin row, identical on every rank)
 *   drg_sinkhorn_shard_begin sets the normalisation constants from global_counts [B,2] = (valid src rows over ALL ranks,
 *   valid tgt columns) and zeroes the potentials; drg_sinkhorn_shard_final writes the rank's rows of the output (out_mode).
 * ------------------------------------------------------------------------------------ */
int drg_sinkhorn_shard_begin(const drg_sinkhorn_args* args, const int* global_counts, void* workspace, size_t workspace_bytes,
                             void* stream);
int drg_sinkhorn_shard_local(const drg_sinkhorn_args* args, void* workspace, size_t workspace_bytes, float* partial, void* stream);
int drg_sinkhorn_shard_update(const drg_sinkhorn_args* args, void* workspace, size_t workspace_bytes, const float* reduced,
                              void* stream);
int drg_sinkhorn_shard_final(const drg_sinkhorn_args* args, void* workspace, size_t workspace_bytes, void* stream);

/* The same iteration with the all-reduce INSIDE the kernel (one process per GPU on one NVLink / NVSwitch node):
 *   drg_p2p_create     allocates this rank's exchange buffer (inboxes for `slot_elems` (max, sum) pairs per sender --
 *                      at least B * (M + 1) -- and `nflags` >= B * ceil((M + 1) / 32) flags), returns an opaque comm and
 *                      its CUDA IPC handle (drg_p2p_handle_bytes() bytes) for the other ranks
 *   drg_p2p_connect    maps every rank's buffer: `handles` = the world handles, rank-major (exchange them with any
 *                      host-side all-gather)
 *   drg_sinkhorn_shard_local_exchange   = shard_local + all-reduce + shard_update: after the row pass one kernel merges
 *                      this rank's column partials, stores them into every rank's inbox over NVLink, waits for the
 *                      peers' partials of the same columns and updates v (bit-identical on every rank).  Every rank must
 *                      issue the same sequence of calls on a comm.
 *   drg_p2p_status     0, or 1 if a wait ever timed out (a peer never arrived; that call's result is invalid)      */
size_t drg_p2p_handle_bytes(void);
int drg_p2p_create(int rank, int world, size_t slot_elems, int nflags, void** comm_out, void* handle_out);
int drg_p2p_connect(void* comm, const void* handles);
int drg_p2p_status(void* comm);
int drg_p2p_destroy(void* comm);
int drg_sinkhorn_shard_local_exchange(const drg_sinkhorn_args* args, void* workspace, size_t workspace_bytes, void* comm,
                                      void* stream);
/* `iters` iterations of drg_sinkhorn_shard_local_exchange enqueued back to back (one host call per Sinkhorn instead of one
 * per iteration: at 8 GPUs an iteration is shorter than a host round trip). */
int drg_sinkhorn_shard_iterate(const drg_sinkhorn_args* args, void* workspace, size_t workspace_bytes, void* comm, int iters,
                               void* stream);

/* Dual-softmax confidence: conf = softmax_src(sim/T | src mask) * softmax_tgt(sim/T | tgt mask)
 *   replaces Diff-Reg-4dmatch/models/matching.py:147-157 (sim already divided by nothing:
 *   the temperature is applied here).  out[B,N,M]. */
int drg_dual_softmax(const float* sim, const uint8_t* src_mask, const uint8_t* tgt_mask, int B, int N, int M,
                     float temperature, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Feature-similarity contraction on the tensor cores (tcgen05 / TMEM / TMA)
 *   C[b] = alpha * A[b] . B[b]^T,  A [batch,N,K], B [batch,M,K], C [batch,N,M], all fp32 row-major.
 *   replaces torch.einsum("bsc,btc->bst", src, tgt)   Diff-Reg-4dmatch/models/matching.py:149,161
 *            (= Diff-Reg-3dmatch/models/matching.py:195,207; Diff-Reg-2d3d/experiments/<exp>/matching.py:110,122)
 *   and, with B = the weight [C_out, C_in], the nn.Linear projections  matching.py:127-128.
 *   Arithmetic: tcgen05.mma kind::tf32 with fp32 accumulation (10 mantissa bits of each operand are used).
 *   Requires K % 4 == 0 and 16-byte aligned A, B.  No workspace.
 * ------------------------------------------------------------------------------------ */
int drg_gemm_nt_tf32(const float* A, const float* B, float* C, int batch, int N, int M, int K, float alpha, void* stream);

/* The fp32-accurate product (what the reference's einsum computes with torch's allow_tf32 = False) from 16-bit SPLIT operands.
 *   A split operand of a [rows, K] matrix is a 16-bit array [rows, 2 kc + 8], kc = K rounded up to 64:
 *     every row is scaled by a power of two 2^e that puts its largest magnitude into [2^14, 2^15) (fp16's range);
 *     x 2^e = hi + lo with hi = fp16(x 2^e) (11 significant bits) and lo = fp16(x 2^e - hi);
 *     the row holds [lo | hi | tail] (left operand, pattern 0) or [hi | lo | tail] (right operand, pattern 1): two segments of
 *     kc columns (padding zero) and a tail of four floats (2^-e, Euclidean norm of the row, 0, 0).
 *   The kernel accumulates lo.hi + hi.lo + hi.hi with tcgen05.mma kind::f16 in fp32 (dropped: lo.lo ~ 2^-22 relative) and
 *   multiplies 2^-e of the row and of the column back in its epilogue (exact).  drg_prep_operand(split = 1) and the split
 *   epilogue below produce the operands.  K here is the column count of the ORIGINAL operand.  A16 [batch, N, 2 kc + 8],
 *   B16 [batch, M, 2 kc + 8].  (fp16 x bf16 products are not available on the tensor core: kind::f16 takes one format.) */
int drg_gemm_nt_split16(const void* A16, const void* B16, float* C, int batch, int N, int M, int K, float alpha, void* stream);

/* Projection with the operand preparation of the similarity GEMM fused into its epilogue:
 *   split_out16[rows, 2 * kc_out + 8] (16-bit split operand, kc_out = C_out rounded up to 64; the caller zeroes padding columns
 *   once) = split(scale * A . W^T): rows < rows_left as the left operand [lo | hi | tail], the others as the right operand
 *   [hi | lo | tail]; an output row's scale comes from the bound |y_ij| <= ||x_i|| max_j ||W_j|| (the norms travel in the
 *   operands' row tails); plain_out (optional) [rows, C_out] fp32 = A . W^T (what the reference stores in data[...]).
 *   replaces src_proj on both feature sets + the 1/sqrt(C) scaling  Diff-Reg-4dmatch/models/matching.py:127-128,144-145
 *   A16 [rows, 2 kc + 8], W16 [C_out, 2 kc + 8]: split operands of the features (pattern 0) and of the weight (pattern 1). */
int drg_project_split16(const void* A16, const void* W16, int rows, int rows_left, int C_out, int K, float scale, float* plain_out,
                        void* split_out16, void* stream);
/* drg_prep_operand over two source tensors in one launch: out rows [0, rows_a) from in_a (pattern_a), then rows_b rows from in_b. */
int drg_prep_operand_pair(const float* in_a, long long rows_a, int pattern_a, const float* in_b, long long rows_b, int pattern_b,
                          int K, float scale, int split, void* out, void* stream);

/* Volumetric position code from point coordinates (next-row widening)
 *   replaces VolumetricPositionEncoding.forward / voxelize   Diff-Reg-4dmatch/models/position_encoding.py:16-24,49-87
 *   xyz [points,3] (device), div_term [feature_dim/6] (device; exp(arange(0, d/3, 2) * -ln(1e4) / (d/3)) as the reference
 *   computes it), origin3 = vol_bnds[0] (HOST pointer to 3 floats), pe_type 1 rotary -> out [points, d, 2] (cos, sin),
 *   2 sinusoidal -> out [points, d].  feature_dim must be a multiple of 6 (as in the reference). */
int drg_position_code(const float* xyz, const float* div_term, long long points, int feature_dim, const float* origin3,
                      float voxel_size, int pe_type, float* out, void* stream);

/* Operand preparation: positional embedding + scaling + optional hi/lo split.
 *   replaces VolPE.embed_pos / embed_rotary   Diff-Reg-4dmatch/models/position_encoding.py:26-46
 *   and       feat / feat.shape[-1] ** .5     Diff-Reg-4dmatch/models/matching.py:144-145
 *   in [rows,K]; pe: rotary [rows,K,2] (cos,sin) for pe_type 1, additive [rows,K] for pe_type 2, NULL for 0;
 *   embedded (optional) [rows,K] receives the features after the embedding and before scaling (data["src_feats"]);
 *   out: split=0 -> fp32 [rows,K] = scale*x;  split=1 -> the 16-bit split operand [rows, 2*kc + 8] of scale*x (pattern 0:
 *   [lo | hi | tail], pattern 1: [hi | lo | tail]; see drg_gemm_nt_split16). */
int drg_prep_operand(const float* in, const float* pe, int pe_type, long long rows, int K, float scale, int split, int pattern,
                     float* embedded, void* out, void* stream);

/* drg_prep_operand with the position code computed inside the kernel from the point coordinates (SURVEY.md 8f rank 1:
 *   VolumetricPositionEncoding.forward fused into the GEMM operand staging): xyz [rows,3], div_term [K/6] and origin3 (HOST
 *   pointer) / voxel_size as for drg_position_code; pe_type 1 rotary, 2 sinusoidal.  Bit-identical to
 *   drg_position_code + drg_prep_operand, without the [rows,K,2] / [rows,K] code tensor in HBM.
 *   replaces Diff-Reg-4dmatch/models/position_encoding.py:49-87 + :26-46 as called at models/transformer.py:165-166 and
 *   models/matching.py:135-137 */
int drg_prep_operand_xyz(const float* in, const float* xyz, const float* div_term, const float* origin3, float voxel_size,
                         int pe_type, long long rows, int K, float scale, int split, int pattern, float* embedded, void* out,
                         void* stream);

/* drg_prep_operand with two more staging options (the attention layer of the denoising transformer):
 *   relu  = 1: max(x, 0) before the scaling -- the nn.ReLU between the two linears of the layer's MLP
 *              Diff-Reg-4dmatch/models/transformer.py:33-37
 *   heads > 1: `in` is [B, seq, heads, K] (rows = B * seq * heads, one row per (b, l, h) head slice; a rotary / additive code
 *              tensor is indexed the same way) and the output rows are head-major (b, h, l): the per-head operands of
 *              torch.einsum("nlhd,nshd->nlsh", qw, kw)   transformer.py:79   leave the staging kernel as [B * heads, seq, .] */
int drg_prep_operand_ext(const float* in, const float* pe, int pe_type, long long rows, int K, float scale, int split, int pattern,
                         int relu, int heads, int seq, float* embedded, void* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Geometry attention layer (denoising transformer): the row kernels between its GEMMs
 *   drg_attn_softmax   logits [B*H, L, S] (the batched Q.K^T of drg_gemm_nt_split16) -> attention probabilities:
 *                      a.masked_fill_(q_mask[:, :, None, None] * (~kv_mask[:, None, :, None]), -inf); a = a / sqrt(d);
 *                      a = softmax(a, dim=2)             Diff-Reg-4dmatch/models/transformer.py:80-84
 *                      (keys are masked for VALID queries only, as the reference's expression does; a valid query without a
 *                      valid key yields NaN, as there).  q_mask [B, L], kv_mask [B, S] bool or NULL; scale = 1 / sqrt(d).
 *                      P (optional) [B*H, L, S] fp32, may alias logits; P16 (optional) the same rows as the LEFT split
 *                      operand [B*H, L, 2 kc(S) + 8] of the P.V product (drg_gemm_nt_split16).
 *   drg_layernorm      out = [residual +] LayerNorm(in) over the last dimension C, affine (weight / bias may be NULL); pre_add = 0
 *                      replaces self.norm1(message), x + self.norm2(message)     transformer.py:88,92-94
 * ------------------------------------------------------------------------------------ */
int drg_attn_softmax(const float* logits, const uint8_t* q_mask, const uint8_t* kv_mask, int B, int H, int L, int S, float scale,
                     float* P, void* P16, void* stream);
int drg_layernorm(const float* in, const float* weight, const float* bias, const float* residual, int pre_add, long long rows, int C,
                  float eps, float* out, void* split16_out /* optional [rows, 2 kc(C) + 8]: out also as the LEFT split operand */,
                  void* stream);
/* 2D-3D flavour of the same row (CrossModalFusionModule, Diff-Reg-2d3d/experiments/<exp>/fusion_module.py:10-107, built from
 * vision3d's TransformerLayer, Diff-Reg-2d3d/vision3d/layers/transformer.py:8-301):
 *   drg_layernorm with pre_add = 1   out = LayerNorm(in + residual)            transformer.py:214,236 (post-norm blocks)
 *   drg_gemm_nt_split16_bias         C = alpha * A . B^T + bias[M]             the nn.Linear layers with bias of that module
 *   drg_fourier_embed                [x | sin(2^l x), cos(2^l x), l < L] of (x - center)
 *                                    replaces FourierEmbedding.forward  vision3d/layers/embedding.py:75-99 and the centring of
 *                                    create_3d_embedding  fusion_module.py:56-60; out [rows, n * (2 L + use_input)]          */
int drg_gemm_nt_split16_bias(const void* A16, const void* B16, const float* bias, float* C, int batch, int N, int M, int K, float alpha,
                             void* stream);
/* Fused attention: out[b, l, h, :] = softmax_s(Q[b,h,l,:] . K[b,h,s,:] * scale + mask) . V[b,h,s,:] in ONE kernel -- logits in
 * TMEM, probabilities split to fp16 hi / lo on the fly, O accumulated in TMEM; the [L, S] attention matrix never reaches HBM.
 *   replaces einsum / masked_fill / softmax / einsum   Diff-Reg-4dmatch/models/transformer.py:79-85 and
 *            einsum / masked_fill / softmax / matmul   Diff-Reg-2d3d/vision3d/layers/transformer.py:127-154
 *   Q16 [B*H, L, 2 kc(d) + 8]   LEFT split operand per head  (drg_prep_operand_ext, pattern 0, head-major)
 *   K16 [B*H, S, 2 kc(d) + 8]   RIGHT split operand per head (pattern 1)
 *   Vt16 [B*H, d, 2 kc(S) + 8]  RIGHT split operand of V^T per head (keys along the row)
 *   q_mask [B, L] / kv_mask [B, S] bool or NULL, 1 = valid: the keys with kv_mask == 0 are masked for the queries with
 *   q_mask != 0 (all queries when q_mask is NULL) -- the reference's expression; a valid query without a valid key yields NaN.
 *   out [B, L, H * d] fp32.  d % 4 == 0, d <= 176 (DRG_ERR_UNSUPPORTED otherwise: the caller keeps the three-kernel path). */
/*   drg_prep_vt_split16: V [B, S, H * d] fp32 -> Vt16 as above (per head V^T, the keys along the row; the tail holds 1 / scale
 *   only).  workspace: B * H * d * 4 bytes (the channels' maxima). */
int drg_prep_vt_split16(const float* V, int B, int H, int S, int d, void* Vt16, void* workspace, void* stream);
/*   nsplit: the keys of one (query tile, head) are shared by nsplit CTAs whose partial results a second kernel combines (small
 *   grids: 2048 queries x 4 heads are 64 CTAs for 148 SMs); 0 = chosen by the library, 1 = never split.  workspace:
 *   drg_attention_workspace_bytes(B, H, L, S, d, nsplit) bytes for the same nsplit (0 when the call will not split: NULL is fine). */
size_t drg_attention_workspace_bytes(int B, int H, int L, int S, int d, int nsplit);
int drg_attention_split16(const void* Q16, const void* K16, const void* Vt16, const uint8_t* q_mask, const uint8_t* kv_mask, int B,
                          int H, int L, int S, int d, float scale, float* out, int nsplit, void* workspace, size_t workspace_bytes,
                          void* stream);
int drg_fourier_embed(const float* x, const float* center, long long rows, int n, int length, float k0, int use_pi, int use_input,
                      float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Correspondence extraction
 *   mode 0: Matching.get_match(conf, thr, mutual)       Diff-Reg-4dmatch/models/matching.py:71-88 (= get_topk_match :90-107)
 *           hit = conf > thr [and conf == row max and conf == column max]
 *   mode 1: mutual_topk_select(score, k=1, largest, threshold, mutual)
 *                                                       Diff-Reg-2d3d/vision3d/ops/mutual_topk_select.py:7-60
 *           hit = (top-1 of its row) AND/OR (top-1 of its column) [and score > threshold]
 *   Two calls because the number of matches is data dependent (the reference's nonzero() synchronises too):
 *   drg_match_count leaves the total in *total_out (device int); the caller reads it, allocates
 *   index_out [total,3] int64 (b, row, col; row-major order like nonzero()) and val_out [total], and calls
 *   drg_match_write with the SAME arguments and workspace.  `capacity` is the number of slots in index_out / val_out
 *   (hits beyond it are dropped, so a caller that sizes the outputs by an upper bound never needs the host read).
 *   mask_out (optional, mode 0) is the [B,N,M] bool mask.
 * ------------------------------------------------------------------------------------ */
size_t drg_match_workspace_bytes(int B, int N, int M);
int drg_match_count(const float* x, int B, int N, int M, int mode, int mutual, int has_thr, float thr, int largest, void* workspace,
                    size_t workspace_bytes, int* total_out, void* stream);
int drg_match_write(const float* x, int B, int N, int M, int mode, int mutual, int has_thr, float thr, int largest, void* workspace,
                    size_t workspace_bytes, long long* index_out, float* val_out, long long capacity, unsigned char* mask_out,
                    void* stream);

/* Top-k selection for k >= 1 (k <= 8), batched, with optional row / column masks applied after the selection:
 *   replaces mutual_topk_select(score_mat, k, ...) and batch_mutual_topk_select(score_mat, k, row_masks, col_masks, ...)
 *            Diff-Reg-2d3d/vision3d/ops/mutual_topk_select.py:7-60, 63-133 (the 2D-3D fine matching calls it with k = 2,
 *            threshold 0.75 on [B, Kc, Kc] patch similarities: experiments/<exp>/model.py:738-746)
 *   hit = (within the top-k of its row) AND/OR (within the top-k of its column) [and score > threshold] [and masks];
 *   ties are broken towards the lower index.  Same two-call protocol and workspace as drg_match_count / drg_match_write
 *   (mode 1 shapes: x [B,N,M]); index_out rows are (b, row, col) in row-major order. */
int drg_topk_match_count(const float* x, int B, int N, int M, int k, int mutual, int has_thr, float thr, int largest,
                         const unsigned char* row_mask, const unsigned char* col_mask, void* workspace, size_t workspace_bytes,
                         int* total_out, void* stream);
int drg_topk_match_write(const float* x, int B, int N, int M, int k, int mutual, int has_thr, float thr, int largest,
                         const unsigned char* row_mask, const unsigned char* col_mask, void* workspace, size_t workspace_bytes,
                         long long* index_out, float* val_out, long long capacity, unsigned char* mask_out, void* stream);

/* Mutual top-1 matches from the packed row / column bests written by drg_sinkhorn (rowbest / colbest): row i matches its
 * best column j iff j's best row is i [and conf > thr].  Same hits as Matching.get_match(conf, thr, mutual=True)
 * (Diff-Reg-4dmatch/models/matching.py:71-88) and mutual_topk_select(k=1, mutual=True) except at exact value ties, where
 * the lowest index wins.  O(B*N) work: the confidence matrix is not read.  index_out [capacity,3] int64 in row-major order,
 * val_out [capacity], *total_out = number of matches (device int). */
int drg_match_from_best(const unsigned long long* rowbest, const unsigned long long* colbest, int B, int N, int M, int has_thr,
                        float thr, long long* index_out, float* val_out, long long capacity, int* total_out, void* stream);

/* ------------------------------------------------------------------------------------
 * SoftProcrustes
 *   replaces SoftProcrustesLayer.forward(conf, src_pcd, tgt_pcd, src_mask, tgt_mask)
 *            Diff-Reg-4dmatch/models/procrustes.py:48-93 (3DMatch: Diff-Reg-3dmatch/models/procrustes.py:61-62)
 *   and the warp of the source points by the gated pose, Diff-Reg-4dmatch/models/pipeline.py:220.
 * ------------------------------------------------------------------------------------ */
typedef struct drg_procrustes_args {
  const float* conf;             /* [B,N,M]                                                          */
  const float* src_pcd;          /* [B,N,3]                                                          */
  const float* tgt_pcd;          /* [B,M,3]                                                          */
  const uint8_t* src_mask;       /* [B,N] bool                                                       */
  const uint8_t* tgt_mask;       /* [B,M] bool                                                       */
  int B, N, M;
  float sample_rate;             /* config.sample_rate                                               */
  float max_condition_num;       /* config.max_condition_num                                         */
  int padded_lengths;            /* 1: 3DMatch variant, lengths are N and M whatever the masks say   */
  float* R;                      /* [B,3,3]                                                          */
  float* t;                      /* [B,3,1]                                                          */
  float* R_forwd;                /* [B,3,3] identity where the condition gate fails                  */
  float* t_forwd;                /* [B,3,1] zero where the condition gate fails                      */
  double* condition;             /* [B] fp64 (the reference returns it on the CPU; here on the device) */
  uint8_t* solution_mask;        /* [B] bool                                                         */
  float* src_warped;             /* optional [B,N,3]: R_forwd s + t_forwd                            */
  int K_max;                     /* slots per batch element in sel_* (>= max(N,M)*sample_rate)       */
  float* sel_w;                  /* optional [B,K_max]: weights of the selected correspondences      */
  int* sel_src;                  /* optional [B,K_max]: their src indices                            */
  int* sel_tgt;                  /* optional [B,K_max]: their tgt indices                            */
} drg_procrustes_args;

size_t drg_soft_procrustes_workspace_bytes(int B, int N, int M);
int drg_soft_procrustes(const drg_procrustes_args* args, void* workspace, size_t workspace_bytes, void* stream);

/* Noisy matching -> pose in one call: log-Sinkhorn on the sampler state followed by SoftProcrustes + warp, WITHOUT
 * materialising the confidence matrix (the top-K search recomputes exp(Z+u+v-norm) from the potentials).
 *   replaces get_warped_from_noising_matching   Diff-Reg-4dmatch/models/pipeline.py:207-223 (3d :293-309,
 *            2d3d get_warped_from_noising_matching3D3D model.py:830-846)
 *   s: the Sinkhorn problem with out_mode DRG_OUT_NONE; a: the Procrustes problem with conf = NULL and the same masks. */
int drg_sinkhorn_soft_procrustes(const drg_sinkhorn_args* s, const drg_procrustes_args* a, void* skh_workspace,
                                 size_t skh_workspace_bytes, void* procr_workspace, size_t procr_workspace_bytes, void* stream);

/* Weighted Kabsch on given correspondences.
 *   replaces SoftProcrustesLayer.batch_weighted_procrustes(X, Y, w, eps)  Diff-Reg-4dmatch/models/procrustes.py:18-44
 *   X, Y [B,K,3], w [B,K] -> R [B,3,3], t [B,3,1], condition [B] fp64 */
int drg_weighted_procrustes(const float* X, const float* Y, const float* w, int B, int K, float eps, float* R, float* t,
                            double* condition, void* stream);

/* ... and its backward pass with respect to the weights (the reference's motion loss differentiates R, t through the host SVD,
 * Diff-Reg-4dmatch/models/loss.py:110-131): grad_R [B,3,3], grad_t [B,3,1] -> grad_w [B,K]; R = the forward's result. */
int drg_weighted_procrustes_backward(const float* X, const float* Y, const float* w, const float* R, const float* grad_R,
                                     const float* grad_t, int B, int K, float eps, float* grad_w, void* stream);

/* Correspondence RANSAC: the consumer of get_match's correspondences in the 3DMatch / 4DMatch evaluation.
 *   replaces ransac_pose_estimation(src_pcd, tgt_pcd, corrs, distance_threshold, ransac_n)   Diff-Reg-4dmatch/models/loss.py:13-24
 *            (open3d 0.13 registration_ransac_based_on_correspondence, TransformationEstimationPointToPoint(False),
 *             RANSACConvergenceCriteria(50000, ...)) inside MatchMotionLoss.ransac_regist_coarse   loss.py:366-398
 *   src [B,N,3], tgt [B,M,3]; match [C,3] int64 rows (b, i, j) grouped by b -- get_match's match_pred as it stands;
 *   num_match = C; offsets [B+1] int32 DEVICE: first row of every batch element, or NULL: found on the device by binary
 *   search on the batch column (rows grouped by ascending b).  Fewer than 3 rows -> identity (loss.py:384-387).
 *   Point numbers outside [0, N) / [0, M) are clamped (the reference raises an IndexError on the host).
 *   Every one of the max_iteration trials runs (one thread each): ransac_n draws with replacement from a counter-based
 *   generator (seed, b, trial, draw), rigid fit, inlier count with |R s + t - g| < max_correspondence_distance; the best trial
 *   is the one with the most inliers, then the smaller rmse, then the lower trial number.
 *   -> pose [B,4,4] row-major, fitness [B] (inliers / C), inlier_rmse [B], best_trial [B] (-1: identity returned),
 *      inlier_count [B]; optional per-trial records trial_count [B,max_iteration] (-1: degenerate sample) and
 *      trial_err2 [B,max_iteration] (both or neither). */
size_t drg_ransac_workspace_bytes(int B, int max_iteration);
int drg_ransac_correspondence(const float* src, const float* tgt, int B, int N, int M, const long long* match, long long num_match,
                              const int* offsets, float max_correspondence_distance, int ransac_n, int max_iteration,
                              unsigned long long seed, float* pose, float* fitness, float* inlier_rmse, int* best_trial, int* inlier_count,
                              int* trial_count, float* trial_err2, void* workspace, size_t workspace_bytes, void* stream);

/* Elementwise tail / head of the samplers.
 *   drg_sigmoid:   conf_matrix_pred = sigmoid(x)        Diff-Reg-4dmatch/models/pipeline.py:192
 *   drg_min_value: x.min() of the 3DMatch sampler       Diff-Reg-3dmatch/models/pipeline.py:239,264
 *                  (scratch: one device uint32; *out receives the minimum) */
int drg_sigmoid(const float* x, float* y, long long n, void* stream);
/* *counter += inc on the stream (the Philox offset of a graph-captured sampler step). */
int drg_counter_add(unsigned long long* counter, unsigned long long inc, void* stream);
int drg_min_value(const float* x, long long n, float* out, unsigned int* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFREG_B200_H */
