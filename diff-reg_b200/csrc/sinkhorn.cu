// Fused log-domain Sinkhorn (dustbin row/column, masked padding), dual-softmax statistics
// and the final exp / DDIM pass.  sm_100a.
//
// Replaces log_optimal_transport (Diff-Reg-4dmatch/models/matching.py:6-38) and its
// consumers; see include/diffreg_b200.h.
//
// Layout in HBM: scores [B,N,M] fp32 row-major, never copied into an (N+1)x(M+1) matrix --
// the dustbin row and column are the constant alpha and are handled analytically.
//
// One Sinkhorn iteration = ONE read of the score matrix:
//   skh_iter_kernel   persistent CTAs, each owns a contiguous range of R-row slabs.  A slab
//                     (R x M fp32, one contiguous chunk) is brought into shared memory by a
//                     single TMA bulk copy (cp.async.bulk + mbarrier, 2-3 stages in flight).
//                     From shared memory the CTA computes (a) the row log-sum-exp of
//                     Z + v  ->  u_i, and then, with the fresh u_i, (b) this slab's
//                     contribution to every column's log-sum-exp of Z + u, kept as
//                     per-thread (max, sum) register accumulators over all the CTA's slabs.
//   skh_col_kernel    merges the G per-CTA column partials (+ the dustbin row term) into v.
// All log-sum-exps are carried in the log2 domain so that each element costs one FFMA
// and one MUFU.EX2 per direction.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace drg {

constexpr int SKH_THREADS = 512;
constexpr int SKH_WARPS = SKH_THREADS / 32;
constexpr int SKH_STAGE_FLOATS = 16384;  // R * M <= 16384 floats (64 KB) per stage
constexpr int SKH_MAX_M = 16384;
constexpr size_t SKH_SMEM_LIMIT = 227 * 1024;

// The final pass (exp / DDIM update / in-kernel noise / arg-max keys) as the LAST PHASE of the persistent kernel: the score
// slab is L2-hot right after the iterations and the potentials are final, so the pass costs one read of x_t and one write of
// the output in DRAM instead of three matrices, and one launch boundary less.  out == NULL: the stand-alone final kernels run.
struct SkhFused {
  float* out;                 // [B,N,M] conf (ddim == 0) or x_next (ddim == 1)
  const float* x_t;           // ddim: [B,N,M]
  const float* xt_shift;      // device scalar or NULL
  const float* noise;         // [B,N,M] caller-supplied N(0,1) draws or NULL
  float* conf;                // ddim: optional x0 = conf
  unsigned long long* rowbest;  // optional packed arg-max keys (entries above best_floor only)
  unsigned long long* colbest;
  float k_x0, k_xt, sigma, best_floor;
  int ddim, gen_noise;
  unsigned long long noise_offset;
  const unsigned long long* noise_offset_dev;
  unsigned int rk[14];        // Philox round keys (seed_lo + r W0, seed_hi + r W1), r = 0..6, precomputed on the host
};

struct SkhParams {
  const float* scores;
  const uint8_t* src_mask;
  const uint8_t* tgt_mask;
  const float* alpha;
  const float* shift;
  float* u;         // [B, N+1]
  float* v;         // [B, M+1]
  float2* colpart;  // [B, G, M]   (max, sum) in the log2 domain
  float2* upart;    // [B, G]
  const SkhConst* bc;
  SkhConst* bc_out;  // persistent kernel: writes the constants it computes
  float2* shard_partial;  // row-sharded mode: the column merge writes (max, sum) per column here instead of updating v
  int B, N, M, G;
  int ldu, ldv;  // row pitch of u / v (multiples of 4 floats)
  int apply_mask;
  int dual;       // 1: dual-softmax statistics (no dustbins, no potentials)
  float zscale2;  // log2(e) (Sinkhorn) or log2(e)/temperature (dual softmax)
  int nstage;
  long long* dbg_times;  // tuning only: CTA 0 writes clock64() stamps here (NULL = off; drg_tuning_set_stamp_buffer)
  unsigned long long* zero_a;  // optional: arrays the persistent kernel clears on its way in (rowbest / colbest of the
  unsigned long long* zero_b;  //   final pass of the same call: two memset nodes less per step)
  size_t zero_a_n, zero_b_n;
  SkhCollect col;  // col.state != NULL: the persistent kernel runs the top-K candidate search of SoftProcrustes as its last phase
  SkhFused fin;    // fin.out != NULL: the persistent kernel runs the final pass as its last phase
};

// ---------------------------------------------------------------------------------------
// prep: mask counts -> constants; v = 0
// ---------------------------------------------------------------------------------------
// 16 mask bytes per load when the row of masks is 16-byte aligned, else byte loads
__device__ __forceinline__ int count_mask_bytes(const uint8_t* __restrict__ m, int n, int tid, int nthreads) {
  int c = 0;
  if ((((uintptr_t)m) & 15u) == 0) {
    const int n16 = n >> 4;
    const uint4* m4 = reinterpret_cast<const uint4*>(m);
    for (int i = tid; i < n16; i += nthreads) {
      const uint4 q = m4[i];
      // bools are 0 / 1 bytes: the popcount of the low bit of every byte
      c += __popc(q.x & 0x01010101u) + __popc(q.y & 0x01010101u) + __popc(q.z & 0x01010101u) + __popc(q.w & 0x01010101u);
    }
    for (int i = (n16 << 4) + tid; i < n; i += nthreads) c += m[i] ? 1 : 0;
  } else {
    for (int i = tid; i < n; i += nthreads) c += m[i] ? 1 : 0;
  }
  return c;
}

__global__ void __launch_bounds__(1024) skh_prep_kernel(const uint8_t* __restrict__ src_mask, const uint8_t* __restrict__ tgt_mask,
                                                        int N, int M, SkhConst* __restrict__ bc, float* __restrict__ v,
                                                        float* __restrict__ u, int ldu, int ldv) {
  const int b = blockIdx.x;
  __shared__ int cnt[2];
  if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
  __syncthreads();
  int cs = count_mask_bytes(src_mask + (size_t)b * N, N, threadIdx.x, blockDim.x);
  int ct = count_mask_bytes(tgt_mask + (size_t)b * M, M, threadIdx.x, blockDim.x);
  cs = __reduce_add_sync(0xffffffffu, cs);
  ct = __reduce_add_sync(0xffffffffu, ct);
  if ((threadIdx.x & 31) == 0) {
    if (cs) atomicAdd(&cnt[0], cs);
    if (ct) atomicAdd(&cnt[1], ct);
  }
  // u, v rows are padded to multiples of 4 floats and 16-byte aligned
  float4* v4 = reinterpret_cast<float4*>(v + (size_t)b * ldv);
  float4* u4 = reinterpret_cast<float4*>(u + (size_t)b * ldu);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = threadIdx.x; j < (ldv >> 2); j += blockDim.x) v4[j] = z4;
  for (int i = threadIdx.x; i < (ldu >> 2); i += blockDim.x) u4[i] = z4;
  __syncthreads();
  if (threadIdx.x == 0) {
    // reference: norm = -(ms+ns).log() on int64 -> fp32 (matching.py:24); ns.log() + norm (:26-27)
    float ms = (float)cnt[0], ns = (float)cnt[1];
    float norm = -logf(ms + ns);
    SkhConst c;
    c.norm = norm;
    c.log_mu_bin = logf(ns) + norm;
    c.log_nu_bin = logf(ms) + norm;
    c.pad = (cnt[0] == N && cnt[1] == M) ? 1.f : 0.f;  // 1: no padded row or column (the final pass skips its mask logic)
    bc[b] = c;
  }
}

// ---------------------------------------------------------------------------------------
// one Sinkhorn iteration over the score matrix
// ---------------------------------------------------------------------------------------
template <int R, int KQ, bool VEC>
__global__ void __launch_bounds__(SKH_THREADS, 1) skh_iter_kernel(const SkhParams p) {
  constexpr int SEG = SKH_WARPS / R;  // warps cooperating on one row
  static_assert(SEG >= 1 && SEG * R == SKH_WARPS, "R must divide the warp count");
  extern __shared__ __align__(128) unsigned char smem_raw[];

  const int N = p.N, M = p.M;
  const int b = blockIdx.y, g = blockIdx.x, G = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  const int nslab = (N + R - 1) / R;
  const int s_begin = (int)(((long long)nslab * g) / G);
  const int s_end = (int)(((long long)nslab * (g + 1)) / G);
  const int nstage = p.nstage;

  // ---- shared memory carve-up
  const int Mv = (M + 1 + 3) & ~3;            // v2 vector, index M = dustbin column term
  const int stage_floats = ((R * M) + 3) & ~3;
  float* v2_s = reinterpret_cast<float*>(smem_raw);
  float* stage0 = v2_s + Mv;
  float* after = stage0 + (size_t)nstage * stage_floats;
  float2* rowpart = reinterpret_cast<float2*>(after);  // [R][SEG]
  float* u2_s = reinterpret_cast<float*>(rowpart + R * SEG);  // [R]
  float* red_s = u2_s + R;                                      // [2*SKH_WARPS] scratch
  uint64_t* full = reinterpret_cast<uint64_t*>(red_s + 2 * SKH_WARPS + ((R & 1) ? 1 : 0));  // 8-byte aligned

  const float* sc_b = p.scores + (size_t)b * N * M;
  const SkhConst bc = p.bc[b];
  const float zs = p.zscale2;
  const float shift = p.shift ? *p.shift : 0.f;

  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) mbar_init(&full[s], VEC ? 1u : (uint32_t)SKH_THREADS);
    fence_mbar_init();
  }
  __syncthreads();

  // ---- producer: bring slab `s` into stage `st`
  auto issue_slab = [&](int s, int st) {
    const int i0 = s * R;
    const int rows = min(R, N - i0);
    const float* src = sc_b + (size_t)i0 * M;
    float* dst = stage0 + (size_t)st * stage_floats;
    if constexpr (VEC) {
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)rows * (uint32_t)M * 4u;
        fence_proxy_async();
        mbar_arrive_expect_tx(&full[st], bytes);
        tma_bulk_g2s(dst, src, bytes, &full[st]);
      }
    } else {
      const int n = rows * M;
      for (int e = tid; e < n; e += SKH_THREADS) cp_async_4(dst + e, src + e);
      cp_async_mbar_arrive_noinc(&full[st]);
    }
  };
  for (int k = 0; k < nstage; ++k)
    if (s_begin + k < s_end) issue_slab(s_begin + k, k);

  // ---- prologue: column potentials into shared memory (log2 domain), dustbin-row potential
  float uN = 0.f;       // u of the dustbin row for this iteration
  float uN2 = NEG_BIG;  // its log2-domain value used in the dustbin-column partial
  if (!p.dual) {
    const float* v_b = p.v + (size_t)b * p.ldv;
    const float alpha = *p.alpha;
    // LSE over v[0..M] (block reduction), needed for the dustbin row: r_N = alpha + LSE(v)
    float mloc = NEG_BIG;
    for (int j = tid; j <= M; j += SKH_THREADS) mloc = fmaxf(mloc, v_b[j] * LOG2E);
    mloc = warp_max(mloc);
    if (lane == 0) red_s[warp] = mloc;
    __syncthreads();
    float mall = red_s[0];
#pragma unroll
    for (int w = 1; w < SKH_WARPS; ++w) mall = fmaxf(mall, red_s[w]);
    float sloc = 0.f;
    for (int j = tid; j <= M; j += SKH_THREADS) {
      const float vj = v_b[j];
      sloc += ex2(vj * LOG2E - mall);
      float v2;
      if (j < M) {
        v2 = (vj - shift) * LOG2E;
        if (p.apply_mask && !p.tgt_mask[(size_t)b * M + j]) v2 = -INFINITY;
      } else {
        v2 = (alpha + vj) * LOG2E;  // dustbin column entry of every real row: alpha + v_M
      }
      v2_s[j] = v2;
    }
    sloc = warp_sum(sloc);
    if (lane == 0) red_s[SKH_WARPS + warp] = sloc;
    __syncthreads();
    float sall = 0.f;
#pragma unroll
    for (int w = 0; w < SKH_WARPS; ++w) sall += red_s[SKH_WARPS + w];
    const float vlse = (mall + lg2(sall)) * LN2;
    uN = bc.log_mu_bin - (alpha + vlse);
    uN2 = uN * LOG2E;
    if (g == 0 && tid == 0) p.u[(size_t)b * p.ldu + N] = uN;
  } else {
    for (int j = tid; j < M; j += SKH_THREADS)
      v2_s[j] = (p.tgt_mask[(size_t)b * M + j]) ? 0.f : -INFINITY;
    if (tid == 0) v2_s[M] = -INFINITY;
  }
  for (int j = M + 1 + tid; j < Mv; j += SKH_THREADS) v2_s[j] = -INFINITY;
  __syncthreads();

  // ---- per-thread column accumulators (log2 domain)
  float cm[KQ * 4], cs[KQ * 4];
#pragma unroll
  for (int e = 0; e < KQ * 4; ++e) {
    cm[e] = NEG_BIG;
    cs[e] = 0.f;
  }
  LseAcc uacc = lse_empty();  // threads < R: running LSE of the u_i they produced (dustbin column)

  const int seg_len = VEC ? ((((M + SEG - 1) / SEG) + 127) & ~127) : ((((M + SEG - 1) / SEG) + 31) & ~31);

  for (int s = s_begin; s < s_end; ++s) {
    const int it = s - s_begin;
    const int st = it % nstage;
    const uint32_t parity = (uint32_t)((it / nstage) & 1);
    const float* slab = stage0 + (size_t)st * stage_floats;
    const int i0 = s * R;
    const int rows = min(R, N - i0);

    mbar_wait(&full[st], parity);

    // ---- (a) row pass: warp -> (row r, segment seg)
    {
      const int r = warp / SEG, seg = warp % SEG;
      // with the fused mask a padded src row holds only its dustbin entry
      const bool row_live = (r < rows) && !(p.apply_mask && !p.dual && !p.src_mask[(size_t)b * N + i0 + r]);
      if (r < rows && !row_live) {
        if (lane == 0) rowpart[r * SEG + seg] = make_float2(NEG_BIG, 0.f);
      } else if (r < rows) {
        const float* row = slab + (size_t)r * M;
        const int c0 = seg * seg_len;
        const int c1 = min(M, c0 + seg_len);
        float xs[32];
        float m = NEG_BIG;
        if constexpr (VEC) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = c0 + 4 * lane + 128 * k;
            if (c < c1) {
              const float4 z = *reinterpret_cast<const float4*>(row + c);
              const float4 vv = *reinterpret_cast<const float4*>(v2_s + c);
              xs[4 * k + 0] = fmaf(z.x, zs, vv.x);
              xs[4 * k + 1] = fmaf(z.y, zs, vv.y);
              xs[4 * k + 2] = fmaf(z.z, zs, vv.z);
              xs[4 * k + 3] = fmaf(z.w, zs, vv.w);
              m = fmaxf(m, fmaxf(fmaxf(xs[4 * k], xs[4 * k + 1]), fmaxf(xs[4 * k + 2], xs[4 * k + 3])));
            } else {
              xs[4 * k + 0] = xs[4 * k + 1] = xs[4 * k + 2] = xs[4 * k + 3] = -INFINITY;
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int c = c0 + lane + 32 * k;
            if (c < c1) {
              xs[k] = fmaf(row[c], zs, v2_s[c]);
              m = fmaxf(m, xs[k]);
            } else {
              xs[k] = -INFINITY;
            }
          }
        }
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) sum += ex2(xs[k] - m);
        sum = warp_sum(sum);
        if (lane == 0) rowpart[r * SEG + seg] = make_float2(m, sum);
      }
    }
    __syncthreads();

    // ---- (b) one thread per row merges the segments (+ dustbin column entry) -> u_i
    if (tid < R) {
      const int r = tid;
      float u2 = -INFINITY;
      if (r < rows) {
        const int i = i0 + r;
        LseAcc a = lse_empty();
#pragma unroll
        for (int sg = 0; sg < SEG; ++sg) {
          const float2 ps = rowpart[r * SEG + sg];
          lse_merge(a, ps.x, ps.y);
        }
        float ui;
        if (!p.dual) {
          lse_add_value(a, v2_s[M]);  // alpha + v_M
          ui = bc.norm - lse_value(a) * LN2;
        } else {
          ui = -lse_value(a) * LN2;  // -(row log-sum-exp), natural log
        }
        p.u[(size_t)b * p.ldu + i] = ui;
        const bool src_ok = (!p.apply_mask && !p.dual) || p.src_mask[(size_t)b * N + i];
        if (!p.dual) {
          lse_add_value(uacc, ui * LOG2E);
          u2 = src_ok ? (ui - shift) * LOG2E : -INFINITY;  // column pass sees (S - shift) + u
        } else {
          u2 = src_ok ? 0.f : -INFINITY;
        }
      }
      u2_s[r] = u2;
    }
    __syncthreads();

    // ---- (c) column pass: thread -> KQ column quads, all R rows of the slab
    {
      float u2r[R];
#pragma unroll
      for (int r = 0; r < R; ++r) u2r[r] = u2_s[r];  // rows beyond `rows` hold -inf
#pragma unroll
      for (int k = 0; k < KQ; ++k) {
        if constexpr (VEC) {
          const int c = 4 * (tid + SKH_THREADS * k);
          if (c < M) {
            float x[R][4];
            float mx[4] = {NEG_BIG, NEG_BIG, NEG_BIG, NEG_BIG};
#pragma unroll
            for (int r = 0; r < R; ++r) {
              if (r < rows) {
                const float4 z = *reinterpret_cast<const float4*>(slab + (size_t)r * M + c);
                x[r][0] = fmaf(z.x, zs, u2r[r]);
                x[r][1] = fmaf(z.y, zs, u2r[r]);
                x[r][2] = fmaf(z.z, zs, u2r[r]);
                x[r][3] = fmaf(z.w, zs, u2r[r]);
              } else {
                x[r][0] = x[r][1] = x[r][2] = x[r][3] = -INFINITY;
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) mx[e] = fmaxf(mx[e], x[r][e]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float& am = cm[4 * k + e];
              float& as = cs[4 * k + e];
              if (mx[e] > am + 32.f) {  // lazy re-reference: rare after the first slab
                as *= ex2(am - mx[e]);
                am = mx[e];
              }
              float acc = 0.f;
#pragma unroll
              for (int r = 0; r < R; ++r) acc += ex2(x[r][e] - am);
              as += acc;
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = tid + SKH_THREADS * (4 * k + e);
            if (c < M) {
              float x[R];
              float mx = NEG_BIG;
#pragma unroll
              for (int r = 0; r < R; ++r) {
                x[r] = (r < rows) ? fmaf(slab[(size_t)r * M + c], zs, u2r[r]) : -INFINITY;
                mx = fmaxf(mx, x[r]);
              }
              float& am = cm[4 * k + e];
              float& as = cs[4 * k + e];
              if (mx > am + 32.f) {
                as *= ex2(am - mx);
                am = mx;
              }
              float acc = 0.f;
#pragma unroll
              for (int r = 0; r < R; ++r) acc += ex2(x[r] - am);
              as += acc;
            }
          }
        }
      }
    }
    __syncthreads();  // every warp is done with stage `st`
    if (s + nstage < s_end) issue_slab(s + nstage, st);
  }

  // ---- write this CTA's column partials
  float2* cp = p.colpart + ((size_t)b * G + g) * M;
#pragma unroll
  for (int k = 0; k < KQ; ++k) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = VEC ? (4 * (tid + SKH_THREADS * k) + e) : (tid + SKH_THREADS * (4 * k + e));
      if (c < M) cp[c] = make_float2(cm[4 * k + e], cs[4 * k + e]);
    }
  }
  // dustbin-column partial: LSE of this CTA's u_i (threads < R hold disjoint rows)
  if (!p.dual) {
    __syncthreads();
    if (tid < R) rowpart[tid] = make_float2(uacc.m, uacc.s);
    __syncthreads();
    if (tid == 0) {
      LseAcc a = lse_empty();
      for (int r = 0; r < R; ++r) lse_merge(a, rowpart[r].x, rowpart[r].y);
      p.upart[(size_t)b * G + g] = make_float2(a.m, a.s);
    }
  }
}

// ---------------------------------------------------------------------------------------
// one Sinkhorn iteration, software-pipelined (16-byte aligned rows: M % 4 == 0)
//   Same data flow as skh_iter_kernel but ONE block barrier per slab: in loop step k every warp
//   first produces its row-segment partial of slab k+1, then merges the partials of slab k (lanes
//   < R of every warp, redundantly -- a handful of instructions -- so no serial phase and no extra
//   barrier) and runs the column pass of slab k.  The column potentials of the warp's row segment
//   live in registers for the whole kernel (KQ <= 2).  FULL = no bounds checks (M == SEG*128*NCH,
//   and M == 2048*KQ for the column pass).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

template <int R, int KQ, int NCH, bool FULL>
__global__ void __launch_bounds__(SKH_THREADS, 1) skh_iter2_kernel(const SkhParams p) {
  constexpr int SEG = SKH_WARPS / R;  // warps cooperating on one row
  constexpr bool V2REG = (KQ <= 2);
  static_assert(SEG >= 1 && SEG * R == SKH_WARPS, "R must divide the warp count");
  extern __shared__ __align__(128) unsigned char smem_raw[];

  const int N = p.N, M = p.M;
  const int b = blockIdx.y, g = blockIdx.x, G = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  const int nslab = (N + R - 1) / R;
  const int s_begin = (int)(((long long)nslab * g) / G);
  const int s_end = (int)(((long long)nslab * (g + 1)) / G);
  const int nstage = p.nstage;

  // ---- shared memory carve-up
  const int Mv = (M + 1 + 3) & ~3;  // v2 vector, index M = dustbin column term
  const int stage_floats = R * M;   // M % 4 == 0
  float* v2_s = reinterpret_cast<float*>(smem_raw);
  float* stage0 = v2_s + Mv;
  float2* rowpart = reinterpret_cast<float2*>(stage0 + (size_t)nstage * stage_floats);  // [2][R*SEG]
  float* red_s = reinterpret_cast<float*>(rowpart + 2 * SKH_WARPS);                      // [2*SKH_WARPS]
  uint64_t* full = reinterpret_cast<uint64_t*>(red_s + 2 * SKH_WARPS);

  const float* sc_b = p.scores + (size_t)b * N * M;
  const SkhConst bc = p.bc[b];
  const float zs = p.zscale2;
  const float shift = p.shift ? *p.shift : 0.f;

  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) mbar_init(&full[s], 1u);
    fence_mbar_init();
  }
  __syncthreads();

  // ---- producer (thread 0): bring slab `s` into stage `st`
  auto issue_slab = [&](int s, int st) {
    const int i0 = s * R;
    const int rows = min(R, N - i0);
    const float* src = sc_b + (size_t)i0 * M;
    float* dst = stage0 + (size_t)st * stage_floats;
    const uint32_t bytes = (uint32_t)rows * (uint32_t)M * 4u;
    fence_proxy_async();
    mbar_arrive_expect_tx(&full[st], bytes);
    tma_bulk_g2s(dst, src, bytes, &full[st]);
  };
  if (tid == 0)
    for (int k = 0; k < nstage; ++k)
      if (s_begin + k < s_end) issue_slab(s_begin + k, k);

  // ---- prologue: column potentials into shared memory (log2 domain), dustbin-row potential
  float uN = 0.f;
  if (!p.dual) {
    const float* v_b = p.v + (size_t)b * p.ldv;
    const float alpha = *p.alpha;
    float mloc = NEG_BIG;
    for (int j = tid; j <= M; j += SKH_THREADS) mloc = fmaxf(mloc, v_b[j] * LOG2E);
    mloc = warp_max(mloc);
    if (lane == 0) red_s[warp] = mloc;
    __syncthreads();
    float mall = red_s[0];
#pragma unroll
    for (int w = 1; w < SKH_WARPS; ++w) mall = fmaxf(mall, red_s[w]);
    float sloc = 0.f;
    for (int j = tid; j <= M; j += SKH_THREADS) {
      const float vj = v_b[j];
      sloc += ex2(vj * LOG2E - mall);
      float v2;
      if (j < M) {
        v2 = (vj - shift) * LOG2E;
        if (p.apply_mask && !p.tgt_mask[(size_t)b * M + j]) v2 = -INFINITY;
      } else {
        v2 = (alpha + vj) * LOG2E;  // dustbin column entry of every real row: alpha + v_M
      }
      v2_s[j] = v2;
    }
    sloc = warp_sum(sloc);
    if (lane == 0) red_s[SKH_WARPS + warp] = sloc;
    __syncthreads();
    float sall = 0.f;
#pragma unroll
    for (int w = 0; w < SKH_WARPS; ++w) sall += red_s[SKH_WARPS + w];
    const float vlse = (mall + lg2(sall)) * LN2;
    uN = bc.log_mu_bin - (alpha + vlse);
    if (g == 0 && tid == 0) p.u[(size_t)b * p.ldu + N] = uN;
  } else {
    for (int j = tid; j < M; j += SKH_THREADS) v2_s[j] = (p.tgt_mask[(size_t)b * M + j]) ? 0.f : -INFINITY;
    if (tid == 0) v2_s[M] = -INFINITY;
  }
  for (int j = M + 1 + tid; j < Mv; j += SKH_THREADS) v2_s[j] = -INFINITY;
  __syncthreads();

  if constexpr (R == 1) {
    // ---- wide rows (8192 < M <= 16384: one row per stage, two stages).  The software-pipelined loop below keeps TWO rows
    //      resident (row partial of s + 1, column pass of s), i.e. the whole ring, so every row's load started only after
    //      the row before it had been consumed: DRAM latency exposed once per row (measured 9 k cycles per row against the
    //      3 k its 64 KB need at the SM's share of HBM).  Here a thread pulls its KQ quads of the row into registers once
    //      and runs BOTH directions from them (as the persistent kernel does): the stage is free after the first barrier
    //      and is re-armed right there, so two rows are always in flight.
    //      Sinkhorn rows take the SCALED form of the persistent kernel: e_ij = 2^(x_ij + v_j - ref_i) is computed once and
    //      serves the row sum and, divided by it, the column sums (one MUFU per element, one barrier per row).  ref_i comes
    //      from the row potential of the previous iteration (u in HBM; 0 before the first: any finite reference is exact as
    //      long as nothing overflows) and every row CHECKS its sum: outside [2^-60, 2^60] (or NaN) the row is redone from
    //      shared memory with its exact maximum -- no global "the potentials moved little" condition is needed.  Column
    //      sums are carried against the uniform reference 2^(norm2 - shift2 - v2_j) (row-normalised entries are <= 1, so
    //      nothing overflows), which is what the (max, sum) pairs handed to the column merge hold.  Dual-softmax rows keep
    //      the exact two-reduction log-domain pass.
    float cs[KQ * 4];
#pragma unroll
    for (int e = 0; e < KQ * 4; ++e) cs[e] = 0.f;
    float cm[KQ * 4];           // dual mode only: online column maxima
    if (p.dual) {
#pragma unroll
      for (int e = 0; e < KQ * 4; ++e) cm[e] = NEG_BIG;
    }
    LseAcc uacc = lse_empty();  // thread 0: running LSE of the u_i of this CTA's rows (dustbin column)
    float* rmax_s = reinterpret_cast<float*>(rowpart);  // [SKH_WARPS]
    float* rsum_s = rmax_s + SKH_WARPS;                 // [SKH_WARPS]
    float* rsum2_s = rsum_s + SKH_WARPS;                // [SKH_WARPS] (second buffer: consecutive rows alternate)
    const float dust2 = v2_s[M];
    const float norm2 = bc.norm * LOG2E, shift2 = shift * LOG2E;
    auto block_sum16 = [&](const float* buf) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < SKH_WARPS; w += 4) {
        const float4 q = *reinterpret_cast<const float4*>(buf + w);
        t += (q.x + q.y) + (q.z + q.w);
      }
      return t;
    };
    for (int s = s_begin; s < s_end; ++s) {
      const int it = s - s_begin;
      const int st = it % nstage;
      const int i = s;
      const bool row_live = !(p.apply_mask && !p.dual && !p.src_mask[(size_t)b * N + i]);
      const bool src_ok = (!p.apply_mask && !p.dual) || p.src_mask[(size_t)b * N + i];
      float mh = 0.f;
      if (!p.dual) mh = (bc.norm - __ldcg(p.u + (size_t)b * p.ldu + i)) * LOG2E;  // row reference: last iteration's row log-sum-exp
      mbar_wait(&full[st], (uint32_t)((it / nstage) & 1));
      const float* row = stage0 + (size_t)st * stage_floats;
      float4 z[KQ];
      float ui, u2;
      if (!p.dual) {
        float* rs_buf = (it & 1) ? rsum2_s : rsum_s;
        float rs = 0.f;
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          const int c = 4 * (tid + SKH_THREADS * k);
          if ((FULL || c < M) && row_live) {
            const float4 zz = *reinterpret_cast<const float4*>(row + c);
            const float4 vv = *reinterpret_cast<const float4*>(v2_s + c);
            float4 e;
            e.x = ex2(fmaf(zz.x, zs, vv.x) - mh);
            e.y = ex2(fmaf(zz.y, zs, vv.y) - mh);
            e.z = ex2(fmaf(zz.z, zs, vv.z) - mh);
            e.w = ex2(fmaf(zz.w, zs, vv.w) - mh);
            z[k] = e;
            rs += (e.x + e.y) + (e.z + e.w);
          } else {
            z[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        rs = warp_sum(rs);
        if (lane == 0) rs_buf[warp] = rs;
        __syncthreads();  // every thread holds its part of the row
        float tot = block_sum16(rs_buf);
        const float edust = ex2(dust2 - mh);
        const bool healthy = !row_live || (tot + edust > 8.6736174e-19f && tot + edust < 1.1529215e18f);  // uniform over the CTA
        if (!healthy) {
          // rare (a reference far from the row's scale): exact row maximum from the row still in shared memory
          float m = NEG_BIG;
#pragma unroll
          for (int k = 0; k < KQ; ++k) {
            const int c = 4 * (tid + SKH_THREADS * k);
            if (FULL || c < M) {
              const float4 zz = *reinterpret_cast<const float4*>(row + c);
              const float4 vv = *reinterpret_cast<const float4*>(v2_s + c);
              m = fmaxf(m, fmaxf(fmaxf(fmaf(zz.x, zs, vv.x), fmaf(zz.y, zs, vv.y)), fmaxf(fmaf(zz.z, zs, vv.z), fmaf(zz.w, zs, vv.w))));
            }
          }
          m = warp_max(m);
          if (lane == 0) rmax_s[warp] = m;
          __syncthreads();
          float mrow = dust2;
#pragma unroll
          for (int w = 0; w < SKH_WARPS; ++w) mrow = fmaxf(mrow, rmax_s[w]);
          mh = mrow;
          rs = 0.f;
#pragma unroll
          for (int k = 0; k < KQ; ++k) {
            const int c = 4 * (tid + SKH_THREADS * k);
            if (FULL || c < M) {
              const float4 zz = *reinterpret_cast<const float4*>(row + c);
              const float4 vv = *reinterpret_cast<const float4*>(v2_s + c);
              float4 e;
              e.x = ex2(fmaf(zz.x, zs, vv.x) - mh);
              e.y = ex2(fmaf(zz.y, zs, vv.y) - mh);
              e.z = ex2(fmaf(zz.z, zs, vv.z) - mh);
              e.w = ex2(fmaf(zz.w, zs, vv.w) - mh);
              z[k] = e;
              rs += (e.x + e.y) + (e.z + e.w);
            }
          }
          rs = warp_sum(rs);
          __syncthreads();          // rmax_s / the other sum buffer have been read
          float* rs_buf2 = (it & 1) ? rsum_s : rsum2_s;
          if (lane == 0) rs_buf2[warp] = rs;
          __syncthreads();
          tot = block_sum16(rs_buf2);
          __syncthreads();          // ... before the next row writes that buffer
        }
        if (tid == 0 && s + nstage < s_end) issue_slab(s + nstage, st);  // the stage is free: re-arm it
        const float srow = tot + ex2(dust2 - mh);       // + the dustbin column entry (alpha + v_M)
        const float rowlse2 = row_live ? mh + lg2(srow) : dust2;  // a padded row (fused mask) holds its dustbin entry only
        ui = bc.norm - rowlse2 * LN2;
        u2 = src_ok ? (ui - shift) * LOG2E : -INFINITY;
        const float w = (row_live && src_ok) ? 1.f / srow : 0.f;
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          cs[4 * k + 0] = fmaf(z[k].x, w, cs[4 * k + 0]);
          cs[4 * k + 1] = fmaf(z[k].y, w, cs[4 * k + 1]);
          cs[4 * k + 2] = fmaf(z[k].z, w, cs[4 * k + 2]);
          cs[4 * k + 3] = fmaf(z[k].w, w, cs[4 * k + 3]);
        }
      } else {
        // dual softmax: exact log-domain pass (row maximum, sum, online column maxima)
        float m = NEG_BIG;
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          const int c = 4 * (tid + SKH_THREADS * k);
          if (FULL || c < M) {
            z[k] = *reinterpret_cast<const float4*>(row + c);
            const float4 vv = *reinterpret_cast<const float4*>(v2_s + c);
            m = fmaxf(m, fmaxf(fmaxf(fmaf(z[k].x, zs, vv.x), fmaf(z[k].y, zs, vv.y)), fmaxf(fmaf(z[k].z, zs, vv.z), fmaf(z[k].w, zs, vv.w))));
          } else {
            z[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        m = warp_max(m);
        if (lane == 0) rmax_s[warp] = m;
        __syncthreads();  // every thread holds its part of the row: the stage is free
        if (tid == 0 && s + nstage < s_end) issue_slab(s + nstage, st);
        float mrow = NEG_BIG;
#pragma unroll
        for (int w = 0; w < SKH_WARPS; ++w) mrow = fmaxf(mrow, rmax_s[w]);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          const int c = 4 * (tid + SKH_THREADS * k);
          if (FULL || c < M) {
            const float4 vv = *reinterpret_cast<const float4*>(v2_s + c);
            sum += (ex2(fmaf(z[k].x, zs, vv.x) - mrow) + ex2(fmaf(z[k].y, zs, vv.y) - mrow)) +
                   (ex2(fmaf(z[k].z, zs, vv.z) - mrow) + ex2(fmaf(z[k].w, zs, vv.w) - mrow));
          }
        }
        sum = warp_sum(sum);
        if (lane == 0) rsum_s[warp] = sum;
        __syncthreads();
        const float tot = block_sum16(rsum_s);
        ui = -(mrow + lg2(tot)) * LN2;  // -(row log-sum-exp), natural log
        u2 = src_ok ? 0.f : -INFINITY;
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          const int c = 4 * (tid + SKH_THREADS * k);
          if (FULL || c < M) {
            const float y[4] = {fmaf(z[k].x, zs, u2), fmaf(z[k].y, zs, u2), fmaf(z[k].z, zs, u2), fmaf(z[k].w, zs, u2)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float& am = cm[4 * k + e];
              float& as = cs[4 * k + e];
              if (y[e] > am + 32.f) {  // lazy re-reference: rare after the first rows
                as *= ex2(am - y[e]);
                am = y[e];
              }
              as += ex2(y[e] - am);
            }
          }
        }
      }
      if (tid == 0) {
        p.u[(size_t)b * p.ldu + i] = ui;
        if (!p.dual) lse_add_value(uacc, ui * LOG2E);
      }
    }
    float2* cp = p.colpart + ((size_t)b * G + g) * M;
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
      const int c = 4 * (tid + SKH_THREADS * k);
      if (FULL || c < M) {
        float ref[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (p.dual) {
            ref[e] = cm[4 * k + e];
          } else {  // sum_i 2^(x_ij + u_i log2e) = 2^(norm2 - shift2 - v2_j) * sum_i e_ij / srow_i
            const float v2j = v2_s[c + e];
            const bool okc = v2j > -INFINITY;
            ref[e] = okc ? (norm2 - v2j - shift2) : NEG_BIG;
            if (!okc) cs[4 * k + e] = 0.f;
          }
        }
        *reinterpret_cast<float4*>(cp + c) = make_float4(ref[0], cs[4 * k], ref[1], cs[4 * k + 1]);
        *reinterpret_cast<float4*>(cp + c + 2) = make_float4(ref[2], cs[4 * k + 2], ref[3], cs[4 * k + 3]);
      }
    }
    if (!p.dual && tid == 0) p.upart[(size_t)b * G + g] = make_float2(uacc.m, uacc.s);
    return;
  }

  // ---- this warp's row segment
  const int wr = warp / SEG, wseg = warp % SEG;
  const int seg_len = FULL ? NCH * 128 : ((((M + SEG - 1) / SEG) + 127) & ~127);
  const int c0 = wseg * seg_len;
  const int c1 = min(M, c0 + seg_len);
  const float dust2 = v2_s[M];
  float4 v2r[V2REG ? NCH : 1];
  if constexpr (V2REG) {
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = c0 + 4 * lane + 128 * k;
      v2r[k] = (FULL || c < c1) ? *reinterpret_cast<const float4*>(v2_s + c) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
  }

  // ---- per-thread column accumulators (log2 domain)
  float cm[KQ * 4], cs[KQ * 4];
#pragma unroll
  for (int e = 0; e < KQ * 4; ++e) {
    cm[e] = NEG_BIG;
    cs[e] = 0.f;
  }
  LseAcc uacc = lse_empty();  // warp 0, lanes < R: running LSE of the u_i they produced (dustbin column)

  // ---- (a) row-segment partial of slab s (stage st) -> rowpart[buf]
  auto row_partial = [&](int s, int st, int buf) {
    const int i0 = s * R;
    const int rows = min(R, N - i0);
    if (wr >= rows) return;
    const bool row_live = !(p.apply_mask && !p.dual && !p.src_mask[(size_t)b * N + i0 + wr]);
    if (!row_live) {  // with the fused mask a padded src row holds only its dustbin entry
      if (lane == 0) rowpart[buf * SKH_WARPS + warp] = make_float2(NEG_BIG, 0.f);
      return;
    }
    const float* row = stage0 + (size_t)st * stage_floats + (size_t)wr * M;
    float xs[4 * NCH];
    float m = NEG_BIG;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = c0 + 4 * lane + 128 * k;
      if (FULL || c < c1) {
        const float4 z = *reinterpret_cast<const float4*>(row + c);
        float4 vv;
        if constexpr (V2REG)
          vv = v2r[k];
        else
          vv = *reinterpret_cast<const float4*>(v2_s + c);
        xs[4 * k + 0] = fmaf(z.x, zs, vv.x);
        xs[4 * k + 1] = fmaf(z.y, zs, vv.y);
        xs[4 * k + 2] = fmaf(z.z, zs, vv.z);
        xs[4 * k + 3] = fmaf(z.w, zs, vv.w);
        m = fmaxf(m, fmaxf(fmaxf(xs[4 * k], xs[4 * k + 1]), fmaxf(xs[4 * k + 2], xs[4 * k + 3])));
      } else {
        xs[4 * k + 0] = xs[4 * k + 1] = xs[4 * k + 2] = xs[4 * k + 3] = -INFINITY;
      }
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 4 * NCH; ++k) sum += ex2(xs[k] - m);
    sum = warp_sum(sum);
    if (lane == 0) rowpart[buf * SKH_WARPS + warp] = make_float2(m, sum);
  };

  // ---- (b)+(c) merge the partials of slab s into u_i, then the column pass over the slab
  auto col_pass = [&](int s, int st, int buf) {
    const int i0 = s * R;
    const int rows = min(R, N - i0);
    const float* slab = stage0 + (size_t)st * stage_floats;
    float u2 = -INFINITY;
    if (lane < R && lane < rows) {
      const int r = lane, i = i0 + r;
      LseAcc a = lse_empty();
#pragma unroll
      for (int sg = 0; sg < SEG; ++sg) {
        const float2 ps = rowpart[buf * SKH_WARPS + r * SEG + sg];
        lse_merge(a, ps.x, ps.y);
      }
      float ui;
      if (!p.dual) {
        lse_add_value(a, dust2);  // alpha + v_M
        ui = bc.norm - lse_value(a) * LN2;
      } else {
        ui = -lse_value(a) * LN2;  // -(row log-sum-exp), natural log
      }
      const bool src_ok = (!p.apply_mask && !p.dual) || p.src_mask[(size_t)b * N + i];
      if (warp == 0) {
        p.u[(size_t)b * p.ldu + i] = ui;
        if (!p.dual) lse_add_value(uacc, ui * LOG2E);
      }
      if (!p.dual)
        u2 = src_ok ? (ui - shift) * LOG2E : -INFINITY;  // column pass sees (S - shift) + u
      else
        u2 = src_ok ? 0.f : -INFINITY;
    }
    float u2r[R];
#pragma unroll
    for (int r = 0; r < R; ++r) u2r[r] = __shfl_sync(0xffffffffu, u2, r);  // rows beyond `rows` hold -inf
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
      const int c = 4 * (tid + SKH_THREADS * k);
      if (FULL || c < M) {
        float x[R][4];
        float mx[4] = {NEG_BIG, NEG_BIG, NEG_BIG, NEG_BIG};
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r < rows) {
            const float4 z = *reinterpret_cast<const float4*>(slab + (size_t)r * M + c);
            x[r][0] = fmaf(z.x, zs, u2r[r]);
            x[r][1] = fmaf(z.y, zs, u2r[r]);
            x[r][2] = fmaf(z.z, zs, u2r[r]);
            x[r][3] = fmaf(z.w, zs, u2r[r]);
          } else {
            x[r][0] = x[r][1] = x[r][2] = x[r][3] = -INFINITY;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) mx[e] = fmaxf(mx[e], x[r][e]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float& am = cm[4 * k + e];
          float& as = cs[4 * k + e];
          if (mx[e] > am + 32.f) {  // lazy re-reference: rare after the first slab
            as *= ex2(am - mx[e]);
            am = mx[e];
          }
          float acc = 0.f;
#pragma unroll
          for (int r = 0; r < R; ++r) acc += ex2(x[r][e] - am);
          as += acc;
        }
      }
    }
  };

  // ---- main loop
  if (s_begin < s_end) {
    mbar_wait(&full[0], 0u);
    row_partial(s_begin, 0, 0);
  }
  __syncthreads();
  for (int s = s_begin; s < s_end; ++s) {
    const int it = s - s_begin;
    const int st = it % nstage;
    const int buf = it & 1;
    if (s + 1 < s_end) {
      const int it1 = it + 1;
      mbar_wait(&full[it1 % nstage], (uint32_t)((it1 / nstage) & 1));
      row_partial(s + 1, it1 % nstage, buf ^ 1);
    }
    col_pass(s, st, buf);
    __syncthreads();  // every warp is done with stage `st`, and rowpart[buf ^ 1] is complete
    if (tid == 0 && s + nstage < s_end) issue_slab(s + nstage, st);
  }

  // ---- write this CTA's column partials
  float2* cp = p.colpart + ((size_t)b * G + g) * M;
#pragma unroll
  for (int k = 0; k < KQ; ++k) {
    const int c = 4 * (tid + SKH_THREADS * k);
    if (FULL || c < M) {
      *reinterpret_cast<float4*>(cp + c) = make_float4(cm[4 * k], cs[4 * k], cm[4 * k + 1], cs[4 * k + 1]);
      *reinterpret_cast<float4*>(cp + c + 2) = make_float4(cm[4 * k + 2], cs[4 * k + 2], cm[4 * k + 3], cs[4 * k + 3]);
    }
  }
  // dustbin-column partial: LSE of this CTA's u_i (warp 0, lanes < R hold disjoint rows)
  if (!p.dual && warp == 0) {
    float am = uacc.m, as = uacc.s;
#pragma unroll
    for (int o = 1; o < R; o <<= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, am, o);
      const float s2 = __shfl_xor_sync(0xffffffffu, as, o);
      LseAcc a{am, as};
      lse_merge(a, m2, s2);
      am = a.m;
      as = a.s;
    }
    if (lane == 0) p.upart[(size_t)b * G + g] = make_float2(am, as);
  }
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Philox4x32-7 counter-based generator (Salmon et al., SC'11: 7 rounds is the fewest that passes BigCrush; 10 is the
// library default with extra margin) + Box-Muller: four N(0,1) draws per counter.
// Counter = (quad index of the element, noise_offset); key = noise_seed.  Replaces torch.randn_like(x)
// (Diff-Reg-4dmatch/models/pipeline.py:188) in throughput mode; parity tests pass the noise tensor instead.
__device__ __forceinline__ uint4 philox4x32_7(uint4 c, uint2 k) {
  const unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const unsigned long long p0 = (unsigned long long)M0 * c.x, p1 = (unsigned long long)M1 * c.z;  // one IMAD.WIDE each
    const unsigned int hi0 = (unsigned int)(p0 >> 32), lo0 = (unsigned int)p0;
    const unsigned int hi1 = (unsigned int)(p1 >> 32), lo1 = (unsigned int)p1;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
// the same generator with the round keys precomputed (kernel parameters: constant-bank operands, no key schedule in the loop)
__device__ __forceinline__ uint4 philox4x32_7_rk(uint4 c, const unsigned int (&rk)[14]) {
  const unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const unsigned long long p0 = (unsigned long long)M0 * c.x, p1 = (unsigned long long)M1 * c.z;
    const unsigned int hi0 = (unsigned int)(p0 >> 32), lo0 = (unsigned int)p0;
    const unsigned int hi1 = (unsigned int)(p1 >> 32), lo1 = (unsigned int)p1;
    c = make_uint4(hi1 ^ c.y ^ rk[2 * r], lo1, hi0 ^ c.w ^ rk[2 * r + 1], lo0);
  }
  return c;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_ftz(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 philox_normal4(unsigned long long quad, unsigned long long offset, unsigned long long seed) {
  const uint4 r = philox4x32_7(make_uint4((unsigned int)quad, (unsigned int)(quad >> 32), (unsigned int)offset,
                                           (unsigned int)(offset >> 32)),
                                make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = fmaf((float)r.x, k, 0.5f * k);   // in [2^-33, 1]: never denormal
  const float u2 = fmaf((float)r.z, k, 0.5f * k);
  // The pass is issue-bound, so the flush-to-zero forms are used: the default lg2 / rsqrt carry a denormal guard
  // (compare, two predicated multiplies / adds each) that these arguments can never need.
  // sqrt(x) = x * rsqrt(x): two instructions, ~1 ulp -- irrelevant for noise draws
  const float xa = -1.3862943611198906f * lg2_ftz(u0), xb = -1.3862943611198906f * lg2_ftz(u2);  // -2 ln u
  const float ra = xa * rsqrt_ftz(fmaxf(xa, 1e-30f)), rb = xb * rsqrt_ftz(fmaxf(xb, 1e-30f));
  float sa, ca, sb, cb;
  const float k2pi = 1.4629180792671596e-09f;  // 2 pi 2^-32: the angle in one multiply
  __sincosf((float)r.y * k2pi, &sa, &ca);
  __sincosf((float)r.w * k2pi, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

// philox_normal4 with precomputed round keys and the counter words given separately (identical draws)
__device__ __forceinline__ float4 philox_normal4_rk(unsigned int quad_lo, unsigned int quad_hi, unsigned int off_lo, unsigned int off_hi,
                                                    const unsigned int (&rk)[14]) {
  const uint4 r = philox4x32_7_rk(make_uint4(quad_lo, quad_hi, off_lo, off_hi), rk);
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = fmaf((float)r.x, k, 0.5f * k);
  const float u2 = fmaf((float)r.z, k, 0.5f * k);
  const float xa = -1.3862943611198906f * lg2_ftz(u0), xb = -1.3862943611198906f * lg2_ftz(u2);  // -2 ln u
  const float ra = xa * rsqrt_ftz(fmaxf(xa, 1e-30f)), rb = xb * rsqrt_ftz(fmaxf(xb, 1e-30f));
  float sa, ca, sb, cb;
  const float k2pi = 1.4629180792671596e-09f;  // 2 pi 2^-32
  __sincosf((float)r.y * k2pi, &sa, &ca);
  __sincosf((float)r.w * k2pi, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

__device__ __forceinline__ void atomic_min_float(float* addr, float value) {
  if (value >= 0.f)
    atomicMin(reinterpret_cast<int*>(addr), __float_as_int(value));
  else
    atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(value));
}

// ---------------------------------------------------------------------------------------
// Register-slab persistent Sinkhorn (M % 4 == 0, M <= 4096, G*B <= #SMs)
//   skh_persist_kernel above hands every element from the row warps to the column warps through shared memory and
//   is bound by shared-memory bandwidth (TMA write + row read + potentials read + e write + column read: 16-20 bytes
//   per element against 128 B/clk/SM).  Here shared memory is only the landing zone of the TMA ring (4 B written +
//   4 B read per element): a CTA is two row groups of 256 threads, a thread owns KQ column quads (columns
//   4*(ct + 256k) .. +3) and pulls RR rows x KQ quads of a mini-slab into registers, where BOTH directions are
//   computed.  Row log-sum-exps are reduced across the 256 threads of the group (warp shuffles + 8 partials through
//   shared memory + ONE named barrier per mini-slab); the column log-sum-exp partials stay in the thread's registers
//   for the whole pass.  The ring holds up to ~190 KB of loads in flight per SM (deep enough for the L2 / HBM
//   latency; loading straight into registers one mini-slab ahead was latency-bound at 5.6 TB/s); a stage is re-armed
//   by the group that consumed it right after its barrier, so no "empty" barriers are needed.
//   Scaled pass (iterations >= 1, see skh_persist_kernel): e_ij = 2^(x_ij - ref_i) is computed once and used for both
//   directions -- 1 MUFU, 2 FFMA, 2 FADD per element.
//   Iteration structure (prologue, grid barrier, merge of the per-CTA column partials, grid barrier) as above.
// ---------------------------------------------------------------------------------------
constexpr int P2_GROUPS = 2;           // row groups per CTA (each consumes every P2_GROUPS-th mini-slab); 3 measured slower (82 vs 73 us)
constexpr int P2_TPR = 256;            // threads per row group
constexpr int P2_THREADS = P2_GROUPS * P2_TPR;
constexpr int P2_GW = P2_TPR / 32;     // warps per row group
constexpr int P2_MAX_STAGES = 8;
// candidate-search tail: sampled rows per column, and the histogram of the samples' log2 confidences
// (SH_PER_OCTAVE bins per octave over [2^-SH_OFFSET, 2^(TK_BINS / SH_PER_OCTAVE - SH_OFFSET)); bin 0 also takes everything below)
constexpr int SH_ROWS = P2_THREADS / 32;
constexpr int SH_PER_OCTAVE = 16;
constexpr float SH_OFFSET = 120.f;

__device__ __forceinline__ float4 ldg_stream4(const float* ptr) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(ptr));
  return r;
}
__device__ __forceinline__ void group_barrier(int rg) { asm volatile("bar.sync %0, %1;" ::"r"(rg + 1), "r"(P2_TPR) : "memory"); }

// grid barrier with a release add / acquire poll by one thread per CTA (the block barriers make it cumulative)
__device__ __forceinline__ void grid_barrier_ra(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    while (ld_acquire_u32(counter) < target) {
    }
  }
  __syncthreads();
}

template <int KQ, int RR, bool FULL>
__global__ void __launch_bounds__(P2_THREADS, 1) skh_persist2_kernel(const SkhParams p, const int iters, unsigned int* gsync) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int N = p.N, M = p.M;
  const int b = blockIdx.y, g = blockIdx.x, G = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rg = tid >> 8, ct = tid & (P2_TPR - 1), wg = warp & (P2_GW - 1);
  const int row0 = (int)(((long long)N * g) / G), row1 = (int)(((long long)N * (g + 1)) / G);
  const int nrows = row1 - row0;              // <= 1024 (host check)
  const int ns = (nrows + RR - 1) / RR;       // mini-slabs of this CTA; row group rg takes s = rg, rg + 2, ...
  if (p.zero_a) {
    const size_t nthr = (size_t)gridDim.x * gridDim.y * P2_THREADS;
    const size_t me = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * P2_THREADS + tid;
    for (size_t i = me; i < p.zero_a_n; i += nthr) p.zero_a[i] = 0ull;
    for (size_t i = me; i < p.zero_b_n; i += nthr) p.zero_b[i] = 0ull;
  }

  const bool do_collect = p.col.state != nullptr;
  if (do_collect) {  // ordered before their first use (the last iteration's merge) by the grid barriers in between
    for (int q = g * P2_THREADS + tid; q < TK_BINS; q += G * P2_THREADS) {
      p.col.sample_hist[(size_t)b * TK_BINS + q] = 0u;
      p.col.sample_hist2[(size_t)b * TK_BINS + q] = 0u;
      p.col.cand_hist[(size_t)b * TK_BINS + q] = 0u;
    }
    if (g == 0 && tid == 0) {
      p.col.state[b].n_cand = 0u;
      p.col.state[b].seg_broken = 0u;
      p.col.sample_list_n[b] = 0u;
    }
  }

  const int Mv = (M + 1 + 3) & ~3;
  float* v2_s = reinterpret_cast<float*>(smem_raw);             // [Mv]  column potentials, log2 domain, minus the shift
  float* lse_prev_s = v2_s + Mv;                                // [1024] row log-sum-exp (log2) of the previous iteration
  float* red_part = lse_prev_s + 1024;                          // [2 buffers][P2_GROUPS][8 rows][8 warps]
  float* red_s = red_part + 2 * P2_GROUPS * 64;                                // [64]
  float2* upart_s = reinterpret_cast<float2*>(red_s + 64);      // [2] (+2 pad)
  float2* comb = upart_s + 4;                                   // [16 warps][32] merge scratch
  unsigned int* tail_s = reinterpret_cast<unsigned int*>(comb + (P2_THREADS / 32) * 32);  // [TK_BINS] candidate-search tail: sample histogram, then the CTA's candidate list
  uint64_t* full = reinterpret_cast<uint64_t*>(tail_s + TK_BINS);  // [P2_MAX_STAGES]
  float* ring = reinterpret_cast<float*>(full + P2_MAX_STAGES);  // [D][RR * M] TMA landing zone, shared by the two row groups
  float2* xcomb = reinterpret_cast<float2*>(ring);              // [KQ*4][256] column partials of row group 1 (pass is over: ring idle)
  const int D = p.nstage;
  const int stage_floats = RR * M;
  __shared__ int cnt_s[2], cnt2_s[2];

#define DRG_STAMP(slot) do { if (p.dbg_times && g == 0 && b == 0) p.dbg_times[(slot)] = clock64(); } while (0)
  const float* sc_b = p.scores + (size_t)b * N * M;
  const float zs = p.zscale2;
  const float shift = p.shift ? *p.shift : 0.f;
  const float shift2 = shift * LOG2E;
  const float alpha = *p.alpha;
  unsigned int* gcount = gsync + b;
  unsigned int* dv_slots = gsync + gridDim.y + 2 * b;

  // stream of mini-slabs: element T = it * ns + s lands in stage T % D; row group (s & 1) consumes it
  auto issue_slab_to = [&](int stg, int s_) {
    const int i0 = row0 + s_ * RR;
    const uint32_t bytes = (uint32_t)min(RR, row1 - i0) * (uint32_t)M * 4u;
    fence_proxy_async();
    mbar_arrive_expect_tx(&full[stg], bytes);
    tma_bulk_g2s(ring + (size_t)stg * stage_floats, sc_b + (size_t)i0 * M, bytes, &full[stg]);
  };
  // final phase (fin.out != NULL): its stream interleaves the score slab and (DDIM) the x_t slab of every mini-slab: element
  // q = NWf * s + w (w = 0 scores, 1 x_t) lands in stage q % Df (Df even in DDIM mode, so that a mini-slab's two stages are
  // re-armed for the same later mini-slab)
  const bool do_final = p.fin.out != nullptr;
  const bool f_ddim = do_final && p.fin.ddim != 0;
  const int NWf = f_ddim ? 2 : 1;
  const int Df = f_ddim ? (D & ~1) : D;
  const float* xt_b = f_ddim ? p.fin.x_t + (size_t)b * N * M : nullptr;
  auto issue_final = [&](int q) {
    const int s_ = q / NWf, wsrc = q - s_ * NWf;
    const int i0 = row0 + s_ * RR;
    const int stg_ = q % Df;
    const uint32_t bytes = (uint32_t)min(RR, row1 - i0) * (uint32_t)M * 4u;
    fence_proxy_async();
    mbar_arrive_expect_tx(&full[stg_], bytes);
    tma_bulk_g2s(ring + (size_t)stg_ * stage_floats, (wsrc ? xt_b : sc_b) + (size_t)i0 * M, bytes, &full[stg_]);
  };
  if (tid == 0) {
    DRG_STAMP(0);
    if (do_collect) DRG_STAMP(704);
    cnt_s[0] = cnt_s[1] = 0;
    for (int s = 0; s < D; ++s) mbar_init(&full[s], 1u);
    fence_mbar_init();
    if (iters > 0)
      for (int s = 0; s < D && s < ns; ++s) issue_slab_to(s, s);
  }
  __syncthreads();
  {
    int cs = count_mask_bytes(p.src_mask + (size_t)b * N, N, tid, P2_THREADS);
    int ctg = count_mask_bytes(p.tgt_mask + (size_t)b * M, M, tid, P2_THREADS);
    cs = __reduce_add_sync(0xffffffffu, cs);
    ctg = __reduce_add_sync(0xffffffffu, ctg);
    if (lane == 0) {
      if (cs) atomicAdd(&cnt_s[0], cs);
      if (ctg) atomicAdd(&cnt_s[1], ctg);
    }
  }
  __syncthreads();
  SkhConst bc;
  {
    const float ms = (float)cnt_s[0], nsv = (float)cnt_s[1];
    bc.norm = -logf(ms + nsv);
    bc.log_mu_bin = logf(nsv) + bc.norm;
    bc.log_nu_bin = logf(ms) + bc.norm;
    bc.pad = (cnt_s[0] == N && cnt_s[1] == M) ? 1.f : 0.f;  // 1: no padded row or column
    if (g == 0 && tid == 0) p.bc_out[b] = bc;
  }
  const float norm2 = bc.norm * LOG2E;
  // K_b of the candidate search: the mean over the batch of the per-element caps (procrustes.py:61-65), from the masks
  int Kb = 0;
  if (do_collect) {
    float cap_sum = 0.f;
    int my_cap = 0;
    for (int bb = 0; bb < (int)gridDim.y; ++bb) {
      int nsrc = N, ntgt = M;
      if (!p.col.padded_lengths) {
        if (bb == b) {
          nsrc = cnt_s[0];
          ntgt = cnt_s[1];
        } else {
          __syncthreads();
          if (tid == 0) cnt2_s[0] = cnt2_s[1] = 0;
          __syncthreads();
          int cs = count_mask_bytes(p.src_mask + (size_t)bb * N, N, tid, P2_THREADS);
          int ctg = count_mask_bytes(p.tgt_mask + (size_t)bb * M, M, tid, P2_THREADS);
          cs = __reduce_add_sync(0xffffffffu, cs);
          ctg = __reduce_add_sync(0xffffffffu, ctg);
          if (lane == 0) {
            if (cs) atomicAdd(&cnt2_s[0], cs);
            if (ctg) atomicAdd(&cnt2_s[1], ctg);
          }
          __syncthreads();
          nsrc = cnt2_s[0];
          ntgt = cnt2_s[1];
        }
      }
      const int cap = (int)((float)max(nsrc, ntgt) * p.col.sample_rate);  // (max(len) * sample_rate).int(), fp32 product
      cap_sum += (float)cap;
      if (bb == b) my_cap = cap;
    }
    Kb = min(min((int)(cap_sum / (float)gridDim.y), my_cap), p.col.K_max);
  }
  unsigned int barriers_done = 0;
  const uint8_t* smask = p.src_mask + (size_t)b * N;
  // row slots: log2-domain row log-sum-exp of the previous iteration; >= 1e30 marks a row without real entries
  // (masked source row: 1e30, written by the log-domain pass; slot past the CTA's rows: 2e30)
  for (int q = nrows + tid; q < 1024; q += P2_THREADS) lse_prev_s[q] = 2.0e30f;

  if (tid == 0) DRG_STAMP(1);
  for (int it = 0; it < iters; ++it) {
    if (tid == 0) DRG_STAMP(10 + it * 100 + 0);
    // how far the column potentials moved in the last merge decides between the scaled pass and the log-domain one.
    // A plain L2 load (the grid barrier already ordered the merge's atomicMax before us): an acquire load here would
    // hold back the loads of v below by a full round trip.
    unsigned int dv_bits = 0u;
    if (it >= 1) dv_bits = __ldcg(dv_slots + ((it - 1) & 1));
    // First iteration: same arithmetic, but the row reference is this pass's exact row maximum (one more reduction per
    // mini-slab).  Flushing e_ij / srow_i < 2^-126 cannot hurt the column sums there: with v = 0 the dustbin-row entry of
    // every column is ~2^norm2, 2^126 / N times larger than anything that can be flushed.
    const bool semi = (it == 0) && alpha >= -20.f;

    // ---- prologue: column potentials into shared memory (log2 domain).  The dustbin-row potential u_N needs the
    //      log-sum-exp of ALL of v; nothing in the pass uses it, so only per-warp (max, sum) partials are formed here (no
    //      block-wide reduction, no barrier) and they are combined after the pass, where the CTA synchronises anyway.
    if (it == 0) {
      for (int j = tid; j <= M; j += P2_THREADS) {
        float v2;
        if (j < M) {
          v2 = -shift2;
          if (p.apply_mask && !p.tgt_mask[(size_t)b * M + j]) v2 = -INFINITY;
        } else {
          v2 = alpha * LOG2E;
        }
        v2_s[j] = v2;
      }
    } else {
      const float* v_b = p.v + (size_t)b * p.ldv;
      constexpr int VPT = (4096 + 1 + P2_THREADS - 1) / P2_THREADS;  // M <= 4096
      float vr[VPT];
      float mloc = NEG_BIG;
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        const int j = tid + k * P2_THREADS;
        vr[k] = (j <= M) ? __ldcg(v_b + j) : -INFINITY;
        mloc = fmaxf(mloc, vr[k] * LOG2E);
      }
      float sloc = 0.f;
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        const int j = tid + k * P2_THREADS;
        if (j <= M) {
          const float vj = vr[k];
          sloc += ex2(vj * LOG2E - mloc);
          float v2;
          if (j < M) {
            v2 = (vj - shift) * LOG2E;
            if (p.apply_mask && !p.tgt_mask[(size_t)b * M + j]) v2 = -INFINITY;
          } else {
            v2 = (alpha + vj) * LOG2E;
          }
          v2_s[j] = v2;
        }
      }
      // warp-level log-sum-exp of the (max, sum) pairs
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, mloc, o);
        const float s2 = __shfl_xor_sync(0xffffffffu, sloc, o);
        const float mn = fmaxf(mloc, m2);
        sloc = sloc * ex2(mloc - mn) + s2 * ex2(m2 - mn);
        mloc = mn;
      }
      if (lane == 0) {
        red_s[warp] = mloc;
        red_s[32 + warp] = sloc;
      }
    }
    for (int j = M + 1 + tid; j < Mv; j += P2_THREADS) v2_s[j] = -INFINITY;
    const bool fast = it >= 1 && alpha >= -20.f && __uint_as_float(dv_bits) <= 50.f;
    const bool scaled = fast || semi;
    if (p.dbg_times && g == 0 && b == 0 && tid == 0) p.dbg_times[400 + it] = fast ? 1 : semi ? 2 : 0;
    __syncthreads();
    if (tid == 0) DRG_STAMP(10 + it * 100 + 1);

    // ---- the pass over this CTA's rows
    const float dust2 = v2_s[M];
    float cs[KQ * 4], cm[KQ * 4];
#pragma unroll
    for (int e = 0; e < KQ * 4; ++e) {
      cs[e] = 0.f;
      cm[e] = NEG_BIG;
    }
    LseAcc uacc = lse_empty();
    int buf = 0;
    int stg = (it * ns + rg) % D;                         // stream element T = it * ns + s lands in stage T % D ...
    uint32_t ph = (uint32_t)(((it * ns + rg) / D) & 1);   // ... on its (T / D)-th use
    const bool ragged = (nrows % RR) != 0;
    for (int s = rg; s < ns; s += P2_GROUPS) {
      const float* slab = ring + (size_t)stg * stage_floats;
      mbar_wait(&full[stg], ph);
      float4 z[RR][KQ];
      bool live[RR];
      if (fast) {
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          live[r] = true;
#pragma unroll
          for (int k = 0; k < KQ; ++k) {
            const int c = 4 * (ct + P2_TPR * k);
            if (FULL || c < M) z[r][k] = *reinterpret_cast<const float4*>(slab + (size_t)r * M + c);
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          const int i = row0 + s * RR + r;
          const bool row_in = i < row1;  // rows past the CTA's range were not copied: stale shared memory
          live[r] = row_in && !(p.apply_mask && !smask[i < N ? i : 0]);
#pragma unroll
          for (int k = 0; k < KQ; ++k) {
            const int c = 4 * (ct + P2_TPR * k);
            z[r][k] = (row_in && (FULL || c < M)) ? *reinterpret_cast<const float4*>(slab + (size_t)r * M + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      float* part = red_part + ((buf * P2_GROUPS + rg) * 8) * 8;
      if (scaled) {
        float mh[RR], rs[RR], tot[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          // row reference (log2 domain).  Later iterations: the row's log-sum-exp of the previous iteration.  First
          // iteration: the dustbin-column entry, which every row holds -- the row sum is then >= 1 (no underflow), and
          // whenever the row maximum lies within ~2^100 of it nothing overflows either; rows outside that range are
          // caught below and redone with their exact maximum.  >= 1e30: a row without real entries (masked, or past the
          // CTA's range): every e is 0.
          mh[r] = semi ? (live[r] ? dust2 : 1.0e30f) : lse_prev_s[s * RR + r];
        }
        // e_ij = 2^(x_ij - ref_i) in place of z, row sums reduced over the group (ONE named barrier)
        auto exp_and_row_sums = [&]() {
#pragma unroll
          for (int r = 0; r < RR; ++r) rs[r] = 0.f;
#pragma unroll
          for (int k = 0; k < KQ; ++k) {
            const int c = 4 * (ct + P2_TPR * k);
            if (FULL || c < M) {
              const float4 vv = *reinterpret_cast<const float4*>(v2_s + c);
#pragma unroll
              for (int r = 0; r < RR; ++r) {
                float4 e;
                e.x = ex2(fmaf(z[r][k].x, zs, vv.x) - mh[r]);
                e.y = ex2(fmaf(z[r][k].y, zs, vv.y) - mh[r]);
                e.z = ex2(fmaf(z[r][k].z, zs, vv.z) - mh[r]);
                e.w = ex2(fmaf(z[r][k].w, zs, vv.w) - mh[r]);
                z[r][k] = e;
                rs[r] += (e.x + e.y) + (e.z + e.w);
              }
            } else {
#pragma unroll
              for (int r = 0; r < RR; ++r) z[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if (!semi && ragged && s == ns - 1) {  // rows past the CTA's range were not copied: stale shared memory, possibly NaN
#pragma unroll
            for (int r = 0; r < RR; ++r) {
              if (s * RR + r >= nrows) {
                rs[r] = 0.f;
#pragma unroll
                for (int k = 0; k < KQ; ++k) z[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
          }
#pragma unroll
          for (int r = 0; r < RR; ++r) rs[r] = warp_sum(rs[r]);
          if (lane == 0) {
#pragma unroll
            for (int r = 0; r < RR; ++r) part[r * 8 + wg] = rs[r];
          }
          group_barrier(rg);
#pragma unroll
          for (int r = 0; r < RR; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(part + r * 8);
            const float4 c4 = *reinterpret_cast<const float4*>(part + r * 8 + 4);
            tot[r] = ((a.x + a.y) + (a.z + a.w)) + ((c4.x + c4.y) + (c4.z + c4.w));
          }
          buf ^= 1;
          part = red_part + ((buf * P2_GROUPS + rg) * 8) * 8;
        };
        exp_and_row_sums();
        if (semi) {
          // every thread of the group sees the same totals: a uniform decision
          bool bad = false;
#pragma unroll
          for (int r = 0; r < RR; ++r)
            if (live[r] && !(tot[r] + 1.f < 1.2676506e30f)) bad = true;   // overflow / NaN: the row maximum is > 2^100 above the dustbin entry
          if (bad) {
            // rare (scores with a huge dynamic range): the slab is still in shared memory -- redo with exact row maxima
#pragma unroll
            for (int r = 0; r < RR; ++r) {
              const int i = row0 + s * RR + r;
              const bool row_in = i < row1;
#pragma unroll
              for (int k = 0; k < KQ; ++k) {
                const int c = 4 * (ct + P2_TPR * k);
                z[r][k] = (row_in && (FULL || c < M)) ? *reinterpret_cast<const float4*>(slab + (size_t)r * M + c) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
            float tm[RR];
#pragma unroll
            for (int r = 0; r < RR; ++r) tm[r] = NEG_BIG;
#pragma unroll
            for (int k = 0; k < KQ; ++k) {
              const int c = 4 * (ct + P2_TPR * k);
              if (FULL || c < M) {
                const float4 vv = *reinterpret_cast<const float4*>(v2_s + c);
#pragma unroll
                for (int r = 0; r < RR; ++r)
                  tm[r] = fmaxf(tm[r], fmaxf(fmaxf(fmaf(z[r][k].x, zs, vv.x), fmaf(z[r][k].y, zs, vv.y)),
                                             fmaxf(fmaf(z[r][k].z, zs, vv.z), fmaf(z[r][k].w, zs, vv.w))));
              }
            }
#pragma unroll
            for (int r = 0; r < RR; ++r) tm[r] = warp_max(tm[r]);
            if (lane == 0) {
#pragma unroll
              for (int r = 0; r < RR; ++r) part[r * 8 + wg] = tm[r];
            }
            group_barrier(rg);
#pragma unroll
            for (int r = 0; r < RR; ++r) {
              const float4 a = *reinterpret_cast<const float4*>(part + r * 8);
              const float4 c4 = *reinterpret_cast<const float4*>(part + r * 8 + 4);
              const float m8 = fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(c4.x, c4.y), fmaxf(c4.z, c4.w)));
              mh[r] = live[r] ? fmaxf(m8, dust2) : 1.0e30f;
            }
            buf ^= 1;
            part = red_part + ((buf * P2_GROUPS + rg) * 8) * 8;
            exp_and_row_sums();
          }
        }
        if (ct == 0 && s + D < ns) issue_slab_to(stg, s + D);  // every thread of the group has its registers: re-arm the stage
        float w[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          const bool dead = mh[r] > 1.0e29f;
          const float srow = tot[r] + ex2(dust2 - mh[r]);  // + dustbin column entry
          w[r] = dead ? 0.f : 1.f / srow;               // = 2^(ref_i + u_i log2e - norm2)
          if (ct == 0 && s * RR + r < nrows) {
            const float rowlse2 = dead ? dust2 : mh[r] + lg2(srow);  // a masked row holds the dustbin entry only
            const float ui = fmaf(-rowlse2, LN2, bc.norm);  // (explicit: the candidate-search tail recomputes it bit for bit)
            p.u[(size_t)b * p.ldu + row0 + s * RR + r] = ui;
            lse_add_value(uacc, ui * LOG2E);
            lse_prev_s[s * RR + r] = dead ? 1.0e30f : rowlse2;
          }
        }
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
#pragma unroll
          for (int r = 0; r < RR; ++r) {
            cs[4 * k + 0] = fmaf(z[r][k].x, w[r], cs[4 * k + 0]);
            cs[4 * k + 1] = fmaf(z[r][k].y, w[r], cs[4 * k + 1]);
            cs[4 * k + 2] = fmaf(z[r][k].z, w[r], cs[4 * k + 2]);
            cs[4 * k + 3] = fmaf(z[r][k].w, w[r], cs[4 * k + 3]);
          }
        }
      } else {
        // log-domain pass: exact row maximum first, then the exponential sum, then an online column log-sum-exp
        float tm[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) tm[r] = NEG_BIG;
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          const int c = 4 * (ct + P2_TPR * k);
          if (FULL || c < M) {
            const float4 vv = *reinterpret_cast<const float4*>(v2_s + c);
#pragma unroll
            for (int r = 0; r < RR; ++r) {
              const float t = fmaxf(fmaxf(fmaf(z[r][k].x, zs, vv.x), fmaf(z[r][k].y, zs, vv.y)),
                                    fmaxf(fmaf(z[r][k].z, zs, vv.z), fmaf(z[r][k].w, zs, vv.w)));
              if (live[r]) tm[r] = fmaxf(tm[r], t);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < RR; ++r) tm[r] = warp_max(tm[r]);
        if (lane == 0) {
#pragma unroll
          for (int r = 0; r < RR; ++r) part[r * 8 + wg] = tm[r];
        }
        group_barrier(rg);
        if (ct == 0 && s + D < ns) issue_slab_to(stg, s + D);  // every thread of the group has its registers: re-arm the stage
        float mrow[RR], rs[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(part + r * 8);
          const float4 c4 = *reinterpret_cast<const float4*>(part + r * 8 + 4);
          mrow[r] = fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(c4.x, c4.y), fmaxf(c4.z, c4.w)));
          mrow[r] = fmaxf(mrow[r], dust2);  // the dustbin column entry is always finite
          rs[r] = 0.f;
        }
        buf ^= 1;
        part = red_part + ((buf * P2_GROUPS + rg) * 8) * 8;
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          const int c = 4 * (ct + P2_TPR * k);
          if (FULL || c < M) {
            const float4 vv = *reinterpret_cast<const float4*>(v2_s + c);
#pragma unroll
            for (int r = 0; r < RR; ++r) {
              if (live[r]) {
                rs[r] += (ex2(fmaf(z[r][k].x, zs, vv.x) - mrow[r]) + ex2(fmaf(z[r][k].y, zs, vv.y) - mrow[r])) +
                         (ex2(fmaf(z[r][k].z, zs, vv.z) - mrow[r]) + ex2(fmaf(z[r][k].w, zs, vv.w) - mrow[r]));
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < RR; ++r) rs[r] = warp_sum(rs[r]);
        if (lane == 0) {
#pragma unroll
          for (int r = 0; r < RR; ++r) part[r * 8 + wg] = rs[r];
        }
        group_barrier(rg);
        float u2[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(part + r * 8);
          const float4 c4 = *reinterpret_cast<const float4*>(part + r * 8 + 4);
          const float tot = ((a.x + a.y) + (a.z + a.w)) + ((c4.x + c4.y) + (c4.z + c4.w));
          const float srow = tot + ex2(dust2 - mrow[r]);
          const float rowlse2 = mrow[r] + lg2(srow);
          const float ui = fmaf(-rowlse2, LN2, bc.norm);  // (explicit: the candidate-search tail recomputes it bit for bit)
          u2[r] = live[r] ? (ui - shift) * LOG2E : -INFINITY;  // column pass sees (S - shift) + u
          const int i = row0 + s * RR + r;
          if (ct == 0 && i < row1) {
            p.u[(size_t)b * p.ldu + i] = ui;
            lse_add_value(uacc, ui * LOG2E);
            lse_prev_s[s * RR + r] = live[r] ? rowlse2 : 1.0e30f;
          }
        }
        buf ^= 1;
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          const int c = 4 * (ct + P2_TPR * k);
          if (FULL || c < M) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float y[RR];
              float mx = NEG_BIG;
#pragma unroll
              for (int r = 0; r < RR; ++r) {
                const float zz = (e == 0) ? z[r][k].x : (e == 1) ? z[r][k].y : (e == 2) ? z[r][k].z : z[r][k].w;
                y[r] = fmaf(zz, zs, u2[r]);
                mx = fmaxf(mx, y[r]);
              }
              float& am = cm[4 * k + e];
              float& as = cs[4 * k + e];
              if (mx > am + 32.f) {  // lazy re-reference
                as *= ex2(am - mx);
                am = mx;
              }
              float acc = 0.f;
#pragma unroll
              for (int r = 0; r < RR; ++r) acc += ex2(y[r] - am);
              as += acc;
            }
          }
        }
      }
      stg += P2_GROUPS;
      while (stg >= D) {
        stg -= D;
        ph ^= 1u;
      }
    }
    if (tid == 0) DRG_STAMP(10 + it * 100 + 2);
    if (scaled) {
      // sum_i 2^(x_ij + u_i log2e) = 2^(norm2 - v_j log2e) * sum_i e_ij w_i   (v2_s holds (v_j - shift) log2e)
#pragma unroll
      for (int k = 0; k < KQ; ++k) {
        const int c = 4 * (ct + P2_TPR * k);
        if (FULL || c < M) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v2j = v2_s[c + e];
            const bool okc = v2j > -INFINITY;
            cm[4 * k + e] = okc ? (norm2 - v2j - shift2) : NEG_BIG;
            if (!okc) cs[4 * k + e] = 0.f;
          }
        }
      }
    }
    // ---- the two row groups combine their column partials; row group 0 writes them
    __syncthreads();  // the ring is idle from here until the next iteration's first slabs are requested below
    // the dustbin-row potential from the prologue's per-warp partials (red_s is not touched by the pass)
    float uN;
    if (it == 0) {
      uN = bc.log_mu_bin - (alpha + logf((float)(M + 1)));
    } else {
      float mall = red_s[0];
#pragma unroll
      for (int w = 1; w < P2_THREADS / 32; ++w) mall = fmaxf(mall, red_s[w]);
      float sall = 0.f;
#pragma unroll
      for (int w = 0; w < P2_THREADS / 32; ++w) sall += red_s[32 + w] * ex2(red_s[w] - mall);
      uN = bc.log_mu_bin - (alpha + (mall + lg2(sall)) * LN2);
    }
    if (g == 0 && tid == 0) p.u[(size_t)b * p.ldu + N] = uN;
    if (rg >= 1) {
#pragma unroll
      for (int e = 0; e < KQ * 4; ++e) xcomb[((rg - 1) * KQ * 4 + e) * P2_TPR + ct] = make_float2(cm[e], cs[e]);
    }
    if (ct == 0) upart_s[rg] = make_float2(uacc.m, uacc.s);
    __syncthreads();
    if (rg == 0) {
#pragma unroll
      for (int e = 0; e < KQ * 4; ++e) {
#pragma unroll
        for (int og = 0; og < P2_GROUPS - 1; ++og) {
          const float2 o = xcomb[(og * KQ * 4 + e) * P2_TPR + ct];
          if (scaled) {  // all groups carry the same reference
            cs[e] += o.y;
          } else {
            LseAcc a{cm[e], cs[e]};
            lse_merge(a, o.x, o.y);
            cm[e] = a.m;
            cs[e] = a.s;
          }
        }
      }
      float2* cp = p.colpart + ((size_t)b * G + g) * M;
#pragma unroll
      for (int k = 0; k < KQ; ++k) {
        const int c = 4 * (ct + P2_TPR * k);
        if (FULL || c < M) {
          *reinterpret_cast<float4*>(cp + c) = make_float4(cm[4 * k], cs[4 * k], cm[4 * k + 1], cs[4 * k + 1]);
          *reinterpret_cast<float4*>(cp + c + 2) = make_float4(cm[4 * k + 2], cs[4 * k + 2], cm[4 * k + 3], cs[4 * k + 3]);
        }
      }
      group_barrier(0);  // xcomb (aliased on the ring) has been read
      // (issued by thread 32, not thread 0: thread 0 carries the CTA's arrival at the grid barrier below, and six TMA issues
      //  with their proxy fences would sit in front of it)
      if (tid == 32 && (it + 1 < iters || do_collect))
        for (int s = 0; s < D && s < ns; ++s) issue_slab_to(((it + 1) * ns + s) % D, s);  // the scores do not change: prefetch across the barriers
      if (tid == 32 && it + 1 == iters && do_final)
        for (int q = 0; q < Df && q < NWf * ns; ++q) issue_final(q);   // neither do the final phase's inputs
    }
    if (tid == 0) {
      LseAcc a{upart_s[0].x, upart_s[0].y};
      for (int og = 1; og < P2_GROUPS; ++og) lse_merge(a, upart_s[og].x, upart_s[og].y);
      p.upart[(size_t)b * G + g] = make_float2(a.m, a.s);
      DRG_STAMP(10 + it * 100 + 3);
    }
    grid_barrier_ra(gcount, (unsigned int)G * (++barriers_done));
    if (tid == 0) DRG_STAMP(10 + it * 100 + 4);

    // ---- merge: a CTA takes 32 consecutive columns at a time (lane = column), its 16 warps split the G partials
    {
      float* v_b = p.v + (size_t)b * p.ldv;
      if (g == 0 && tid == 0) dv_slots[(it + 1) & 1] = 0u;
      float dv_loc = 0.f;
      const bool sampling = do_collect && it + 1 == iters;  // the final v of this CTA's columns is known right here
      if (sampling)
        for (int q = tid; q < TK_BINS; q += P2_THREADS) tail_s[q] = 0u;
      for (int j0 = g * 32; j0 <= M; j0 += G * 32) {
        const int j = j0 + lane;
        const bool in_range = j <= M;
        const bool is_bin = (j == M);
        const bool col_ok = in_range && (is_bin || !p.apply_mask || p.tgt_mask[(size_t)b * M + j]);
        float m = NEG_BIG, sum = 0.f;
        if (col_ok) {
          const float2* src = is_bin ? (p.upart + (size_t)b * G) : (p.colpart + (size_t)b * G * M + j);
          const size_t gstride = is_bin ? 1 : (size_t)M;
          constexpr int NW = P2_THREADS / 32;
          constexpr int PER_WARP = (NUM_SMS + NW - 1) / NW;
          float2 q[PER_WARP];
#pragma unroll
          for (int k = 0; k < PER_WARP; ++k) {
            const int gg = warp + NW * k;
            q[k] = (gg < G) ? __ldcg(src + (size_t)gg * gstride) : make_float2(NEG_BIG, 0.f);
            m = fmaxf(m, q[k].x);
          }
#pragma unroll
          for (int k = 0; k < PER_WARP; ++k) sum += q[k].y * ex2(q[k].x - m);
        }
        comb[warp * 32 + lane] = make_float2(m, sum);
        __syncthreads();
        if (warp == 0 && in_range) {
          float mm = NEG_BIG;
#pragma unroll
          for (int w = 0; w < P2_THREADS / 32; ++w) mm = fmaxf(mm, comb[w * 32 + lane].x);
          float ss = 0.f;
#pragma unroll
          for (int w = 0; w < P2_THREADS / 32; ++w) {
            const float2 c2 = comb[w * 32 + lane];
            ss += c2.y * ex2(c2.x - mm);
          }
          LseAcc a{mm, ss};
          const float v_old = (it == 0) ? 0.f : __ldcg(v_b + j);
          float v_new;
          if (!is_bin) {
            lse_add_value(a, (alpha + uN) * LOG2E);  // dustbin row entry
            v_new = bc.norm - lse_value(a) * LN2;
          } else {
            lse_add_value(a, uN * LOG2E);  // c_M = alpha + LSE(u[0..N])
            v_new = bc.log_nu_bin - (alpha + lse_value(a) * LN2);
          }
          v_b[j] = v_new;
          const float d = fabsf(v_new - v_old) * LOG2E;
          dv_loc = fmaxf(dv_loc, (d == d) ? d : INFINITY);
          if (sampling) red_s[lane] = (col_ok && !is_bin) ? v_new : -INFINITY;
        } else if (sampling && warp == 0) {
          red_s[lane] = -INFINITY;
        }
        __syncthreads();
        if (sampling) {
          // SH_ROWS sampled rows per column (warp w takes the w-th), their log2 confidences into the CTA's histogram.
          // A sample is exp(Z + u + v - norm) at its FINAL potentials: u was final before the grid barrier above.
          const float vn = red_s[lane];
          const int jc = j0 + lane;
          int i = warp;
          if (N > SH_ROWS) i = (int)(((unsigned long long)hash_u32((unsigned int)(jc * SH_ROWS + warp) + 0x9e3779b9u * (unsigned int)(b + 1)) *
                                      (unsigned long long)N) >> 32);
          float la2 = -INFINITY;
          if (vn > -INFINITY && i < N && !(p.apply_mask && !smask[i])) {
            const float z = __ldcg(sc_b + (size_t)i * M + jc);
            const float ui = __ldcg(p.u + (size_t)b * p.ldu + i);
            la2 = ((((z - shift) + ui) + vn) - bc.norm) * LOG2E;
            if (la2 > -INFINITY) {  // (false for NaN)
              const float fb = fminf(fmaxf((la2 + SH_OFFSET) * (float)SH_PER_OCTAVE, 0.f), (float)(TK_BINS - 1));
              atomicAdd(&tail_s[(int)fb], 1u);
            } else {
              la2 = -INFINITY;
            }
          }
          if (jc <= M) p.col.sample_val[((size_t)b * SH_ROWS + warp) * p.ldv + jc] = la2;  // kept for a second-level binning
        }
      }
      if (sampling) {
        __syncthreads();
        for (int q = tid; q < TK_BINS; q += P2_THREADS) {
          const unsigned int c = tail_s[q];
          if (c) atomicAdd(&p.col.sample_hist[(size_t)b * TK_BINS + q], c);
        }
      }
      if (warp == 0) {
        dv_loc = warp_max(dv_loc);
        if (lane == 0 && dv_loc > 0.f) atomicMax(dv_slots + (it & 1), __float_as_uint(dv_loc));
      }
    }
    if (tid == 0) DRG_STAMP(10 + it * 100 + 5);
    if (it + 1 < iters || do_collect || do_final) grid_barrier_ra(gcount, (unsigned int)G * (++barriers_done));
    if (tid == 0) DRG_STAMP(10 + it * 100 + 6);
  }
  if (do_final) {
    // =====================================================================================================
    // Final pass over this CTA's rows (exp(Z + u + v - norm)[:-1, :-1], matching.py:169-170, or the DDIM update
    // pipeline.py:180-190 with the noise drawn here), the arithmetic of skh_final_tile_kernel element for element (the two
    // paths agree bit for bit).  The score slab comes out of L2 through the same ring, x_t out of DRAM through the other
    // half of it.  The pass is ISSUE-bound (Philox + Box-Muller + exp per element on 16 warps per SM): the Philox round keys
    // come precomputed from the host (constant-bank operands) and the running minimum of the 3DMatch flavour is not
    // offered here (those calls take the stand-alone kernel).
    // =====================================================================================================
    if (tid == 0) DRG_STAMP(750);
    const SkhFused& f = p.fin;
    const float* v_b = p.v + (size_t)b * p.ldv;
    const float* u_b = p.u + (size_t)b * p.ldu;
    const bool masked = p.apply_mask && bc.pad != 1.f;
    const bool track = f.rowbest != nullptr;
    const float floor_v = f.best_floor;
    const float xt_shift = f.xt_shift ? *f.xt_shift : 0.f;
    const unsigned long long noise_offset = f.noise_offset + (f.noise_offset_dev ? *f.noise_offset_dev : 0ull);
    const unsigned int off_lo = (unsigned int)noise_offset, off_hi = (unsigned int)(noise_offset >> 32);
    float vj[KQ][4];
    unsigned int tmbits = 0xffffffffu;   // bit 4 k + e: target column valid
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
      const int c = 4 * (ct + P2_TPR * k);
      float4 vq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (FULL || c < M) {
        vq = __ldcg(reinterpret_cast<const float4*>(v_b + c));
        if (masked) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (!p.tgt_mask[(size_t)b * M + c + e]) tmbits &= ~(1u << (4 * k + e));
        }
      }
      vj[k][0] = vq.x; vj[k][1] = vq.y; vj[k][2] = vq.z; vj[k][3] = vq.w;
    }
    // mbarrier phase of every stage after the iteration passes: stage k has completed ceil((Ttot - k) / D) uses
    const int Ttot = iters * ns;
    auto base_parity = [&](int k) -> uint32_t { return (uint32_t)((Ttot > k ? (Ttot - k + D - 1) / D : 0) & 1); };
    float* out_b = f.out + (size_t)b * N * M;
    float* conf_b = f.conf ? f.conf + (size_t)b * N * M : nullptr;
    const float* noise_b = f.noise ? f.noise + (size_t)b * N * M : nullptr;
    const unsigned long long quad_base = ((unsigned long long)b * N * M) >> 2;   // Philox counter = global quad index
    // NOISE: 0 none, 1 caller tensor, 2 in-kernel Philox.  The slab's RR x KQ quads are straight-line code (the compiler
    // interleaves their Philox / exp chains -- the phase is latency-bound on 16 warps per SM); the arg-max bookkeeping of
    // the rare quads that hold an entry above the floor is deferred to after the slab (a flag bit per quad, the few
    // confidences recomputed from the slab still in shared memory), so no data-dependent branch splits the hot code.
    auto body = [&](auto masked_c, auto ddim_c, auto noise_c) {
      constexpr bool MASKED = decltype(masked_c)::value;
      constexpr bool DDIM = decltype(ddim_c)::value;
      constexpr int NOISE = decltype(noise_c)::value;
      for (int s = rg; s < ns; s += P2_GROUPS) {
        const int q0 = NWf * s;
        const int stg_z = q0 % Df;
        mbar_wait(&full[stg_z], (base_parity(stg_z) + (uint32_t)(q0 / Df)) & 1u);
        const float* slab_z = ring + (size_t)stg_z * stage_floats;
        const float* slab_x = slab_z;
        if (DDIM) {
          const int stg_x = (q0 + 1) % Df;
          mbar_wait(&full[stg_x], (base_parity(stg_x) + (uint32_t)((q0 + 1) / Df)) & 1u);
          slab_x = ring + (size_t)stg_x * stage_floats;
        }
        unsigned int hot = 0u;   // bit r * KQ + k: that quad holds an entry above the floor
        float ui_r[RR];
        bool ok_r[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          const int i = min(row0 + s * RR + r, row1 - 1);   // (rows past the CTA's range: clamped, nothing is stored for them)
          ui_r[r] = __ldcg(u_b + i);
          ok_r[r] = !MASKED || smask[i];
        }
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          const int i = row0 + s * RR + r;
          const bool row_in = i < row1;   // rows past the CTA's range were not copied
          const size_t row_off = (size_t)min(i, row1 - 1) * M + 4 * ct;
#pragma unroll
          for (int k = 0; k < KQ; ++k) {
            const int c = 4 * (ct + P2_TPR * k);
            if (FULL || c < M) {
              const float4 z4 = *reinterpret_cast<const float4*>(slab_z + (size_t)r * M + c);
              float nz[4] = {0.f, 0.f, 0.f, 0.f};
              float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (DDIM) {
                t4 = *reinterpret_cast<const float4*>(slab_x + (size_t)r * M + c);
                if (NOISE == 1) {
                  const float4 n4 = ldg_stream4(noise_b + row_off + 4 * P2_TPR * k);
                  nz[0] = n4.x; nz[1] = n4.y; nz[2] = n4.z; nz[3] = n4.w;
                } else if (NOISE == 2) {
                  const unsigned long long quad = quad_base + ((row_off + 4 * P2_TPR * k) >> 2);
                  const float4 n4 = philox_normal4_rk((unsigned int)quad, (unsigned int)(quad >> 32), off_lo, off_hi, f.rk);
                  nz[0] = n4.x; nz[1] = n4.y; nz[2] = n4.z; nz[3] = n4.w;
                }
              }
              const float z[4] = {z4.x, z4.y, z4.z, z4.w};
              const float xt[4] = {t4.x, t4.y, t4.z, t4.w};
              float cf[4], o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const bool ok = !MASKED || (ok_r[r] && ((tmbits >> (4 * k + e)) & 1u));
                const float zz = ok ? (z[e] - shift) : -INFINITY;
                const float la = ((zz + ui_r[r]) + vj[k][e]) - bc.norm;  // same association as matching.py:34-36
                cf[e] = ex2(la * LOG2E);
                if (DDIM) o[e] = ok ? fmaf(f.k_x0, cf[e], fmaf(f.k_xt, xt[e] - xt_shift, f.sigma * nz[e])) : -INFINITY;
                else o[e] = cf[e];
              }
              if (row_in) {
                *reinterpret_cast<float4*>(out_b + row_off + 4 * P2_TPR * k) = make_float4(o[0], o[1], o[2], o[3]);
                if (DDIM && conf_b) *reinterpret_cast<float4*>(conf_b + row_off + 4 * P2_TPR * k) = make_float4(cf[0], cf[1], cf[2], cf[3]);
              }
              if (row_in && fmaxf(fmaxf(cf[0], cf[1]), fmaxf(cf[2], cf[3])) > floor_v) hot |= 1u << (r * KQ + k);
            }
          }
        }
        // arg-max keys: only entries above the floor (a row of the plan sums to <= 1: a handful per row) -- straight to the
        // packed 64-bit keys with atomicMax (largest value, then lowest index); the confidences are recomputed exactly
        if (track && hot) {
#pragma unroll 1
          while (hot) {
            const int bit = __ffs(hot) - 1;
            hot &= hot - 1u;
            const int r = bit / KQ, k = bit - r * KQ;
            const int i = row0 + s * RR + r;
            const int c = 4 * (ct + P2_TPR * k);
            const float4 z4 = *reinterpret_cast<const float4*>(slab_z + (size_t)r * M + c);
            const float z[4] = {z4.x, z4.y, z4.z, z4.w};
            const float ui = __ldcg(u_b + i);
            const bool row_ok = !MASKED || smask[i];
            const float4 vq = __ldcg(reinterpret_cast<const float4*>(v_b + c));
            const float vv[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool ok = !MASKED || (row_ok && p.tgt_mask[(size_t)b * M + c + e]);
              const float zz = ok ? (z[e] - shift) : -INFINITY;
              const float cfe = ex2((((zz + ui) + vv[e]) - bc.norm) * LOG2E);
              if (cfe > floor_v) {
                const unsigned long long hi = (unsigned long long)float_to_ordered(cfe) << 32;
                atomicMax(&f.rowbest[(size_t)b * N + i], hi | (unsigned long long)(0xFFFFFFFFu - (unsigned int)(c + e)));
                atomicMax(&f.colbest[(size_t)b * M + c + e], hi | (unsigned long long)(0xFFFFFFFFu - (unsigned int)i));
              }
            }
          }
        }
        group_barrier(rg);   // every thread of the group is done with the two stages: re-arm them for mini-slab s + Df / NWf
        if (ct == 0) {
          const int s2 = s + Df / NWf;
          if (s2 < ns) {
            issue_final(NWf * s2);
            if (DDIM) issue_final(NWf * s2 + 1);
          }
        }
      }
    };
    using T_ = std::true_type;
    using F_ = std::false_type;
    using N0 = std::integral_constant<int, 0>;
    using N1 = std::integral_constant<int, 1>;
    using N2 = std::integral_constant<int, 2>;
    const int noise_mode = !f_ddim ? 0 : noise_b ? 1 : f.gen_noise ? 2 : 0;
    if (masked) {
      if (!f_ddim) body(T_{}, F_{}, N0{});
      else if (noise_mode == 2) body(T_{}, T_{}, N2{});
      else if (noise_mode == 1) body(T_{}, T_{}, N1{});
      else body(T_{}, T_{}, N0{});
    } else {
      if (!f_ddim) body(F_{}, F_{}, N0{});
      else if (noise_mode == 2) body(F_{}, T_{}, N2{});
      else if (noise_mode == 1) body(F_{}, T_{}, N1{});
      else body(F_{}, T_{}, N0{});
    }
    if (tid == 0) DRG_STAMP(751);
  }
  if (do_collect) {
    // =====================================================================================================
    // Candidate search of SoftProcrustes (procrustes.py:61-76) as a last pass over the L2-resident slab:
    //   bound   the sampled log2 confidences (SH_ROWS per column, drawn by the CTAs that merged the columns) are in a
    //           global histogram; every CTA walks it and takes the lower edge of the bin holding the t-th largest
    //           sample, t ~ twice the expected number of top-K_b entries in the sample (+16) -- a value Lv that
    //           ~(2 K_b + N M / (SH_ROWS M / 16)) entries of the matrix reach, whatever their distribution;
    //   pass    1 FFMA + 1 compare per element against thr_i = log2(Lv) + (row log-sum-exp of the last iteration);
    //           the few that pass are evaluated exactly (the association of the final pass, so that the weights equal
    //           the confidences a stored matrix would hold) and appended with conf >= Lv;
    //   the candidates' histogram (first level of the select in procr_pose_kernel) is filled as they are flushed.
    // =====================================================================================================
    __shared__ int tail_i[8];     // Kb, crossing bin, top non-empty bin, crowded?, rank wanted inside the crossing bin, sub-bin, rank inside it, target
    __shared__ unsigned long long tail_k64;   // exact 64-bit key bound (third level); 0: the bound is the value Lv alone
    __shared__ float tail_f[3];   // Lv, log2 Lv, log2 of (an upper bound of) the largest sample
    __shared__ unsigned int cand_n_s, cand_base_s;
    const float* v_b = p.v + (size_t)b * p.ldv;
    if (tid == 0) DRG_STAMP(700);
    // this thread's columns: the final v (raw, for the exact expression) and its log2-domain form (for the filter)
    float vraw[KQ][4], v2c[KQ][4];
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
      const int c = 4 * (ct + P2_TPR * k);
      float4 vq = make_float4(0.f, 0.f, 0.f, 0.f);
      unsigned char mm[4] = {1, 1, 1, 1};
      const bool in = FULL || c < M;
      if (in) {
        vq = __ldcg(reinterpret_cast<const float4*>(v_b + c));
        if (p.apply_mask) {
#pragma unroll
          for (int e = 0; e < 4; ++e) mm[e] = p.tgt_mask[(size_t)b * M + c + e];
        }
      }
      const float vr[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        vraw[k][e] = vr[e];
        v2c[k][e] = (in && mm[e]) ? (vr[e] - shift) * LOG2E : -INFINITY;
      }
    }
    for (int q = tid; q < TK_BINS; q += P2_THREADS) tail_s[q] = __ldcg(p.col.sample_hist + (size_t)b * TK_BINS + q);
    __syncthreads();
    if (warp == 0) {
      long long target = Kb;  // N <= SH_ROWS: the sample is the whole matrix
      if (N > SH_ROWS) target = (long long)(2.0 * (double)Kb * (double)SH_ROWS / (double)N + 16.0);
      int bin = 0;
      unsigned int cum = 0u, hsel = 0u;
      if (target >= 1) warp_walk_hist(tail_s, (unsigned int)min(target, 0x7fffffffll), bin, cum, hsel);
      // highest non-empty bin (range of the candidate histogram)
      int top = 0;
      for (int q = TK_BINS - 1 - lane; q >= 0; q -= 32)
        if (tail_s[q]) {
          top = q;
          break;
        }
      top = __reduce_max_sync(0xffffffffu, top);
      if (lane == 0) {
        tail_i[0] = Kb;
        tail_i[1] = bin;
        tail_i[2] = top;
        tail_i[3] = (bin > 0 && target >= 1 && (long long)hsel > target) ? 1 : 0;   // crowded: refine
        tail_i[7] = (int)min(target, 0x7fffffffll);
        tail_i[4] = (int)min(target - (long long)cum, 0x7fffffffll);                // rank still wanted inside the bin
        tail_f[1] = (float)bin / (float)SH_PER_OCTAVE - SH_OFFSET;
        tail_f[0] = bin > 0 ? exp2f(tail_f[1]) : 0.f;
        tail_f[2] = (float)(top + 1) / (float)SH_PER_OCTAVE - SH_OFFSET;
      }
    }
    __syncthreads();
    const int bin = tail_i[1];
    // A crowded crossing bin (more samples in it than the whole target: a narrow distribution -- the noise-free 3DMatch /
    // 2D-3D samplers produce nearly flat plans -- whose bin edge would let a multiple of the wanted candidates through)
    // is re-binned: every CTA adds ITS samples of that bin to a second-level histogram (TK_BINS sub-bins across the bin),
    // one more grid barrier, and the bound becomes a sub-bin edge: 1 / (16 * 2048) of an octave.  All CTAs read the same
    // first-level histogram, so they take this branch together.
    if (tail_i[3]) {
      const float edge_lo = (float)bin / (float)SH_PER_OCTAVE - SH_OFFSET;
      for (int j0 = g * 32; j0 <= M; j0 += G * 32) {
        const int jc = j0 + lane;
        if (jc < M) {
          const float la2 = __ldcg(p.col.sample_val + ((size_t)b * SH_ROWS + warp) * p.ldv + jc);
          if (la2 > -INFINITY) {
            const float fb = fminf(fmaxf((la2 + SH_OFFSET) * (float)SH_PER_OCTAVE, 0.f), (float)(TK_BINS - 1));
            if ((int)fb == bin) {
              const float fs = fminf(fmaxf((la2 - edge_lo) * (float)(SH_PER_OCTAVE * TK_BINS), 0.f), (float)(TK_BINS - 1));
              atomicAdd(&p.col.sample_hist2[(size_t)b * TK_BINS + (int)fs], 1u);
            }
          }
        }
      }
      grid_barrier_ra(gcount, (unsigned int)G * (++barriers_done));
      for (int q = tid; q < TK_BINS; q += P2_THREADS) tail_s[q] = __ldcg(p.col.sample_hist2 + (size_t)b * TK_BINS + q);
      __syncthreads();
      if (warp == 0) {
        int sub = 0;
        unsigned int cum2, hsel2;
        warp_walk_hist(tail_s, (unsigned int)tail_i[4], sub, cum2, hsel2);   // the remaining rank inside the bin
        int top2 = 0;
        for (int q = TK_BINS - 1 - lane; q >= 0; q -= 32)
          if (tail_s[q]) {
            top2 = q;
            break;
          }
        top2 = __reduce_max_sync(0xffffffffu, top2);
        if (lane == 0) {
          const float thr = edge_lo + (float)sub / (float)(SH_PER_OCTAVE * TK_BINS);
          tail_f[0] = exp2f(thr);
          tail_f[1] = thr;
          // the largest sample: inside this bin (then known to a sub-bin) or in a higher first-level bin
          tail_f[2] = (tail_i[2] == bin) ? edge_lo + (float)(top2 + 1) / (float)(SH_PER_OCTAVE * TK_BINS) : tail_f[2];
          tail_i[5] = sub;
          tail_i[6] = (int)tail_i[4] - (int)cum2;                     // rank still wanted inside the sub-bin
          tail_i[3] = ((long long)hsel2 > 2ll * tail_i[7] + 64) ? 2 : 1;  // still crowded: values (nearly) tied -> exact keys
        }
      }
      __syncthreads();
      if (tail_i[3] == 2) {
        // Third level, exact: the samples of the crowded sub-bin as 64-bit keys (value, ~flat index) in a global list, one
        // more grid barrier, and every CTA selects the wanted rank among them itself (general radix select: slow, but this
        // is the path of nearly flat / tied plans).  The bound is then a KEY: ties in value are cut by index, so the
        // candidate count stays ~2 K_b whatever the distribution.
        const int sub = tail_i[5];
        for (int j0 = g * 32; j0 <= M; j0 += G * 32) {
          const int jc = j0 + lane;
          if (jc < M) {
            const float la2 = __ldcg(p.col.sample_val + ((size_t)b * SH_ROWS + warp) * p.ldv + jc);
            if (la2 > -INFINITY) {
              const float fb = fminf(fmaxf((la2 + SH_OFFSET) * (float)SH_PER_OCTAVE, 0.f), (float)(TK_BINS - 1));
              const float fs = fminf(fmaxf((la2 - edge_lo) * (float)(SH_PER_OCTAVE * TK_BINS), 0.f), (float)(TK_BINS - 1));
              if ((int)fb == bin && (int)fs == sub) {
                int i = warp;
                if (N > SH_ROWS) i = (int)(((unsigned long long)hash_u32((unsigned int)(jc * SH_ROWS + warp) + 0x9e3779b9u * (unsigned int)(b + 1)) *
                                            (unsigned long long)N) >> 32);
                const unsigned int pos = atomicAdd(&p.col.sample_list_n[b], 1u);
                p.col.sample_list[(size_t)b * SH_ROWS * p.ldv + pos] = make_key64(float_to_ordered(ex2(la2)), (unsigned int)((size_t)i * M + jc));
              }
            }
          }
        }
        grid_barrier_ra(gcount, (unsigned int)G * (++barriers_done));
        const unsigned int n3 = __ldcg(p.col.sample_list_n + b);
        const unsigned long long* lst = p.col.sample_list + (size_t)b * SH_ROWS * p.ldv;
        __shared__ SelectCtl sel_ctl;
        const int want3 = max(1, min(tail_i[6], (int)n3));
        const unsigned long long T3 = block_select_kth<P2_THREADS>([&](size_t e) { return __ldcg(lst + e); }, (size_t)n3, want3, tail_s, sel_ctl);
        if (tid == 0) {
          tail_k64 = T3;
          tail_f[0] = ordered_to_float((unsigned int)(T3 >> 32));
        }
        __syncthreads();
      }
    }
    const float Lv = tail_f[0];
    const bool exact_bound = tail_i[3] == 2;
    // candidates are the entries with 64-bit key (value, ~flat index) >= lower64; without the third level that is conf >= Lv
    const unsigned long long lower64 = exact_bound ? tail_k64 : ((unsigned long long)float_to_ordered(Lv) << 32);
    const float thr2 = bin > 0 ? tail_f[1] - 1.0e-3f : -INFINITY;  // pre-filter margin (its arithmetic is folded differently)
    const unsigned int kmin = float_to_ordered(Lv);
    const unsigned int smax = float_to_ordered(exp2f(tail_f[2]));
    const int hist_sh = cand_hist_shift(kmin, smax, bin > 0);
    if (g == 0 && tid == 0) {
      ProcrState* sp = p.col.state + b;  // n_cand is being accumulated by the CTAs: field-wise
      sp->Kb = Kb;
      sp->lower_key = lower64;
      sp->T = 0ull;
      sp->hist_kmin = kmin;
      sp->hist_sh = hist_sh;
      sp->sel_count = 0u;
      sp->pad_ = 0u;
      sp->seg_G = G;
    }
    if (tid == 0) cand_n_s = 0u;
    if (tid == 0) DRG_STAMP(701);
    __syncthreads();  // tail_s becomes the candidate list: keys [0, CL_CAP), flat indices [CL_CAP, 2 CL_CAP)
    constexpr int CL_CAP = TK_BINS / 2;
    unsigned int* gkey = p.col.cand_key + (size_t)b * N * M;
    unsigned int* gidx = p.col.cand_idx + (size_t)b * N * M;
    int stg = (iters * ns + rg) % D;
    uint32_t ph = (uint32_t)(((iters * ns + rg) / D) & 1);
    if (Kb > 0) {
      for (int s = rg; s < ns; s += P2_GROUPS) {
        const float* slab = ring + (size_t)stg * stage_floats;
        mbar_wait(&full[stg], ph);
        float4 z[RR][KQ];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
#pragma unroll
          for (int k = 0; k < KQ; ++k) {
            const int c = 4 * (ct + P2_TPR * k);
            if (FULL || c < M) z[r][k] = *reinterpret_cast<const float4*>(slab + (size_t)r * M + c);
          }
        }
        group_barrier(rg);
        if (ct == 0 && s + D < ns) issue_slab_to(stg, s + D);  // every thread of the group has its registers: re-arm the stage
        float tr[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          const float rl = lse_prev_s[s * RR + r];          // >= 1e30: masked row / slot past the CTA's rows (stale data)
          tr[r] = rl > 1.0e29f ? INFINITY : thr2 + rl;     // conf_ij = 2^(x_ij - rowlse_i)
        }
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          const int c = 4 * (ct + P2_TPR * k);
          if (FULL || c < M) {
#pragma unroll
            for (int r = 0; r < RR; ++r) {
              const bool hit = (fmaf(z[r][k].x, zs, v2c[k][0]) >= tr[r]) | (fmaf(z[r][k].y, zs, v2c[k][1]) >= tr[r]) |
                               (fmaf(z[r][k].z, zs, v2c[k][2]) >= tr[r]) | (fmaf(z[r][k].w, zs, v2c[k][3]) >= tr[r]);
              if (hit) {  // rare (~0.1 % of the quads): the exact expression of the final pass, then conf >= Lv; no global load
                const int i = row0 + s * RR + r;
                const float ui = fmaf(-lse_prev_s[s * RR + r], LN2, bc.norm);  // = p.u[i], bit for bit (see the passes)
                const float zq[4] = {z[r][k].x, z[r][k].y, z[r][k].z, z[r][k].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (!(v2c[k][e] > -INFINITY)) continue;  // masked column
                  const float cf = ex2(((((zq[e] - shift) + ui) + vraw[k][e]) - bc.norm) * LOG2E);
                  const unsigned int key = float_to_ordered(cf);
                  const unsigned int fi = (unsigned int)((size_t)i * M + c + e);
                  if (make_key64(key, fi) >= lower64) {
                    const unsigned int pos = atomicAdd(&cand_n_s, 1u);
                    if (pos < (unsigned int)CL_CAP) {
                      tail_s[pos] = key;
                      tail_s[CL_CAP + pos] = fi;
                    } else {  // the CTA's list is full (dense candidates: degenerate inputs): straight to the global list
                      p.col.state[b].seg_broken = 1u;
                      const unsigned int gp = atomicAdd(&p.col.state[b].n_cand, 1u);
                      gkey[gp] = key;
                      gidx[gp] = fi;
                      atomicAdd(&p.col.cand_hist[(size_t)b * TK_BINS + cand_bin(key, kmin, hist_sh)], 1u);
                    }
                  }
                }
              }
            }
          }
        }
        stg += P2_GROUPS;
        while (stg >= D) {
          stg -= D;
          ph ^= 1u;
        }
      }
    } else {
      // nothing to select: drain the prefetched slabs so that no bulk copy is in flight when the CTA exits
      for (int s = rg; s < ns && s < D; s += P2_GROUPS) {
        mbar_wait(&full[stg], ph);
        stg += P2_GROUPS;
        while (stg >= D) {
          stg -= D;
          ph ^= 1u;
        }
      }
    }
    __syncthreads();
    if (tid == 0) DRG_STAMP(702);
    const unsigned int cnt = min(cand_n_s, (unsigned int)CL_CAP);
    if (cnt == 0u && tid == 0) p.col.cand_seg[(size_t)b * NUM_SMS + g] = make_uint2(0u, 0u);
    if (cnt) {
      if (tid == 0) {
        cand_base_s = atomicAdd(&p.col.state[b].n_cand, cnt);
        p.col.cand_seg[(size_t)b * NUM_SMS + g] = make_uint2(cand_base_s, cnt);  // this CTA's rows, one contiguous segment
      }
      __syncthreads();
      const unsigned int base = cand_base_s;
      for (unsigned int e = tid; e < cnt; e += P2_THREADS) {
        const unsigned int k32 = tail_s[e];
        gkey[base + e] = k32;
        gidx[base + e] = tail_s[CL_CAP + e];
        atomicAdd(&p.col.cand_hist[(size_t)b * TK_BINS + cand_bin(k32, kmin, hist_sh)], 1u);
      }
    }
    if (tid == 0) DRG_STAMP(703);
  }
#undef DRG_STAMP
}

// ---------------------------------------------------------------------------------------
// merge the per-CTA column partials into v (and the dustbin column entry v_M)
//   block = 32 columns x 8 slices of the G partials; every thread first pulls its <= 19
//   partials into registers (all loads in flight at once), reduces them with one ex2 each,
//   then the 8 slices are merged through shared memory.
// ---------------------------------------------------------------------------------------
constexpr int COL_SLICES = 8;
constexpr int COL_MAXG = (NUM_SMS + COL_SLICES - 1) / COL_SLICES;  // 19

__global__ void __launch_bounds__(256) skh_col_kernel(const SkhParams p) {
  const int b = blockIdx.y;
  const int jj = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + jj;
  const int M = p.M, N = p.N, G = p.G;
  __shared__ float2 part_s[COL_SLICES][32];

  LseAcc a = lse_empty();
  const bool in_range = (j <= M) && !(p.dual && j == M);
  if (in_range) {
    const bool is_bin = (j == M);
    const bool col_ok = is_bin || p.dual || !p.apply_mask || p.tgt_mask[(size_t)b * M + j];
    if (col_ok) {
      const float2* src = is_bin ? (p.upart + (size_t)b * G) : (p.colpart + (size_t)b * G * M + j);
      const size_t gstride = is_bin ? 1 : (size_t)M;
      float2 q[COL_MAXG];
      float m = NEG_BIG;
#pragma unroll
      for (int k = 0; k < COL_MAXG; ++k) {
        const int g = sl + COL_SLICES * k;
        q[k] = (g < G) ? src[(size_t)g * gstride] : make_float2(NEG_BIG, 0.f);
        m = fmaxf(m, q[k].x);
      }
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < COL_MAXG; ++k) sum += q[k].y * ex2(q[k].x - m);
      a.m = m;
      a.s = sum;
    }
  }
  part_s[sl][jj] = make_float2(a.m, a.s);
  __syncthreads();
  if (sl != 0 || !in_range) return;

  float m = part_s[0][jj].x;
#pragma unroll
  for (int k = 1; k < COL_SLICES; ++k) m = fmaxf(m, part_s[k][jj].x);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < COL_SLICES; ++k) sum += part_s[k][jj].y * ex2(part_s[k][jj].x - m);
  a.m = m;
  a.s = sum;

  if (p.shard_partial) {  // row-sharded Sinkhorn: the ranks all-reduce these before any of them updates v
    p.shard_partial[(size_t)b * (M + 1) + j] = make_float2(a.m, a.s);
    return;
  }
  float* v_b = p.v + (size_t)b * p.ldv;
  if (p.dual) {
    v_b[j] = -lse_value(a) * LN2;
    return;
  }
  const SkhConst bc = p.bc[b];
  const float alpha = *p.alpha;
  const float uN = p.u[(size_t)b * p.ldu + N];
  if (j < M) {
    lse_add_value(a, (alpha + uN) * LOG2E);  // dustbin row entry
    v_b[j] = bc.norm - lse_value(a) * LN2;
  } else {
    lse_add_value(a, uN * LOG2E);  // c_M = alpha + LSE(u[0..N])
    v_b[M] = bc.log_nu_bin - (alpha + lse_value(a) * LN2);
  }
}

// ---------------------------------------------------------------------------------------
// row-sharded Sinkhorn (rows of one matrix spread over several GPUs, BASELINE.json configs[4])
//   begin : constants from the GLOBAL valid counts, u = v = 0
//   local : skh_iter*_kernel + skh_col_kernel(shard_partial) -> this rank's (max, sum) per column
//   [caller: all-reduce the partials over the ranks -- log-sum-exp is associative]
//   update: v from the reduced partials + the dustbin row term (identical on every rank)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) skh_shard_begin_kernel(const int* __restrict__ global_counts, SkhConst* __restrict__ bc,
                                                               float* __restrict__ v, float* __restrict__ u, int ldu, int ldv) {
  const int b = blockIdx.x;
  float4* v4 = reinterpret_cast<float4*>(v + (size_t)b * ldv);
  float4* u4 = reinterpret_cast<float4*>(u + (size_t)b * ldu);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = threadIdx.x; j < (ldv >> 2); j += blockDim.x) v4[j] = z4;
  for (int i = threadIdx.x; i < (ldu >> 2); i += blockDim.x) u4[i] = z4;
  if (threadIdx.x == 0) {
    const float ms = (float)global_counts[2 * b], ns = (float)global_counts[2 * b + 1];
    SkhConst c;
    c.norm = -logf(ms + ns);
    c.log_mu_bin = logf(ns) + c.norm;
    c.log_nu_bin = logf(ms) + c.norm;
    c.pad = 0.f;
    bc[b] = c;
  }
}

__global__ void __launch_bounds__(256) skh_shard_update_kernel(const SkhParams p, const float2* __restrict__ reduced) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int M = p.M, N = p.N;
  if (j > M) return;
  const float2 r = reduced[(size_t)b * (M + 1) + j];
  LseAcc a{r.x, r.y};
  const SkhConst bc = p.bc[b];
  const float alpha = *p.alpha;
  const float uN = p.u[(size_t)b * p.ldu + N];
  float* v_b = p.v + (size_t)b * p.ldv;
  if (j < M) {
    lse_add_value(a, (alpha + uN) * LOG2E);  // dustbin row entry (the dustbin row exists once, on every rank alike)
    v_b[j] = bc.norm - lse_value(a) * LN2;
  } else {
    lse_add_value(a, uN * LOG2E);
    v_b[M] = bc.log_nu_bin - (alpha + lse_value(a) * LN2);
  }
}

// Fused column merge + peer-to-peer all-reduce + potential update of the row-sharded path: ONE kernel per iteration
// instead of skh_col_kernel -> NCCL all-reduce (MAX, SUM) -> skh_shard_update_kernel.  A CTA owns 32 columns: it
// merges this rank's G per-CTA partials, stores the (max, sum) pair straight into every rank's inbox over NVLink
// (peer mappings from CUDA IPC), publishes them with a system-scope release on one flag per (rank, column chunk), waits
// until all ranks have published the same chunk, combines the P partials from its own inbox in rank order (every
// rank computes bit-identical potentials) and writes v.  CTAs depend only on the same-index CTA of the peers.
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void __launch_bounds__(256) skh_shard_exchange_kernel(const SkhParams p, const P2PView c) {
  const int b = blockIdx.y;
  const int jj = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + jj;
  const int M = p.M, N = p.N, G = p.G;
  __shared__ float2 part_s[COL_SLICES][32];

  LseAcc a = lse_empty();
  const bool in_range = (j <= M);
  if (in_range) {
    const bool is_bin = (j == M);
    const bool col_ok = is_bin || !p.apply_mask || p.tgt_mask[(size_t)b * M + j];
    if (col_ok) {
      const float2* src = is_bin ? (p.upart + (size_t)b * G) : (p.colpart + (size_t)b * G * M + j);
      const size_t gstride = is_bin ? 1 : (size_t)M;
      float2 q[COL_MAXG];
      float m = NEG_BIG;
#pragma unroll
      for (int k = 0; k < COL_MAXG; ++k) {
        const int g = sl + COL_SLICES * k;
        q[k] = (g < G) ? src[(size_t)g * gstride] : make_float2(NEG_BIG, 0.f);
        m = fmaxf(m, q[k].x);
      }
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < COL_MAXG; ++k) sum += q[k].y * ex2(q[k].x - m);
      a.m = m;
      a.s = sum;
    }
  }
  part_s[sl][jj] = make_float2(a.m, a.s);
  __syncthreads();
  const size_t eoff = (size_t)b * (M + 1) + j;
  if (sl == 0 && in_range) {
    float m = part_s[0][jj].x;
#pragma unroll
    for (int k = 1; k < COL_SLICES; ++k) m = fmaxf(m, part_s[k][jj].x);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < COL_SLICES; ++k) sum += part_s[k][jj].y * ex2(part_s[k][jj].x - m);
    const float2 mine = make_float2(m, sum);
    for (int r = 0; r < c.world; ++r)  // this rank's slot in every inbox (its own included)
      c.inbox[r][((size_t)c.slot * c.world + c.rank) * c.slot_elems + eoff] = mine;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // the stores above (ordered before this thread by the barrier) become visible to the peers first
    const size_t fidx = (size_t)c.slot * c.nflags + (size_t)b * gridDim.x + blockIdx.x;
    for (int r = 0; r < c.world; ++r) atomicAdd_system(c.flags[r] + fidx, 1u);
    const unsigned int* mine = c.flags[c.rank] + fidx;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys_u32(mine) < c.target) {
      if (global_timer_ns() - t0 > 4000000000ull) {  // a peer never arrived: report instead of hanging the GPU
        *c.status = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (sl != 0 || !in_range) return;
  LseAcc tot = lse_empty();
  for (int r = 0; r < c.world; ++r) {
    const float2 q = __ldcg(c.inbox[c.rank] + ((size_t)c.slot * c.world + r) * c.slot_elems + eoff);  // written by peers: L2
    lse_merge(tot, q.x, q.y);
  }
  const SkhConst bc = p.bc[b];
  const float alpha = *p.alpha;
  const float uN = p.u[(size_t)b * p.ldu + N];
  float* v_b = p.v + (size_t)b * p.ldv;
  if (j < M) {
    lse_add_value(tot, (alpha + uN) * LOG2E);  // dustbin row entry (the dustbin row exists once, on every rank alike)
    v_b[j] = bc.norm - lse_value(tot) * LN2;
  } else {
    lse_add_value(tot, uN * LOG2E);
    v_b[M] = bc.log_nu_bin - (alpha + lse_value(tot) * LN2);
  }
}

// ---------------------------------------------------------------------------------------
// final pass
// ---------------------------------------------------------------------------------------
struct SkhFinalParams {
  const float* scores;
  const uint8_t* src_mask;
  const uint8_t* tgt_mask;
  const float* alpha;
  const float* shift;
  const float* u;
  const float* v;
  const SkhConst* bc;
  int B, N, M;
  int ldu, ldv;
  int apply_mask;
  int mode;  // DRG_OUT_* ; 100 = dual softmax product
  float* out;
  const float* x_t;
  const float* xt_shift;
  const float* noise;
  float* conf;
  float k_x0, k_xt, sigma;
  float* x_min;
  float inv_temp;  // dual softmax
  int gen_noise;   // 1: N(0,1) draws come from the in-kernel Philox stream (noise == NULL)
  unsigned long long noise_seed, noise_offset;
  const unsigned long long* noise_offset_dev;
  unsigned long long* rowbest;  // optional [B,N]: packed (ordered(conf) << 32 | ~column) of every row's best entry
  unsigned long long* colbest;  // optional [B,M]: packed (ordered(conf) << 32 | ~row) of every column's best entry
  int tile_rows;                // rows per CTA of skh_final_tile_kernel (chosen by the host: one full wave of CTAs)
  float best_floor;             // only confidences > best_floor enter rowbest / colbest (-1: all of them)
};

// full (N+1)x(M+1) log-assignment (API parity with log_optimal_transport); odd row pitch -> scalar IO
__global__ void __launch_bounds__(256) skh_final_full_kernel(const SkhFinalParams p) {
  const int b = blockIdx.y;
  const int N = p.N, M = p.M;
  const SkhConst bc = p.bc[b];
  const float alpha = *p.alpha;
  const float shift = p.shift ? *p.shift : 0.f;
  const float* u_b = p.u + (size_t)b * p.ldu;
  const float* v_b = p.v + (size_t)b * p.ldv;
  for (int i = blockIdx.x; i <= N; i += gridDim.x) {
    const float ui = u_b[i];
    const bool row_ok = (i < N) && (!p.apply_mask || p.src_mask[(size_t)b * N + i]);
    const float* z = p.scores + ((size_t)b * N + i) * M;
    float* o = p.out + ((size_t)b * (N + 1) + i) * (M + 1);
    for (int j = threadIdx.x; j <= M; j += blockDim.x) {
      float zz;
      if (i < N && j < M) {
        zz = z[j] - shift;
        if (p.apply_mask && !(row_ok && p.tgt_mask[(size_t)b * M + j])) zz = -INFINITY;
      } else {
        zz = alpha;
      }
      o[j] = ((zz + ui) + v_b[j]) - bc.norm;  // same association as matching.py:34-36
    }
  }
}

// N x M outputs: conf / DDIM / dual-softmax product.  VEC: M % 4 == 0 and 16-byte aligned bases.
template <bool VEC>
__global__ void __launch_bounds__(256) skh_final_kernel(const SkhFinalParams p) {
  const int b = blockIdx.y;
  const int N = p.N, M = p.M;
  const SkhConst bc = p.bc[b];
  const float shift = p.shift ? *p.shift : 0.f;
  const float* u_b = p.u + (size_t)b * p.ldu;
  const float* v_b = p.v + (size_t)b * p.ldv;
  const bool dual = (p.mode == 100);
  const bool ddim = (p.mode == DRG_OUT_DDIM);
  const float xt_shift = p.xt_shift ? *p.xt_shift : 0.f;
  const unsigned long long noise_offset = p.noise_offset + (p.noise_offset_dev ? *p.noise_offset_dev : 0ull);
  float local_min = INFINITY;

  auto one = [&](float z, float ui, float vj, bool ok, float xt, float nz, float& conf_out) -> float {
    float conf;
    if (dual) {
      conf = ok ? ex2((2.f * z * p.inv_temp + ui + vj) * LOG2E) : 0.f;
      conf_out = conf;
      return conf;
    }
    const float zz = ok ? (z - shift) : -INFINITY;
    const float la = ((zz + ui) + vj) - bc.norm;
    conf = ex2(la * LOG2E);
    conf_out = conf;
    if (!ddim) return conf;
    float xn = ok ? fmaf(p.k_x0, conf, fmaf(p.k_xt, xt - xt_shift, p.sigma * nz)) : -INFINITY;
    if (ok && xn > -INFINITY) local_min = fminf(local_min, xn);
    return xn;
  };

  for (int i = blockIdx.x; i < N; i += gridDim.x) {
    const float ui = u_b[i];
    const bool row_ok = (!p.apply_mask && !dual) || p.src_mask[(size_t)b * N + i];
    const size_t base = ((size_t)b * N + i) * M;
    if constexpr (VEC) {
      for (int j = 4 * threadIdx.x; j < M; j += 4 * blockDim.x) {
        const float4 z = *reinterpret_cast<const float4*>(p.scores + base + j);
        const float4 vj = *reinterpret_cast<const float4*>(v_b + j);
        bool ok[4] = {row_ok, row_ok, row_ok, row_ok};
        if (p.apply_mask || dual) {
          const uchar4 tm = *reinterpret_cast<const uchar4*>(p.tgt_mask + (size_t)b * M + j);
          ok[0] = row_ok && tm.x;
          ok[1] = row_ok && tm.y;
          ok[2] = row_ok && tm.z;
          ok[3] = row_ok && tm.w;
        }
        float4 xt = make_float4(0.f, 0.f, 0.f, 0.f), nz = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ddim) {
          xt = *reinterpret_cast<const float4*>(p.x_t + base + j);
          if (p.noise) nz = *reinterpret_cast<const float4*>(p.noise + base + j);
          else if (p.gen_noise) nz = philox_normal4((unsigned long long)((base + j) >> 2), noise_offset, p.noise_seed);
        }
        float4 o, c;
        o.x = one(z.x, ui, vj.x, ok[0], xt.x, nz.x, c.x);
        o.y = one(z.y, ui, vj.y, ok[1], xt.y, nz.y, c.y);
        o.z = one(z.z, ui, vj.z, ok[2], xt.z, nz.z, c.z);
        o.w = one(z.w, ui, vj.w, ok[3], xt.w, nz.w, c.w);
        *reinterpret_cast<float4*>(p.out + base + j) = o;
        if (ddim && p.conf) *reinterpret_cast<float4*>(p.conf + base + j) = c;
      }
    } else {
      for (int j = threadIdx.x; j < M; j += blockDim.x) {
        bool ok = row_ok;
        if (p.apply_mask || dual) ok = row_ok && p.tgt_mask[(size_t)b * M + j];
        const float xt = ddim ? p.x_t[base + j] : 0.f;
        float nz = (ddim && p.noise) ? p.noise[base + j] : 0.f;
        if (ddim && !p.noise && p.gen_noise) nz = philox_normal4((unsigned long long)(base + j), noise_offset, p.noise_seed).x;
        float c;
        p.out[base + j] = one(p.scores[base + j], ui, v_b[j], ok, xt, nz, c);
        if (ddim && p.conf) p.conf[base + j] = c;
      }
    }
  }
  if (ddim && p.x_min) {
    local_min = warp_min(local_min);
    if ((threadIdx.x & 31) == 0 && local_min < INFINITY) atomic_min_float(p.x_min, local_min);
  }
}

// ---------------------------------------------------------------------------------------
// tiled final pass (rows 16-byte aligned): CTA = 32 rows x 1024 columns, thread = one column quad.
//   v, the target mask and the column bests of the quad stay in registers down the band of rows; the row best is a
//   warp max + ballot per row.  Writes conf and / or the DDIM update like skh_final_kernel and, when asked, the
//   packed row / column arg-max keys that the correspondence extraction needs (so that x0 = conf never has to be
//   re-read -- or even stored -- to find the mutual matches; Matching.get_match, matching.py:71-88).
// ---------------------------------------------------------------------------------------
constexpr int FT_THREADS = 256;
constexpr int FT_CTAS_PER_SM = 4;  // 256 threads x <= 64 registers
// Rows per CTA.  Measured at 4096^2 (round-1 tuning build): 8 and 16 rows 49 us, 24 rows 54 us, 28 rows (one
// full wave of 588 resident CTAs) 52 us, 32+ rows 65 us -- many small CTAs balance better than one exact wave.  Every
// thread requests the next row's scores / x_t before it works on the current one.
constexpr int FT_ROWS = 16;

template <bool MASKED, bool TRACK, bool WANT_MIN>
__device__ __forceinline__ void final_tile_body(const SkhFinalParams& p) {
  const int b = blockIdx.z;
  const int N = p.N, M = p.M;
  const int c = blockIdx.x * (FT_THREADS * 4) + 4 * (int)threadIdx.x;
  const bool active = c < M;
  const int lane = threadIdx.x & 31;
  const int i0 = blockIdx.y * p.tile_rows, i1 = min(N, i0 + p.tile_rows);
  const SkhConst bc = p.bc[b];
  const float shift = p.shift ? *p.shift : 0.f;
  const float xt_shift = p.xt_shift ? *p.xt_shift : 0.f;
  const bool ddim = (p.mode == DRG_OUT_DDIM);
  const unsigned long long noise_offset = p.noise_offset + (p.noise_offset_dev ? *p.noise_offset_dev : 0ull);
  const float* u_b = p.u + (size_t)b * p.ldu;
  float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
  bool tm[4] = {true, true, true, true};
  if (active) {
    v4 = *reinterpret_cast<const float4*>(p.v + (size_t)b * p.ldv + c);
    if (MASKED) {
      const uchar4 t4 = *reinterpret_cast<const uchar4*>(p.tgt_mask + (size_t)b * M + c);
      tm[0] = t4.x; tm[1] = t4.y; tm[2] = t4.z; tm[3] = t4.w;
    }
  }
  const float vj[4] = {v4.x, v4.y, v4.z, v4.w};
  // column bests start at the floor: only entries above it are ever recorded (floor -1: every entry, confidences are >= 0)
  const float floor_v = p.best_floor;
  float cbv[4] = {floor_v, floor_v, floor_v, floor_v};
  int cbi[4] = {0, 0, 0, 0};
  float local_min = INFINITY;
  // software prefetch: the next row's operands are requested before the current row is processed
  float4 z_nx = make_float4(0.f, 0.f, 0.f, 0.f), t_nx = make_float4(0.f, 0.f, 0.f, 0.f);
  float u_nx = 0.f;
  auto fetch_row = [&](int i) {
    if (i < i1) {
      u_nx = u_b[i];
      if (active) {
        const size_t base = ((size_t)b * N + i) * M + c;
        z_nx = __ldcs(reinterpret_cast<const float4*>(p.scores + base));
        if (ddim) t_nx = __ldcs(reinterpret_cast<const float4*>(p.x_t + base));
      }
    }
  };
  fetch_row(i0);
  for (int i = i0; i < i1; ++i) {
    const float ui = u_nx;
    const float4 z4 = z_nx, t4 = t_nx;
    fetch_row(i + 1);
    const bool row_ok = !MASKED || p.src_mask[(size_t)b * N + i];
    float cf[4] = {-1.f, -1.f, -1.f, -1.f};
    if (active) {
      const size_t base = ((size_t)b * N + i) * M + c;
      const float z[4] = {z4.x, z4.y, z4.z, z4.w};
      float xt[4] = {0.f, 0.f, 0.f, 0.f}, nz[4] = {0.f, 0.f, 0.f, 0.f};
      if (ddim) {
        xt[0] = t4.x; xt[1] = t4.y; xt[2] = t4.z; xt[3] = t4.w;
        if (p.noise) {
          const float4 n4 = *reinterpret_cast<const float4*>(p.noise + base);
          nz[0] = n4.x; nz[1] = n4.y; nz[2] = n4.z; nz[3] = n4.w;
        } else if (p.gen_noise) {
          const float4 n4 = philox_normal4((unsigned long long)(base >> 2), noise_offset, p.noise_seed);
          nz[0] = n4.x; nz[1] = n4.y; nz[2] = n4.z; nz[3] = n4.w;
        }
      }
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = !MASKED || (row_ok && tm[e]);
        const float zz = ok ? (z[e] - shift) : -INFINITY;
        const float la = ((zz + ui) + vj[e]) - bc.norm;  // same association as matching.py:34-36
        cf[e] = ex2(la * LOG2E);
        if (ddim) {
          const float xn = ok ? fmaf(p.k_x0, cf[e], fmaf(p.k_xt, xt[e] - xt_shift, p.sigma * nz[e])) : -INFINITY;
          if (WANT_MIN && ok && xn > -INFINITY) local_min = fminf(local_min, xn);
          o[e] = xn;
        } else {
          o[e] = cf[e];
        }
      }
      *reinterpret_cast<float4*>(p.out + base) = make_float4(o[0], o[1], o[2], o[3]);
      if (ddim && p.conf) *reinterpret_cast<float4*>(p.conf + base) = make_float4(cf[0], cf[1], cf[2], cf[3]);
    }
    // The arg-max bookkeeping below (~47 instructions per quad) only matters for entries above the floor; with the
    // matcher's confidence threshold as the floor at most a handful of the 32 warps of a row hold one (a row of the
    // transport plan sums to <= 1), so one 4-way max, one compare and one vote skip it for almost every warp-row.
    if (TRACK && __any_sync(0xffffffffu, fmaxf(fmaxf(cf[0], cf[1]), fmaxf(cf[2], cf[3])) > floor_v)) {
      // column bests (rows ascend, strict > keeps the lowest row on ties)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (cf[e] > cbv[e]) {
          cbv[e] = cf[e];
          cbi[e] = i;
        }
      // row best of this warp's 128 columns: confidences are >= 0, so their bit patterns order like unsigned integers
      // and ONE redux.sync finds the warp maximum; the owner is the lowest lane holding it (lowest column)
      float rv = cf[0];
      int re = 0;
#pragma unroll
      for (int e = 1; e < 4; ++e)
        if (cf[e] > rv) {
          rv = cf[e];
          re = e;
        }
      const unsigned int rbits = (active && rv >= 0.f && rv > floor_v) ? __float_as_uint(rv) : 0u;  // NaN / inactive lanes never win
      const unsigned int wbits = __reduce_max_sync(0xffffffffu, rbits);
      const unsigned int owners = __ballot_sync(0xffffffffu, active && rbits == wbits);
      if (owners && lane == __ffs(owners) - 1) {
        const unsigned long long key =
            ((unsigned long long)float_to_ordered(rv) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)(c + re));
        atomicMax(&p.rowbest[(size_t)b * N + i], key);
      }
    }
  }
  if (TRACK && active && i1 > i0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (!(cbv[e] > floor_v)) continue;  // nothing above the floor in this column band: the key stays 0
      const unsigned long long key =
          ((unsigned long long)float_to_ordered(cbv[e]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)cbi[e]);
      atomicMax(&p.colbest[(size_t)b * M + c + e], key);
    }
  }
  if (WANT_MIN && ddim) {
    local_min = warp_min(local_min);
    if (lane == 0 && local_min < INFINITY) atomic_min_float(p.x_min, local_min);
  }
}

// The mask predicates, the arg-max tracking and the running minimum are compiled out when not needed: the pass is
// issue-bound (in-kernel Philox + Box-Muller), not bandwidth-bound, so every instruction per element counts.
__global__ void __launch_bounds__(FT_THREADS, FT_CTAS_PER_SM) skh_final_tile_kernel(const SkhFinalParams p) {
  const bool masked = p.apply_mask && p.bc[blockIdx.z].pad != 1.f;  // pad == 1: the Sinkhorn saw no padded row / column
  const bool track = p.rowbest != nullptr;
  const bool want_min = p.x_min != nullptr;
  if (masked) {
    if (track) {
      if (want_min) final_tile_body<true, true, true>(p);
      else final_tile_body<true, true, false>(p);
    } else {
      if (want_min) final_tile_body<true, false, true>(p);
      else final_tile_body<true, false, false>(p);
    }
  } else {
    if (track) {
      if (want_min) final_tile_body<false, true, true>(p);
      else final_tile_body<false, true, false>(p);
    } else {
      if (want_min) final_tile_body<false, false, true>(p);
      else final_tile_body<false, false, false>(p);
    }
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct SkhPlan {
  int R, KQ, nstage, G;
  size_t smem;
  bool ok;
};

static SkhPlan make_plan(int B, int N, int M) {
  SkhPlan pl{};
  pl.ok = false;
  if (M < 1 || M > SKH_MAX_M || N < 1) return pl;
  int R = 16;
  while (R > 1 && (long long)R * M > SKH_STAGE_FLOATS) R >>= 1;
  pl.R = R;
  pl.KQ = (M <= 2048) ? 1 : (M <= 4096) ? 2 : (M <= 8192) ? 4 : 8;
  const size_t Mv = (size_t)((M + 1 + 3) & ~3);
  const size_t stage_floats = (size_t)(((R * M) + 3) & ~3);
  const int SEG = SKH_WARPS / R;
  const size_t fixed = Mv * 4 + (size_t)R * SEG * 8 + (size_t)R * 4 + 2 * SKH_WARPS * 4 + 8 /*align*/ + 4 * 8 /*bars*/ + 64;
  int nstage = (int)((SKH_SMEM_LIMIT - fixed) / (stage_floats * 4));
  if (nstage > 4) nstage = 4;
  if (nstage < 1) return pl;
  pl.nstage = nstage;
  pl.smem = fixed + (size_t)nstage * stage_floats * 4;
  const int nslab = (N + R - 1) / R;
  int G = NUM_SMS / (B < 1 ? 1 : B);
  if (G < 1) G = 1;
  if (G > nslab) G = nslab;
  pl.G = G;
  pl.ok = true;
  return pl;
}

struct SkhWorkspace {
  SkhConst* bc;
  float* u;
  float* v;
  float2* colpart;
  float2* upart;
  unsigned int* gsync;  // [B] grid-barrier counters of the persistent kernel
  size_t total;
};

static inline int pitch4(int n) { return (n + 3) & ~3; }

static SkhWorkspace carve(void* ws, int B, int N, int M, int G) {
  SkhWorkspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* ptr = ws ? (void*)((char*)ws + off) : nullptr;
    off += align_up(bytes, 256);
    return ptr;
  };
  w.bc = (SkhConst*)take(sizeof(SkhConst) * B);
  w.u = (float*)take(sizeof(float) * (size_t)B * pitch4(N + 1));
  w.v = (float*)take(sizeof(float) * (size_t)B * pitch4(M + 1));
  w.colpart = (float2*)take(sizeof(float2) * (size_t)B * G * M);
  w.upart = (float2*)take(sizeof(float2) * (size_t)B * G);
  w.gsync = (unsigned int*)take(sizeof(unsigned int) * (size_t)B * 3);  // [B] barrier counters + [B][2] dv slots
  w.total = off;
  return w;
}

template <int R, int KQ>
static cudaError_t launch_iter_rk(const SkhParams& p, const SkhPlan& pl, bool vec, cudaStream_t st) {
  dim3 grid(pl.G, p.B);
  cudaError_t e;
  if (vec) {
    e = cudaFuncSetAttribute(skh_iter_kernel<R, KQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    skh_iter_kernel<R, KQ, true><<<grid, SKH_THREADS, pl.smem, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(skh_iter_kernel<R, KQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    skh_iter_kernel<R, KQ, false><<<grid, SKH_THREADS, pl.smem, st>>>(p);
  }
  return cudaGetLastError();
}

static cudaError_t launch_iter(const SkhParams& p, const SkhPlan& pl, cudaStream_t st) {
  // unaligned rows (M % 4 != 0): the scalar cp.async variant
  switch (pl.R) {
    case 16: return launch_iter_rk<16, 1>(p, pl, false, st);
    case 8: return launch_iter_rk<8, 1>(p, pl, false, st);
    case 4: return launch_iter_rk<4, 2>(p, pl, false, st);
    case 2: return launch_iter_rk<2, 4>(p, pl, false, st);
    default: return launch_iter_rk<1, 8>(p, pl, false, st);
  }
}

// ---- plan / launch of the software-pipelined kernel (aligned rows)
struct SkhPlan2 {
  int R, KQ, NCH, nstage, G;
  bool full;
  size_t smem;
  bool ok;
};

static SkhPlan2 make_plan2(int B, int N, int M) {
  SkhPlan2 pl{};
  pl.ok = false;
  if (M < 4 || M > SKH_MAX_M || N < 1 || (M % 4) != 0) return pl;
  const int target = SKH_STAGE_FLOATS;
  int R = 16;
  while (R > 1 && (long long)R * M > target) R >>= 1;
  pl.R = R;
  pl.KQ = (M <= 2048) ? 1 : (M <= 4096) ? 2 : (M <= 8192) ? 4 : 8;
  const int SEG = SKH_WARPS / R;
  pl.full = false;
  pl.NCH = 8;
  if (M >= 2048 && (M & (M - 1)) == 0 && M % (SEG * 128) == 0) {
    const int nch = M / (SEG * 128);
    if (nch == 1 || nch == 2 || nch == 4 || nch == 8) {
      pl.full = true;
      pl.NCH = nch;
    }
  }
  if ((long long)SEG * 1024 < M) return pl;  // a warp segment holds at most 1024 columns
  const size_t Mv = (size_t)((M + 1 + 3) & ~3);
  const size_t stage_bytes = (size_t)R * M * 4;
  const size_t fixed = Mv * 4 + 2 * SKH_WARPS * 8 + 2 * SKH_WARPS * 4 + 8 * 8 + 64;
  int nstage = (int)((SKH_SMEM_LIMIT - fixed) / stage_bytes);
  if (nstage > 8) nstage = 8;
  if (nstage < 2) return pl;
  pl.nstage = nstage;
  pl.smem = fixed + (size_t)nstage * stage_bytes;
  const int nslab = (N + R - 1) / R;
  int G = NUM_SMS / (B < 1 ? 1 : B);
  if (G < 1) G = 1;
  if (G > nslab) G = nslab;
  pl.G = G;
  pl.ok = true;
  return pl;
}

template <int R, int KQ, int NCH, bool FULL>
static cudaError_t launch_iter2_t(const SkhParams& p, const SkhPlan2& pl, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(skh_iter2_kernel<R, KQ, NCH, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return e;
  skh_iter2_kernel<R, KQ, NCH, FULL><<<dim3(pl.G, p.B), SKH_THREADS, pl.smem, st>>>(p);
  return cudaGetLastError();
}

static cudaError_t launch_iter2(const SkhParams& p, const SkhPlan2& pl, cudaStream_t st) {
#define DRG_SKH_CASE(r, kq, nch, fl) \
  if (pl.R == r && pl.KQ == kq && pl.NCH == nch && pl.full == fl) return launch_iter2_t<r, kq, nch, fl>(p, pl, st);
  // power-of-two widths: no bounds checks
  DRG_SKH_CASE(8, 1, 8, true)   // M = 2048, 64 KB stages
  DRG_SKH_CASE(4, 2, 8, true)   // M = 4096, 64 KB stages
  DRG_SKH_CASE(2, 4, 8, true)   // M = 8192, 64 KB stages
  DRG_SKH_CASE(1, 8, 8, true)   // M = 16384
  // everything else: guarded
  DRG_SKH_CASE(16, 1, 8, false)
  DRG_SKH_CASE(8, 1, 8, false)
  DRG_SKH_CASE(4, 1, 8, false)
  DRG_SKH_CASE(2, 1, 8, false)
  DRG_SKH_CASE(1, 1, 8, false)
  DRG_SKH_CASE(4, 2, 8, false)
  DRG_SKH_CASE(2, 2, 8, false)
  DRG_SKH_CASE(1, 2, 8, false)
  DRG_SKH_CASE(2, 4, 8, false)
  DRG_SKH_CASE(1, 4, 8, false)
  DRG_SKH_CASE(1, 8, 8, false)
#undef DRG_SKH_CASE
  return cudaErrorInvalidConfiguration;
}

// ---- register-slab persistent kernel: plan and launch
struct SkhPlanP2 {
  int KQ, RR, G, nstage;
  bool full, ok;
  size_t smem;
};

static SkhPlanP2 make_plan_p2(int B, int N, int M) {
  SkhPlanP2 pl{};
  pl.ok = false;
  if (M < 4 || M > 4096 || (M % 4) != 0 || N < 1 || B < 1 || B > NUM_SMS) return pl;
  pl.KQ = (M <= 1024) ? 1 : (M <= 2048) ? 2 : 4;
  pl.full = (M == 1024 * pl.KQ);
  const int gmax = NUM_SMS / B;
  const int rr_max = 8 / pl.KQ;
  const int rr_min = (pl.KQ == 1) ? 2 : 1;
  pl.RR = ((long long)P2_GROUPS * rr_max * gmax <= N) ? rr_max : rr_min;
  int G = (N + P2_GROUPS * pl.RR - 1) / (P2_GROUPS * pl.RR);  // at least one mini-slab per row group
  if (G > gmax) G = gmax;
  if (G < 1) G = 1;
  pl.G = G;
  if ((N + G - 1) / G > 1024) return pl;  // row references of a CTA live in a 1024-entry shared array
  const size_t Mv = (size_t)((M + 1 + 3) & ~3);
  const size_t fixed = (Mv + 1024 + 2 * P2_GROUPS * 64 + 64 + 8 + (P2_THREADS / 32) * 32 * 2 + TK_BINS) * 4 + P2_MAX_STAGES * 8;
  const size_t stage_bytes = (size_t)pl.RR * M * 4;
  const size_t xcomb_bytes = (size_t)(P2_GROUPS - 1) * P2_TPR * pl.KQ * 4 * 8;
  int nstage = (int)((SKH_SMEM_LIMIT - 256 - fixed) / stage_bytes);
  if (nstage > P2_MAX_STAGES) nstage = P2_MAX_STAGES;
  if (nstage < 2) return pl;
  pl.nstage = nstage;
  const size_t ring_bytes = (size_t)nstage * stage_bytes;
  pl.smem = fixed + (ring_bytes > xcomb_bytes ? ring_bytes : xcomb_bytes);
  pl.ok = true;
  return pl;
}

template <int KQ, int RR, bool FULL>
static cudaError_t launch_persist2_t(const SkhParams& p, const SkhPlanP2& pl, int iters, unsigned int* gsync, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(skh_persist2_kernel<KQ, RR, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return e;
  SkhParams pp = p;
  pp.nstage = pl.nstage;
  void* args[] = {(void*)&pp, (void*)&iters, (void*)&gsync};
  return cudaLaunchCooperativeKernel((const void*)skh_persist2_kernel<KQ, RR, FULL>, dim3(pl.G, p.B), dim3(P2_THREADS), args, pl.smem, st);
}

static cudaError_t launch_persist2(const SkhParams& p, const SkhPlanP2& pl, int iters, unsigned int* gsync, cudaStream_t st) {
#define DRG_P2_CASE(kq, rr)                                                                  \
  if (pl.KQ == kq && pl.RR == rr) {                                                          \
    if (pl.full) return launch_persist2_t<kq, rr, true>(p, pl, iters, gsync, st);            \
    return launch_persist2_t<kq, rr, false>(p, pl, iters, gsync, st);                        \
  }
  DRG_P2_CASE(4, 2) DRG_P2_CASE(4, 1) DRG_P2_CASE(2, 4) DRG_P2_CASE(2, 1) DRG_P2_CASE(1, 8) DRG_P2_CASE(1, 2)
#undef DRG_P2_CASE
  return cudaErrorInvalidConfiguration;
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace drg

using namespace drg;

static int skh_max_g(int B) {
  int G = NUM_SMS / (B < 1 ? 1 : B);
  return G < 1 ? 1 : G;
}

extern "C" size_t drg_sinkhorn_workspace_bytes(int B, int N, int M) {
  if (B < 1 || N < 1 || M < 1) return 0;
  SkhPlan pl = make_plan(B, N, M);
  if (!pl.ok) return 0;
  return carve(nullptr, B, N, M, skh_max_g(B)).total;
}


enum SkhRun { SKH_RUN_ALL = 0, SKH_SHARD_BEGIN, SKH_SHARD_LOCAL, SKH_SHARD_UPDATE, SKH_SHARD_FINAL, SKH_SHARD_LOCAL_X };

static int run_sinkhorn(const drg_sinkhorn_args* a, bool dual, float temperature, void* workspace, size_t workspace_bytes,
                        void* stream, SkhRun run = SKH_RUN_ALL, const int* global_counts = nullptr, float2* shard_partial = nullptr,
                        const P2PView* xview = nullptr, const SkhCollect* collect = nullptr, bool* collected = nullptr) {
  cudaStream_t st = (cudaStream_t)stream;
  const int B = a->B, N = a->N, M = a->M;
  SkhPlan pl = make_plan(B, N, M);
  if (!pl.ok) {
    set_error("sinkhorn: unsupported shape B=%d N=%d M=%d (need 1 <= M <= %d)", B, N, M, SKH_MAX_M);
    return DRG_ERR_UNSUPPORTED;
  }
  // with R rows per stage the column accumulators need KQ*2048 >= M
  if (pl.KQ * 2048 < M) {
    set_error("sinkhorn: internal plan error");
    return DRG_ERR_UNSUPPORTED;
  }
  const bool vec = (M % 4 == 0) && aligned16(a->scores);
  SkhPlan2 pl2 = vec ? make_plan2(B, N, M) : SkhPlan2{};
  const bool use2 = vec && pl2.ok;
  SkhWorkspace w = carve(workspace, B, N, M, skh_max_g(B));
  if (workspace == nullptr || workspace_bytes < w.total) {
    set_error("sinkhorn: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return DRG_ERR_WORKSPACE;
  }
  if (((uintptr_t)workspace & 255u) != 0) {
    set_error("sinkhorn: workspace must be 256-byte aligned");
    return DRG_ERR_INVALID;
  }
  // all iterations in one cooperative launch (register-slab kernel) whenever the shape allows it; otherwise one launch
  // per iteration (skh_iter2_kernel for 16-byte aligned rows, skh_iter_kernel for the rest) + the column merge
  SkhPlanP2 plp = (run == SKH_RUN_ALL && vec && !dual && a->iters >= 1) ? make_plan_p2(B, N, M) : SkhPlanP2{};
  const bool persist2 = plp.ok;
  const bool persist = persist2;
  if (run == SKH_SHARD_BEGIN) {
    skh_shard_begin_kernel<<<B, 1024, 0, st>>>(global_counts, w.bc, w.v, w.u, pitch4(N + 1), pitch4(M + 1));
    DRG_LAUNCH_CHECK();
    return DRG_OK;
  }
  if (!persist && run == SKH_RUN_ALL) {
    ProfScope prof_scope(PROF_SKH_PREP, st);
    skh_prep_kernel<<<B, 1024, 0, st>>>(a->src_mask, a->tgt_mask, N, M, w.bc, w.v, w.u, pitch4(N + 1), pitch4(M + 1));
    DRG_LAUNCH_CHECK();
  }

  SkhParams p{};
  p.scores = a->scores;
  p.src_mask = a->src_mask;
  p.tgt_mask = a->tgt_mask;
  p.alpha = a->alpha;
  p.shift = a->shift;
  p.u = w.u;
  p.v = w.v;
  p.colpart = w.colpart;
  p.upart = w.upart;
  p.bc = w.bc;
  p.bc_out = w.bc;
  p.shard_partial = (run == SKH_SHARD_LOCAL) ? shard_partial : nullptr;
  p.B = B;
  p.N = N;
  p.M = M;
  p.G = persist2 ? plp.G : use2 ? pl2.G : pl.G;
  p.ldu = pitch4(N + 1);
  p.ldv = pitch4(M + 1);
  p.apply_mask = a->apply_mask;
  p.dual = dual ? 1 : 0;
  p.zscale2 = dual ? LOG2E / temperature : LOG2E;
  p.nstage = use2 ? pl2.nstage : pl.nstage;
  p.dbg_times = g_tuning_stamps;

  if (run == SKH_SHARD_UPDATE) {
    skh_shard_update_kernel<<<dim3((M + 1 + 255) / 256, B), 256, 0, st>>>(p, shard_partial);
    DRG_LAUNCH_CHECK();
    return DRG_OK;
  }
  const int iters = (run == SKH_SHARD_LOCAL || run == SKH_SHARD_LOCAL_X) ? 1 : run == SKH_SHARD_FINAL ? 0 : dual ? 1 : a->iters;
  dim3 cgrid((M + 1 + 31) / 32, B);
  bool bests_cleared = false;
  bool final_fused = false;
  if (persist) {
    if (persist2 && run == SKH_RUN_ALL && !dual && a->rowbest && a->colbest) {
      p.zero_a = a->rowbest;
      p.zero_a_n = (size_t)B * N;
      p.zero_b = a->colbest;
      p.zero_b_n = (size_t)B * M;
      bests_cleared = true;
    }
    if (collect && (long long)N * M < (1ll << 32)) {
      p.col = *collect;
      if (collected) *collected = true;
    }
    // The final pass as the last phase of the same launch (conf / DDIM outputs on 16-byte aligned buffers).  The in-kernel
    // arg-max keys go straight to rowbest / colbest with atomics, which is only cheap when a floor keeps them rare (a row of
    // the plan sums to <= 1: at most 1 / floor entries per row exceed it); without a floor the tiled final kernel runs.
    {
      const bool want_best = a->rowbest && a->colbest;
      const bool fuse = persist2 && run == SKH_RUN_ALL && !dual && !p.col.state &&
                        (a->out_mode == DRG_OUT_CONF || a->out_mode == DRG_OUT_DDIM) && aligned16(a->out) &&
                        (a->out_mode != DRG_OUT_DDIM || aligned16(a->x_t)) && (!a->noise || aligned16(a->noise)) &&
                        (!a->conf || aligned16(a->conf)) && (!want_best || (a->has_best_floor && a->best_floor >= 0.01f)) &&
                        a->x_min == nullptr;   // (the running minimum of the 3DMatch flavour stays with the stand-alone kernel)
      if (fuse) {
        p.fin.out = a->out;
        p.fin.ddim = a->out_mode == DRG_OUT_DDIM ? 1 : 0;
        p.fin.x_t = a->x_t;
        p.fin.xt_shift = a->xt_shift;
        p.fin.noise = p.fin.ddim ? a->noise : nullptr;
        p.fin.conf = p.fin.ddim ? a->conf : nullptr;
        p.fin.rowbest = want_best ? a->rowbest : nullptr;
        p.fin.colbest = want_best ? a->colbest : nullptr;
        p.fin.k_x0 = a->k_x0;
        p.fin.k_xt = a->k_xt;
        p.fin.sigma = a->sigma;
        p.fin.best_floor = a->best_floor;
        p.fin.gen_noise = (a->noise == nullptr && a->gen_noise) ? 1 : 0;
        for (int r = 0; r < 7; ++r) {
          p.fin.rk[2 * r] = (unsigned int)a->noise_seed + (unsigned int)r * 0x9E3779B9u;
          p.fin.rk[2 * r + 1] = (unsigned int)(a->noise_seed >> 32) + (unsigned int)r * 0xBB67AE85u;
        }
        p.fin.noise_offset = a->noise_offset;
        p.fin.noise_offset_dev = a->noise_offset_dev;
        final_fused = true;
      }
    }
    DRG_CUDA(cudaMemsetAsync(w.gsync, 0, sizeof(unsigned int) * B * 3, st));
    cudaError_t e;
    {
      // (slot "skh_col" is free in this mode: it times the launches that carry the candidate-search tail)
      ProfScope prof_scope(p.col.state ? PROF_SKH_COL : final_fused ? PROF_SKH_FUSED : PROF_SKH_ITER, st);
      e = launch_persist2(p, plp, iters, w.gsync, st);
    }
    if (e != cudaSuccess) {
      set_error("persistent sinkhorn launch failed: %s (smem=%zu, grid %d x %d)", cudaGetErrorString(e), plp.smem, plp.G, B);
      return DRG_ERR_CUDA;
    }
    count_launch();
  }
  for (int k = 0; k < (persist ? 0 : iters); ++k) {
    cudaError_t e;
    {
      ProfScope prof_scope(PROF_SKH_ITER, st);
      e = use2 ? launch_iter2(p, pl2, st) : launch_iter(p, pl, st);
    }
    if (e != cudaSuccess) {
      set_error("sinkhorn iteration launch failed: %s (smem=%zu)", cudaGetErrorString(e), use2 ? pl2.smem : pl.smem);
      return DRG_ERR_CUDA;
    }
    count_launch();
    {
      ProfScope prof_scope(PROF_SKH_COL, st);
      if (run == SKH_SHARD_LOCAL_X) skh_shard_exchange_kernel<<<cgrid, 256, 0, st>>>(p, *xview);
      else skh_col_kernel<<<cgrid, 256, 0, st>>>(p);
    }
    DRG_LAUNCH_CHECK();
  }

  if (run == SKH_SHARD_LOCAL || run == SKH_SHARD_LOCAL_X) return DRG_OK;
  if ((a->out_mode != DRG_OUT_NONE || dual) && !final_fused) {
    SkhFinalParams f{};
    f.scores = a->scores;
    f.src_mask = a->src_mask;
    f.tgt_mask = a->tgt_mask;
    f.alpha = a->alpha;
    f.shift = a->shift;
    f.u = w.u;
    f.v = w.v;
    f.bc = w.bc;
    f.B = B;
    f.N = N;
    f.M = M;
    f.ldu = pitch4(N + 1);
    f.ldv = pitch4(M + 1);
    f.apply_mask = a->apply_mask;
    f.mode = dual ? 100 : a->out_mode;
    f.out = a->out;
    f.x_t = a->x_t;
    f.xt_shift = a->xt_shift;
    f.noise = a->noise;
    f.conf = a->conf;
    f.k_x0 = a->k_x0;
    f.k_xt = a->k_xt;
    f.sigma = a->sigma;
    f.x_min = a->x_min;
    f.gen_noise = (a->noise == nullptr && a->gen_noise) ? 1 : 0;
    f.noise_seed = a->noise_seed;
    f.noise_offset = a->noise_offset;
    f.noise_offset_dev = a->noise_offset_dev;
    f.rowbest = (a->rowbest && a->colbest) ? a->rowbest : nullptr;
    f.colbest = f.rowbest ? a->colbest : nullptr;
    f.best_floor = a->has_best_floor ? a->best_floor : -1.f;
    f.inv_temp = dual ? 1.f / temperature : 0.f;
    int gx = (NUM_SMS * 8) / B;
    if (gx < 1) gx = 1;
    if (!dual && a->out_mode == DRG_OUT_LOG_FULL) {
      if (gx > N + 1) gx = N + 1;
      {
        ProfScope prof_scope(PROF_SKH_FINAL, st);
        skh_final_full_kernel<<<dim3(gx, B), 256, 0, st>>>(f);
      }
    } else {
      if (gx > N) gx = N;
      if (f.rowbest && !(vec && !dual && (a->out_mode == DRG_OUT_CONF || a->out_mode == DRG_OUT_DDIM))) {
        set_error("sinkhorn: rowbest/colbest need M % 4 == 0 and out_mode CONF or DDIM");
        return DRG_ERR_UNSUPPORTED;
      }
      bool fvec = vec && !dual && aligned16(a->out) && (!f.x_t || aligned16(f.x_t)) && (!f.noise || aligned16(f.noise)) &&
                  (!f.conf || aligned16(f.conf)) && (((uintptr_t)a->tgt_mask & 3u) == 0);
      if (f.rowbest && !fvec) {
        set_error("sinkhorn: rowbest/colbest need 16-byte aligned out / x_t / noise / conf and a 4-byte aligned tgt_mask "
                  "(only the tiled final pass tracks the bests)");
        return DRG_ERR_UNSUPPORTED;
      }
      if (fvec)
        {
          if (f.rowbest && !bests_cleared) {
            DRG_CUDA(cudaMemsetAsync(f.rowbest, 0, sizeof(unsigned long long) * (size_t)B * N, st));
            DRG_CUDA(cudaMemsetAsync(f.colbest, 0, sizeof(unsigned long long) * (size_t)B * M, st));
          }
          ProfScope prof_scope(PROF_SKH_FINAL, st);
          f.tile_rows = FT_ROWS;
          skh_final_tile_kernel<<<dim3((M + FT_THREADS * 4 - 1) / (FT_THREADS * 4), (N + f.tile_rows - 1) / f.tile_rows, B), FT_THREADS, 0, st>>>(f);
        }
      else
        {
          ProfScope prof_scope(PROF_SKH_FINAL, st);
          skh_final_kernel<false><<<dim3(gx, B), 256, 0, st>>>(f);
        }
    }
    DRG_LAUNCH_CHECK();
  }
  if (a->u)
    DRG_CUDA(cudaMemcpy2DAsync(a->u, sizeof(float) * (N + 1), w.u, sizeof(float) * pitch4(N + 1), sizeof(float) * (N + 1), B,
                               cudaMemcpyDeviceToDevice, st));
  if (a->v)
    DRG_CUDA(cudaMemcpy2DAsync(a->v, sizeof(float) * (M + 1), w.v, sizeof(float) * pitch4(M + 1), sizeof(float) * (M + 1), B,
                               cudaMemcpyDeviceToDevice, st));
  return DRG_OK;
}

static int skh_check_args(const drg_sinkhorn_args* a) {
  DRG_CHECK_ARG(a != nullptr, "args is null");
  DRG_CHECK_ARG(a->scores && a->src_mask && a->tgt_mask && a->alpha, "scores/src_mask/tgt_mask/alpha must be non-null");
  DRG_CHECK_ARG(a->B >= 1 && a->N >= 1 && a->M >= 1, "B, N, M must be >= 1");
  DRG_CHECK_ARG(a->iters >= 0, "iters must be >= 0");
  DRG_CHECK_ARG(a->out_mode >= DRG_OUT_LOG_FULL && a->out_mode <= DRG_OUT_NONE, "unknown out_mode");
  DRG_CHECK_ARG(a->out_mode == DRG_OUT_NONE || a->out != nullptr, "out is null");
  DRG_CHECK_ARG(a->out_mode != DRG_OUT_DDIM || a->x_t != nullptr, "DDIM mode needs x_t");
  return DRG_OK;
}

extern "C" int drg_sinkhorn(const drg_sinkhorn_args* a, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = skh_check_args(a);
  if (rc != DRG_OK) return rc;
  return run_sinkhorn(a, false, 1.f, workspace, workspace_bytes, stream);
}

extern "C" int drg_dual_softmax(const float* sim, const uint8_t* src_mask, const uint8_t* tgt_mask, int B, int N, int M,
                                float temperature, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  DRG_CHECK_ARG(sim && src_mask && tgt_mask && out, "sim/src_mask/tgt_mask/out must be non-null");
  DRG_CHECK_ARG(B >= 1 && N >= 1 && M >= 1, "B, N, M must be >= 1");
  DRG_CHECK_ARG(temperature > 0.f, "temperature must be > 0");
  drg_sinkhorn_args a{};
  a.scores = sim;
  a.src_mask = src_mask;
  a.tgt_mask = tgt_mask;
  a.B = B;
  a.N = N;
  a.M = M;
  a.iters = 1;
  a.apply_mask = 1;
  a.out_mode = DRG_OUT_CONF;
  a.out = out;
  return run_sinkhorn(&a, true, temperature, workspace, workspace_bytes, stream);
}

/* ---- row-sharded Sinkhorn: see include/diffreg_b200.h ---- */
static int shard_check(const drg_sinkhorn_args* a) {
  DRG_CHECK_ARG(a != nullptr, "args is null");
  DRG_CHECK_ARG(a->scores && a->src_mask && a->tgt_mask && a->alpha, "scores/src_mask/tgt_mask/alpha must be non-null");
  DRG_CHECK_ARG(a->B >= 1 && a->N >= 1 && a->M >= 1, "B, N (local rows), M must be >= 1");
  return DRG_OK;
}
extern "C" int drg_sinkhorn_shard_begin(const drg_sinkhorn_args* a, const int* global_counts, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  int rc = shard_check(a);
  if (rc != DRG_OK) return rc;
  DRG_CHECK_ARG(global_counts != nullptr, "global_counts is null");
  return run_sinkhorn(a, false, 1.f, workspace, workspace_bytes, stream, SKH_SHARD_BEGIN, global_counts, nullptr);
}
extern "C" int drg_sinkhorn_shard_local(const drg_sinkhorn_args* a, void* workspace, size_t workspace_bytes, float* partial,
                                        void* stream) {
  int rc = shard_check(a);
  if (rc != DRG_OK) return rc;
  DRG_CHECK_ARG(partial != nullptr && (((uintptr_t)partial) & 7u) == 0, "partial must be a non-null 8-byte aligned [B, M+1, 2] buffer");
  return run_sinkhorn(a, false, 1.f, workspace, workspace_bytes, stream, SKH_SHARD_LOCAL, nullptr, reinterpret_cast<float2*>(partial));
}
extern "C" int drg_sinkhorn_shard_local_exchange(const drg_sinkhorn_args* a, void* workspace, size_t workspace_bytes, void* comm,
                                                 void* stream) {
  int rc = shard_check(a);
  if (rc != DRG_OK) return rc;
  DRG_CHECK_ARG(comm != nullptr, "comm is null");
  P2PComm* c = reinterpret_cast<P2PComm*>(comm);
  const int nchunk = (a->M + 1 + 31) / 32;
  DRG_CHECK_ARG((size_t)a->B * (a->M + 1) <= c->slot_elems && (long long)a->B * nchunk <= c->nflags,
                "p2p comm too small for this shape (drg_p2p_create capacity)");
  for (int r = 0; r < c->world; ++r) DRG_CHECK_ARG(c->peer_base[r] != nullptr, "p2p comm is not connected (drg_p2p_connect)");
  const P2PView view = p2p_view(*c);
  rc = run_sinkhorn(a, false, 1.f, workspace, workspace_bytes, stream, SKH_SHARD_LOCAL_X, nullptr, nullptr, &view);
  if (rc == DRG_OK) c->epoch += 1u;
  return rc;
}
extern "C" int drg_sinkhorn_shard_iterate(const drg_sinkhorn_args* a, void* workspace, size_t workspace_bytes, void* comm, int iters,
                                          void* stream) {
  // `iters` iterations of drg_sinkhorn_shard_local_exchange enqueued back to back from C: at 8 GPUs an iteration is ~45 us of
  // GPU work, less than one Python -> ctypes round trip per launch pair costs on the host
  DRG_CHECK_ARG(iters >= 0, "iters must be >= 0");
  for (int k = 0; k < iters; ++k) {
    const int rc = drg_sinkhorn_shard_local_exchange(a, workspace, workspace_bytes, comm, stream);
    if (rc != DRG_OK) return rc;
  }
  return DRG_OK;
}
extern "C" int drg_sinkhorn_shard_update(const drg_sinkhorn_args* a, void* workspace, size_t workspace_bytes, const float* reduced,
                                         void* stream) {
  int rc = shard_check(a);
  if (rc != DRG_OK) return rc;
  DRG_CHECK_ARG(reduced != nullptr && (((uintptr_t)reduced) & 7u) == 0, "reduced must be a non-null 8-byte aligned [B, M+1, 2] buffer");
  return run_sinkhorn(a, false, 1.f, workspace, workspace_bytes, stream, SKH_SHARD_UPDATE, nullptr,
                      const_cast<float2*>(reinterpret_cast<const float2*>(reduced)));
}
extern "C" int drg_sinkhorn_shard_final(const drg_sinkhorn_args* a, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = shard_check(a);
  if (rc != DRG_OK) return rc;
  DRG_CHECK_ARG(a->out_mode >= DRG_OUT_LOG_FULL && a->out_mode <= DRG_OUT_NONE, "unknown out_mode");
  DRG_CHECK_ARG(a->out_mode == DRG_OUT_NONE || a->out != nullptr, "out is null");
  return run_sinkhorn(a, false, 1.f, workspace, workspace_bytes, stream, SKH_SHARD_FINAL, nullptr, nullptr);
}

namespace drg {
int skh_run_with_views(const drg_sinkhorn_args* a, void* workspace, size_t workspace_bytes, void* stream, SkhViews* views,
                       const SkhCollect* collect, bool* collected) {
  if (collected) *collected = false;
  int rc = skh_check_args(a);
  if (rc != DRG_OK) return rc;
  rc = run_sinkhorn(a, false, 1.f, workspace, workspace_bytes, stream, SKH_RUN_ALL, nullptr, nullptr, nullptr, collect, collected);
  if (rc != DRG_OK) return rc;
  SkhWorkspace w = carve(workspace, a->B, a->N, a->M, skh_max_g(a->B));
  views->u = w.u;
  views->v = w.v;
  views->bc = w.bc;
  views->ldu = pitch4(a->N + 1);
  views->ldv = pitch4(a->M + 1);
  return DRG_OK;
}
}  // namespace drg
