// SoftProcrustes: top-K soft correspondences -> weighted Kabsch -> gated pose -> warped points.
// sm_100a.
//
// Replaces
//   SoftProcrustesLayer.forward                    Diff-Reg-4dmatch/models/procrustes.py:48-93
//                                                  (3DMatch variant: Diff-Reg-3dmatch/models/procrustes.py:61-62)
//   SoftProcrustesLayer.batch_weighted_procrustes  Diff-Reg-4dmatch/models/procrustes.py:18-44
//   the warp  (R_forwd @ s_pcd^T + t_forwd)^T      Diff-Reg-4dmatch/models/pipeline.py:220
//
// The reference sorts all N*M confidences to use the best K = max(|src|,|tgt|)*rate of them and
// ships a 3x3 matrix to the host for a LAPACK SVD.  Here:
//   1. topk_threshold_kernel  mask counts -> K_b; 32768 hashed samples of the matrix (32 CTAs fetch, the last one
//                             carries on); the EXACT t-th largest sample (64-bit key = ordered value << 32 |
//                             ~flat index, so ties are ordered too), t ~ 2x the expected number of top-K_b
//                             entries in the sample -> lower bound L; range of the candidate histogram
//   2. topk_collect_*_kernel  ONE pass over the matrix: entries with key >= L are appended to a candidate
//                             list (CTA-local lists, one global atomic per flush) and counted in a 2048-bin
//                             histogram of their values; everything else is never touched again
//   3. procr_select_kernel    one CTA per batch element: histogram walk, short list of the crossing bin, the
//                             exact K_b-th largest key T by rank counting (general radix select for ties)
//   4. procr_moments_kernel   up to 64 CTAs per batch element: fp32 moments of the selected candidates about
//                             per-CTA centres, exact fp64 combination in the last CTA, closed-form 3x3 SVD
//                             (one-sided Jacobi, fp64), reflection fix, condition-number gate, src-point warp
//      (procr_solve_kernel: the earlier single-CTA version of 3 + 4, DRG_PROCR_SINGLE=1)
// Because L is an order statistic of the sample (not a histogram bin edge) the candidate list holds
// ~2 K_b + 16 N M / 32768 entries whatever the value distribution (flat, tied or all-zero matrices
// included).  If the sample still misleads (fewer than K_b candidates) the select kernel falls back to
// collecting the whole matrix itself: slow, but exact.
#include <stdlib.h>

#include "common.cuh"

namespace drg {

constexpr int TS_THREADS = 512;
constexpr int TS_SAMPLES = 32768;  // sample keys held in shared memory (128 KB)
constexpr int TS_FAST_TARGET = 128;  // order statistics up to this rank use the thread-maxima short cut of the bound select
constexpr int TS_FAST_LIST = 256;    // capacity of its short list (TS_THREADS threads rank it)

struct ProcrParams {
  const float* conf;             // [B,N,M], or NULL: potentials mode, conf = exp((scores - shift | mask) + u + v - norm)
  const float* scores;           // potentials mode: [B,N,M]
  const float* pu;               // [B, ldu]
  const float* pv;               // [B, ldv]
  const SkhConst* pbc;           // [B]
  const float* pshift;           // device scalar or NULL
  int ldu, ldv, apply_mask;
  const float* src_pcd;          // [B,N,3]
  const float* tgt_pcd;          // [B,M,3]
  const unsigned char* src_mask; // [B,N]
  const unsigned char* tgt_mask; // [B,M]
  int B, N, M;
  float sample_rate;
  float max_condition_num;
  int padded_lengths;            // 3DMatch variant: lengths are N, M whatever the masks say
  // workspace
  ProcrState* state;             // [B]
  unsigned int* cand_key;        // [B, N*M]
  unsigned int* cand_idx;        // [B, N*M]
  unsigned int* sample_buf;      // [B, TS_SAMPLES]
  unsigned int* sample_arrive;   // [B] arrival counters of the sampling CTAs (zero on entry)
  const uint2* cand_seg;         // [B, NUM_SMS] producer segments of the candidate list (state.seg_G > 0)
  unsigned int* pose_sync;       // [B][4] pose kernel: fallback barrier, T-ready flag, moments arrivals (zero on entry)
  long long* stamps;             // tuning only (drg_tuning_set_stamp_buffer): globaltimer stamps of batch element 0, slots 800+
  int surv_cap;                  // pose kernel: entries of the survivor list (positions of the candidates in the crossing bin or above)
  int mine_cap;                  // pose kernel: entries of its dynamic shared-memory list (>= K_max + SEL_LIST, power of two)
  unsigned int* cand_hist;       // [B, TK_BINS] histogram of the candidates' value keys (zeroed by the threshold kernel)
  double* partials;              // [B, PP_MAX_G, PM_PART] per-CTA moment partials
  // outputs
  float* R;                      // [B,3,3]
  float* t;                      // [B,3]
  float* R_forwd;                // [B,3,3]
  float* t_forwd;                // [B,3]
  double* condition;             // [B]
  unsigned char* solution_mask;  // [B]
  float* src_warped;             // [B,N,3] or NULL
  // optional: the selected correspondences, K_max slots per batch element (weight 0 beyond K_b)
  int K_max;
  float* sel_w;                  // [B,K_max] or NULL
  int* sel_src;                  // [B,K_max] or NULL
  int* sel_tgt;                  // [B,K_max] or NULL
};

// number of non-zero bytes among m[tid], m[tid + nthreads], ... (16 bytes per load when aligned; bools are 0 / 1)
__device__ __forceinline__ unsigned int count_bytes16(const unsigned char* __restrict__ m, int n, int tid, int nthreads) {
  unsigned int c = 0;
  if ((((uintptr_t)m) & 15u) == 0) {
    const int n16 = n >> 4;
    const uint4* m4 = reinterpret_cast<const uint4*>(m);
    for (int i = tid; i < n16; i += nthreads) {
      const uint4 q = m4[i];
      c += __popc(q.x & 0x01010101u) + __popc(q.y & 0x01010101u) + __popc(q.z & 0x01010101u) + __popc(q.w & 0x01010101u);
    }
    for (int i = (n16 << 4) + tid; i < n; i += nthreads) c += m[i] ? 1u : 0u;
  } else {
    for (int i = tid; i < n; i += nthreads) c += m[i] ? 1u : 0u;
  }
  return c;
}

// the confidence at flat position `pos` of batch element b: stored, or recomputed from the Sinkhorn potentials with
// the arithmetic of skh_final_tile_kernel (so that the matrix never has to be materialised for the pose step)
__device__ __forceinline__ float conf_at(const ProcrParams& p, int b, size_t pos) {
  const size_t total = (size_t)p.N * p.M;
  if (p.conf) return p.conf[(size_t)b * total + pos];
  const int i = (int)(pos / (size_t)p.M), j = (int)(pos - (size_t)i * p.M);
  const bool ok = !p.apply_mask || (p.src_mask[(size_t)b * p.N + i] && p.tgt_mask[(size_t)b * p.M + j]);
  const float shift = p.pshift ? *p.pshift : 0.f;
  const float zz = ok ? (p.scores[(size_t)b * total + pos] - shift) : -INFINITY;
  const float la = ((zz + p.pu[(size_t)b * p.ldu + i]) + p.pv[(size_t)b * p.ldv + j]) - p.pbc[b].norm;
  return ex2(la * LOG2E);
}

// ---- 1. K_b and the sample-based lower bound ---------------------------------------------------
__global__ void __launch_bounds__(TS_THREADS) topk_threshold_kernel(const ProcrParams p) {
  extern __shared__ __align__(16) unsigned int sample_key[];  // [TS_SAMPLES]
  __shared__ SelectScratch sc;
  __shared__ unsigned long long cnt_s;
  const int b = blockIdx.y, tid = threadIdx.x;
  // ---- sample the matrix: every CTA of the batch element fetches its share of the hashed positions (scattered
  //      4-byte reads: spread over TS_CTAS CTAs so that they are all in flight at once) into a global buffer; the
  //      last CTA to arrive pulls the whole sample into shared memory and carries on alone.
  const size_t total = (size_t)p.N * p.M;
  const bool all = total <= (size_t)TS_SAMPLES;
  const unsigned int n_s = all ? (unsigned int)total : (unsigned int)TS_SAMPLES;
  const unsigned int salt = 0x9e3779b9u * (unsigned int)(b + 1);
  auto sample_pos = [&](unsigned int q) -> unsigned int {
    return all ? q : (unsigned int)(((unsigned long long)hash_u32(q + salt) * (unsigned long long)total) >> 32);
  };
  unsigned int* sbuf = p.sample_buf + (size_t)b * TS_SAMPLES;
  {
    const unsigned int per_cta = (n_s + gridDim.x - 1) / gridDim.x;
    const unsigned int q_lo = blockIdx.x * per_cta, q_hi = min(n_s, q_lo + per_cta);
    for (unsigned int q0 = q_lo + tid; q0 < q_hi; q0 += TS_THREADS * 4) {
      float val[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned int q = q0 + k * TS_THREADS;
        val[k] = (q < q_hi) ? conf_at(p, b, sample_pos(q)) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned int q = q0 + k * TS_THREADS;
        if (q < q_hi) sbuf[q] = float_to_ordered(val[k]);
      }
    }
  }
  __shared__ unsigned int ticket_s;
  __threadfence();
  __syncthreads();
  if (tid == 0) ticket_s = atomicAdd(&p.sample_arrive[b], 1u);
  __syncthreads();
  if (ticket_s != gridDim.x - 1) return;   // not the last CTA of this batch element
  if (tid == 0) p.sample_arrive[b] = 0u;   // self-reset for the next call
  __threadfence();
  for (int q = tid; q < TK_BINS; q += TS_THREADS) p.cand_hist[(size_t)b * TK_BINS + q] = 0u;  // filled by the collect kernel
  __shared__ unsigned int smax_s;
  if (tid == 0) smax_s = 0u;
  // the staging loop also tracks this thread's largest sample value (fast path of the select below)
  // (16-byte loads, eight in flight per thread: the 128 KB come in two round trips instead of eight)
  unsigned int tmax = 0u;
  {
    const unsigned int n4 = n_s >> 2;
    const uint4* s4 = reinterpret_cast<const uint4*>(sbuf);  // sbuf is 256-byte aligned (workspace carve)
    uint4* d4 = reinterpret_cast<uint4*>(sample_key);
    for (unsigned int q0 = 0; q0 < n4; q0 += TS_THREADS * 8) {
      uint4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const unsigned int q = q0 + k * TS_THREADS + tid;
        v[k] = q < n4 ? __ldcg(s4 + q) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const unsigned int q = q0 + k * TS_THREADS + tid;
        if (q < n4) {
          d4[q] = v[k];
          tmax = max(max(tmax, max(v[k].x, v[k].y)), max(v[k].z, v[k].w));
        }
      }
    }
    for (unsigned int q = (n4 << 2) + tid; q < n_s; q += TS_THREADS) {
      const unsigned int v = __ldcg(sbuf + q);
      sample_key[q] = v;
      tmax = max(tmax, v);
    }
  }
  __syncthreads();
  {
    const unsigned int wm = __reduce_max_sync(0xffffffffu, tmax);
    if ((tid & 31) == 0) atomicMax(&smax_s, wm);
  }
  // ---- mask counts of every batch element (K is a mean over the batch)      procrustes.py:61-65
  float cap_sum = 0.f;
  int my_cap = 0;
  for (int bb = 0; bb < p.B; ++bb) {
    int ns = p.N, nt = p.M;
    if (!p.padded_lengths) {
      if (tid == 0) cnt_s = 0ull;
      __syncthreads();
      unsigned long long c = ((unsigned long long)count_bytes16(p.src_mask + (size_t)bb * p.N, p.N, tid, TS_THREADS) << 32) |
                             (unsigned long long)count_bytes16(p.tgt_mask + (size_t)bb * p.M, p.M, tid, TS_THREADS);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if ((tid & 31) == 0) atomicAdd(&cnt_s, c);
      __syncthreads();
      ns = (int)(cnt_s >> 32);
      nt = (int)(cnt_s & 0xFFFFFFFFull);
      __syncthreads();
    }
    // (max(len) * sample_rate).int()   procrustes.py:63-64 (fp32 product, truncation)
    const int cap = (int)((float)max(ns, nt) * p.sample_rate);
    cap_sum += (float)cap;
    if (bb == b) my_cap = cap;
  }
  // sample_n_points = entry_max.float().mean().int()   procrustes.py:65
  const int K = (int)(cap_sum / (float)p.B);
  const int Kb = min(min(K, my_cap), p.K_max);
  // ---- the t-th largest sample key is the lower bound
  unsigned long long lower = 0ull;
  if (Kb > 0) {
    long long target;
    if (all) {
      target = Kb;  // the sample IS the matrix: the bound is the exact K_b-th largest
    } else {
      // twice the expected number of top-K_b entries inside the sample, plus slack for small counts
      const double expect = (double)Kb * ((double)n_s / (double)total);
      target = (long long)(2.0 * expect + 16.0);
    }
    if (target < (long long)n_s) {
      // Fast path (the normal case: the 32nd largest of 32768 samples).  A radix select over the whole sample spends
      // its levels on ~32 k shared-memory atomics into the few bins the exponents of the confidences occupy.  Instead:
      // the target-th largest of the 512 per-thread maxima is a value b0 that at least `target` samples reach, and
      // hardly more than `target` do (the top samples are spread over the threads at random); those few go to a short
      // list whose target-th largest 64-bit key, found by rank counting, is exactly the key the full select returns.
      bool found = false;
      if (target <= (long long)TS_FAST_TARGET) {
        __shared__ unsigned int tmax_s[TS_THREADS];
        __shared__ unsigned long long list_s[TS_FAST_LIST];
        __shared__ unsigned int list_n, b0_s;
        __shared__ unsigned long long lower_s;
        tmax_s[tid] = tmax;
        if (tid == 0) {
          list_n = 0u;
          b0_s = 0xFFFFFFFFu;
          lower_s = 0ull;
        }
        __syncthreads();
        const unsigned long long tb = block_select_kth<TS_THREADS>([&](size_t e) { return make_key64(tmax_s[e], (unsigned int)e); }, (size_t)TS_THREADS, (int)target, sc.hist, sc.ctl);
        // the select may stop at a bucket edge: the exact bound is the smallest selected maximum
        unsigned int mine = (make_key64(tmax, (unsigned int)tid) >= tb) ? tmax : 0xFFFFFFFFu;
        mine = __reduce_min_sync(0xffffffffu, mine);
        if ((tid & 31) == 0) atomicMin(&b0_s, mine);
        __syncthreads();
        const unsigned int b0 = b0_s;
        for (unsigned int q = tid; q < n_s; q += TS_THREADS) {
          const unsigned int v = sample_key[q];
          if (v >= b0) {
            const unsigned int pos = atomicAdd(&list_n, 1u);
            if (pos < (unsigned int)TS_FAST_LIST) list_s[pos] = make_key64(v, sample_pos(q));
          }
        }
        __syncthreads();
        const unsigned int L = list_n;
        if (L <= (unsigned int)TS_FAST_LIST && (long long)L >= target) {
          if ((unsigned int)tid < L) {
            const unsigned long long my = list_s[tid];
            unsigned int rank = 0u;
            for (unsigned int q = 0; q < L; ++q) {
              const unsigned long long o = list_s[q];
              rank += (o > my || (o == my && q < (unsigned int)tid)) ? 1u : 0u;  // duplicates (hash collisions) ranked by position
            }
            if (rank == (unsigned int)(target - 1)) lower_s = my;
          }
          __syncthreads();
          lower = lower_s;
          found = true;
        }
      }
      if (!found) {
        lower = block_select_kth<TS_THREADS>([&](size_t e) { return make_key64(sample_key[e], sample_pos((unsigned int)e)); }, (size_t)n_s, (int)target, sc.hist, sc.ctl);
      }
    }
  }
  if (tid == 0) {
    ProcrState s;
    s.Kb = Kb;
    s.n_cand = 0u;
    s.lower_key = lower;
    s.T = 0ull;
    // histogram range: from the bound to twice the distance of the largest sample (larger keys share the top bin)
    const unsigned int kmin = (unsigned int)(lower >> 32);
    const unsigned int smax = smax_s;
    unsigned int range = 0xFFFFFFFFu - kmin;
    if (lower != 0ull && smax > kmin) {
      const unsigned long long r2 = 2ull * (unsigned long long)(smax - kmin) + 1ull;
      if (r2 < (unsigned long long)range) range = (unsigned int)r2;
    }
    s.hist_kmin = kmin;
    s.hist_sh = range ? __clz((int)range) : 32;
    s.sel_count = 0u;
    s.pad_ = 0u;
    s.seg_G = 0;          // the stand-alone collect kernels stride over the rows: no per-producer segments
    s.seg_broken = 0u;
    p.state[b] = s;
  }
}

// ---- 4. collect candidates ------------------------------------------------------------------
__device__ __forceinline__ void append_candidates(const ProcrParams& p, int b, size_t total, const bool (&take)[4],
                                                  const unsigned int (&key)[4], unsigned int flat0) {
  const int lane = threadIdx.x & 31;
  int mine = (int)take[0] + (int)take[1] + (int)take[2] + (int)take[3];
  const unsigned int any = __ballot_sync(0xffffffffu, mine > 0);
  if (any == 0u) return;
  // warp-aggregated append: exclusive prefix of `mine` over the lanes
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int tmp = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += tmp;
  }
  const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned int base = 0;
  if (lane == 0) base = atomicAdd(&p.state[b].n_cand, (unsigned int)warp_total);
  base = __shfl_sync(0xffffffffu, base, 0);
  unsigned int pos = base + (unsigned int)(incl - mine);
  const unsigned int hk = p.state[b].hist_kmin;
  const int hs = p.state[b].hist_sh;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (take[e]) {
      p.cand_key[(size_t)b * total + pos] = key[e];
      p.cand_idx[(size_t)b * total + pos] = flat0 + e;
      atomicAdd(&p.cand_hist[(size_t)b * TK_BINS + cand_bin(key[e], hk, hs)], 1u);
      ++pos;
    }
  }
}

// CTA-local variant: candidates go to a shared-memory list first; the CTA reserves its slice of the global list with
// ONE atomic per flush (the single global counter was the bottleneck: ~19 k dependent same-address atomics).
constexpr int COLLECT_CAP = 5120;  // shared-memory candidate list of a CTA; flushed when a worst-case chunk (4096) might not fit
__device__ __forceinline__ void append_candidates_smem(unsigned int* s_key, unsigned int* s_idx, unsigned int* s_n,
                                                       const bool (&take)[4], const unsigned int (&key)[4], unsigned int flat0) {
  const int lane = threadIdx.x & 31;
  int mine = (int)take[0] + (int)take[1] + (int)take[2] + (int)take[3];
  const unsigned int any = __ballot_sync(0xffffffffu, mine > 0);
  if (any == 0u) return;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int tmp = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += tmp;
  }
  const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned int base = 0;
  if (lane == 0) base = atomicAdd(s_n, (unsigned int)warp_total);
  base = __shfl_sync(0xffffffffu, base, 0);
  unsigned int pos = base + (unsigned int)(incl - mine);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (take[e]) {
      s_key[pos] = key[e];
      s_idx[pos] = flat0 + e;
      ++pos;
    }
  }
}

__global__ void __launch_bounds__(256) topk_collect_kernel(const ProcrParams p) {
  const int b = blockIdx.y;
  const size_t total = (size_t)p.N * p.M;
  const float* x = (p.conf ? p.conf : p.scores) + (size_t)b * total;
  const unsigned long long lower = p.state[b].lower_key;
  // vector path: quads never straddle a row in potentials mode (M % 4 == 0)
  const bool vec = ((total & 3) == 0) && ((((uintptr_t)x) & 15u) == 0) && (p.conf || (p.M & 3) == 0);
  const size_t n4 = vec ? (total >> 2) : ((total + 3) >> 2);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float shift = (!p.conf && p.pshift) ? *p.pshift : 0.f;
  const float norm = p.conf ? 0.f : p.pbc[b].norm;
  // all lanes of a warp iterate the same number of times (warp-collective append)
  const size_t iters = (n4 + stride - 1) / stride;
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t it = 0; it < iters; ++it, q += stride) {
    bool take[4] = {false, false, false, false};
    unsigned int key[4] = {0u, 0u, 0u, 0u};
    if (q < n4) {
      float v[4];
      if (vec) {
        const float4 t = *reinterpret_cast<const float4*>(x + q * 4);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        if (!p.conf) {
          const int i = (int)((q * 4) / (size_t)p.M), j = (int)(q * 4 - (size_t)i * p.M);
          const float ui = p.pu[(size_t)b * p.ldu + i];
          const float4 vj = *reinterpret_cast<const float4*>(p.pv + (size_t)b * p.ldv + j);
          bool ok[4] = {true, true, true, true};
          if (p.apply_mask) {
            const bool row_ok = p.src_mask[(size_t)b * p.N + i];
            const uchar4 tm = *reinterpret_cast<const uchar4*>(p.tgt_mask + (size_t)b * p.M + j);
            ok[0] = row_ok && tm.x; ok[1] = row_ok && tm.y; ok[2] = row_ok && tm.z; ok[3] = row_ok && tm.w;
          }
          const float vv[4] = {vj.x, vj.y, vj.z, vj.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float zz = ok[e] ? (v[e] - shift) : -INFINITY;
            v[e] = ex2((((zz + ui) + vv[e]) - norm) * LOG2E);
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (q * 4 + e < total) ? conf_at(p, b, q * 4 + e) : -INFINITY;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        key[e] = float_to_ordered(v[e]);
        take[e] = (q * 4 + e < total) && make_key64(key[e], (unsigned int)(q * 4 + e)) >= lower;
      }
    }
    append_candidates(p, b, total, take, key, (unsigned int)(q * 4));
  }
}

// potentials mode, aligned rows (M % 4 == 0): CTAs stride over rows, threads over column quads -- no index division,
// u_i once per row, v and the target mask as 16-byte loads
__global__ void __launch_bounds__(256) topk_collect_rows_kernel(const ProcrParams p) {
  const int b = blockIdx.y;
  const int N = p.N, M = p.M;
  const size_t total = (size_t)N * M;
  const float* x = p.scores + (size_t)b * total;
  const unsigned long long lower = p.state[b].lower_key;
  const unsigned int hist_kmin = p.state[b].hist_kmin;
  const int hist_sh = p.state[b].hist_sh;
  const float shift = p.pshift ? *p.pshift : 0.f;
  const SkhConst bc = p.pbc[b];
  const float norm = bc.norm;
  const bool masked = p.apply_mask && bc.pad != 1.f;  // pad == 1: the Sinkhorn saw no padded row / column
  const float* u_b = p.pu + (size_t)b * p.ldu;
  const float* v_b = p.pv + (size_t)b * p.ldv;
  // Pre-filter in the log2 domain: conf = 2^(la2) >= L  <=>  la2 >= log2 L up to rounding (~1e-5 here, the pre-filter
  // folds the constants differently from the exact expression); the 1e-3 margin only lets a few extra quads through to
  // the exact 64-bit key comparison.  Almost every quad stops after 1 FADD + 1 FFMA + 1 compare per element and one
  // warp vote -- the pass was issue-bound, not bandwidth-bound, with the exponential and the key compare on every element.
  const float lower_val = ordered_to_float((unsigned int)(lower >> 32));
  const float thr2 = (lower_val > 0.f) ? (log2f(lower_val) - 1.0e-3f) : -INFINITY;
  const int ncol_iters = (M + 1023) / 1024;  // all lanes iterate alike (warp-collective vote / append)
  constexpr int CU = 4;                      // column chunks loaded together: 4 x 16 bytes in flight per thread
  __shared__ unsigned int s_key[COLLECT_CAP], s_idx[COLLECT_CAP];
  __shared__ unsigned int s_n, s_base;
  if (threadIdx.x == 0) s_n = 0u;
  __syncthreads();
  float cterm[CU][4];  // (v_j - norm - shift) * log2e of this thread's columns; -inf: masked or out of range
  auto load_cterm = [&](int k0) {
#pragma unroll
    for (int q = 0; q < CU; ++q) {
      const int j = 4 * (int)threadIdx.x + 1024 * (k0 + q);
      float4 vj = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      bool okc[4] = {true, true, true, true};
      if (k0 + q < ncol_iters && j < M) {
        vj = *reinterpret_cast<const float4*>(v_b + j);
        if (masked) {
          const uchar4 tm = *reinterpret_cast<const uchar4*>(p.tgt_mask + (size_t)b * M + j);
          okc[0] = tm.x; okc[1] = tm.y; okc[2] = tm.z; okc[3] = tm.w;
        }
      }
      const float vv[4] = {vj.x, vj.y, vj.z, vj.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) cterm[q][e] = okc[e] ? ((vv[e] - norm) - shift) * LOG2E : -INFINITY;
    }
  };
  if (ncol_iters <= CU) load_cterm(0);
  auto flush = [&]() {  // CTA-wide: reserve a slice of the global list with one atomic, copy, reset
    __syncthreads();
    const unsigned int cnt = s_n;
    if (cnt) {
      if (threadIdx.x == 0) s_base = atomicAdd(&p.state[b].n_cand, cnt);
      __syncthreads();
      const unsigned int base = s_base;
      for (unsigned int e = threadIdx.x; e < cnt; e += blockDim.x) {
        const unsigned int k32 = s_key[e];
        p.cand_key[(size_t)b * total + base + e] = k32;
        p.cand_idx[(size_t)b * total + base + e] = s_idx[e];
        atomicAdd(&p.cand_hist[(size_t)b * TK_BINS + cand_bin(k32, hist_kmin, hist_sh)], 1u);  // level 1 of the select, for free
      }
      __syncthreads();
      if (threadIdx.x == 0) s_n = 0u;
      __syncthreads();
    }
  };
  auto load_row = [&](float4(&dst)[CU], int i, int k0) {
#pragma unroll
    for (int q = 0; q < CU; ++q) {
      const int j = 4 * (int)threadIdx.x + 1024 * (k0 + q);
      dst[q] = (i < N && k0 + q < ncol_iters && j < M) ? __ldcs(reinterpret_cast<const float4*>(x + (size_t)i * M + j))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // the next chunk (normally: the next row) is requested before this one is examined
  float4 zn[CU];
  load_row(zn, blockIdx.x, 0);
  for (int i = blockIdx.x; i < N; i += gridDim.x) {
    const float ui = u_b[i];
    const bool row_ok = !masked || p.src_mask[(size_t)b * N + i];
    const float ui2 = row_ok ? ui * LOG2E : -INFINITY;
    for (int k0 = 0; k0 < ncol_iters; k0 += CU) {
      if (ncol_iters > CU) load_cterm(k0);
      float4 zq[CU];
#pragma unroll
      for (int q = 0; q < CU; ++q) zq[q] = zn[q];
      if (k0 + CU < ncol_iters) load_row(zn, i, k0 + CU);
      else load_row(zn, i + (int)gridDim.x, 0);
      // room for a worst-case chunk (every element a candidate)?  Block-uniform: s_n is read between barriers.
      __syncthreads();
      if (s_n > (unsigned int)(COLLECT_CAP - CU * 1024)) flush();
#pragma unroll
      for (int q = 0; q < CU; ++q) {
        if (k0 + q >= ncol_iters) break;  // uniform
        const float zz[4] = {zq[q].x, zq[q].y, zq[q].z, zq[q].w};
        bool maybe = false;
#pragma unroll
        for (int e = 0; e < 4; ++e) maybe = maybe || !(fmaf(zz[e], LOG2E, ui2 + cterm[q][e]) < thr2);  // NaN passes too
        if (!__any_sync(0xffffffffu, maybe)) continue;
        // exact path (rare): the expression and association of skh_final_tile_kernel, ordered key, 64-bit compare
        const int j = 4 * (int)threadIdx.x + 1024 * (k0 + q);
        bool take[4] = {false, false, false, false};
        unsigned int key[4] = {0u, 0u, 0u, 0u};
        const unsigned int flat0 = (unsigned int)((size_t)i * M + j);
        if (maybe && j < M) {
          const float4 vj = *reinterpret_cast<const float4*>(v_b + j);
          bool ok[4] = {row_ok, row_ok, row_ok, row_ok};
          if (masked) {
            const uchar4 tm = *reinterpret_cast<const uchar4*>(p.tgt_mask + (size_t)b * M + j);
            ok[0] = row_ok && tm.x; ok[1] = row_ok && tm.y; ok[2] = row_ok && tm.z; ok[3] = row_ok && tm.w;
          }
          const float vv[4] = {vj.x, vj.y, vj.z, vj.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float zs = ok[e] ? (zz[e] - shift) : -INFINITY;
            key[e] = float_to_ordered(ex2((((zs + ui) + vv[e]) - norm) * LOG2E));
            take[e] = make_key64(key[e], flat0 + e) >= lower;
          }
        }
        append_candidates_smem(s_key, s_idx, &s_n, take, key, flat0);
      }
    }
  }
  flush();
}

// ---- 5. select + Kabsch + warp ----------------------------------------------------------------
// Block-wide sums of K fp32 per-thread partials: fp32 warp shuffles (pairwise), then the per-warp partials are added
// in fp64 in a fixed order (bitwise reproducible) and everybody reads the totals.  fp64 vector math is slow on this
// part, so it is kept to this last step and to the 3x3 solve.
struct MomentScratch {
  float part[32][16];
  double total[16];
};
template <int K>
__device__ __forceinline__ void block_sum_f32(const float (&v)[K], double (&out)[K], MomentScratch& ms) {
  static_assert(K <= 16, "MomentScratch holds 16 values");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  __syncthreads();  // scratch reuse
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) ms.part[warp][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += (double)ms.part[w][threadIdx.x];
    ms.total[threadIdx.x] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = ms.total[k];
}

// One-sided Jacobi SVD of a 3x3 matrix (fp64): A = U diag(s) V^T, singular values descending.
__device__ void svd3x3(const double A[3][3], double U[3][3], double s[3], double V[3][3]) {
  double W[3][3];  // working copy, columns get orthogonalised: W = A V
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      W[i][j] = A[i][j];
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  // fp64 divisions and square roots are ~100-cycle software sequences and this runs on ONE thread: a rotation uses one
  // sqrt, one division and one rsqrt (t = 2g sign(d) / (|d| + sqrt(d^2 + 4 g^2)) with d = beta - alpha is the same
  // tangent as sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = d / 2g), and convergence is tested on squares.
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
    for (int pcol = 0; pcol < 2; ++pcol) {
      for (int qcol = pcol + 1; qcol < 3; ++qcol) {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int i = 0; i < 3; ++i) {
          alpha += W[i][pcol] * W[i][pcol];
          beta += W[i][qcol] * W[i][qcol];
          gamma += W[i][pcol] * W[i][qcol];
        }
        if (gamma == 0.0) continue;
        if (gamma * gamma > 1e-30 * fmax(alpha * beta, 1e-300)) rotated = true;  // |gamma| / sqrt(alpha beta) > 1e-15
        const double d = beta - alpha;
        const double tt = (d >= 0.0 ? 2.0 : -2.0) * gamma / (fabs(d) + sqrt(d * d + 4.0 * gamma * gamma));
        const double c = rsqrt(1.0 + tt * tt), sn = c * tt;
        for (int i = 0; i < 3; ++i) {
          const double wp = W[i][pcol], wq = W[i][qcol];
          W[i][pcol] = c * wp - sn * wq;
          W[i][qcol] = sn * wp + c * wq;
          const double vp = V[i][pcol], vq = V[i][qcol];
          V[i][pcol] = c * vp - sn * vq;
          V[i][qcol] = sn * vp + c * vq;
        }
      }
    }
    if (!rotated) break;
  }
  for (int j = 0; j < 3; ++j) s[j] = sqrt(W[0][j] * W[0][j] + W[1][j] * W[1][j] + W[2][j] * W[2][j]);
  // sort descending (columns of W and V move together)
  for (int a = 0; a < 2; ++a)
    for (int c2 = a + 1; c2 < 3; ++c2)
      if (s[c2] > s[a]) {
        double tmp = s[a]; s[a] = s[c2]; s[c2] = tmp;
        for (int i = 0; i < 3; ++i) {
          tmp = W[i][a]; W[i][a] = W[i][c2]; W[i][c2] = tmp;
          tmp = V[i][a]; V[i][a] = V[i][c2]; V[i][c2] = tmp;
        }
      }
  // U columns; rank-deficient columns are completed to an orthonormal basis
  const double tiny = 1e-300;
  for (int j = 0; j < 3; ++j) {
    const double inv = s[j] > tiny ? 1.0 / s[j] : 0.0;
    for (int i = 0; i < 3; ++i) U[i][j] = W[i][j] * inv;
  }
  if (!(s[0] > tiny)) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) U[i][j] = (i == j) ? 1.0 : 0.0;
  } else {
    if (!(s[1] > tiny)) {
      // any unit vector orthogonal to U[:,0]
      int m = 0;
      if (fabs(U[1][0]) < fabs(U[m][0])) m = 1;
      if (fabs(U[2][0]) < fabs(U[m][0])) m = 2;
      double e[3] = {0, 0, 0};
      e[m] = 1.0;
      const double d = U[m][0];
      double n = 0.0;
      for (int i = 0; i < 3; ++i) {
        U[i][1] = e[i] - d * U[i][0];
        n += U[i][1] * U[i][1];
      }
      n = 1.0 / sqrt(n);
      for (int i = 0; i < 3; ++i) U[i][1] *= n;
    }
    if (!(s[2] > tiny)) {
      U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
      U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
      U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
  }
}

__device__ __forceinline__ double det3(const double A[3][3]) {
  return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
         A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
}

// From the centred weighted covariance S = sum w_norm (y - my)(x - mx)^T and the (shrunk) weighted means to
// (R, t, condition); one thread.                                                 procrustes.py:35-43
// (not inlined: the pose kernel calls it twice -- once to warm the caches -- and both calls must run the SAME code)
__device__ __noinline__ void kabsch_solve(const double S[3][3], const double mx[3], const double my[3], float R_out[9], float t_out[3],
                             double* cond_out) {
  double U[3][3], sv[3], V[3][3];
  svd3x3(S, U, sv, V);
  *cond_out = sv[0] / sv[2];  // D.max / D.min   (inf or nan for rank-deficient input, as in the reference)
  const double dd = det3(U) * det3(V);
  double R[3][3];
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) R[a][c] = U[a][0] * V[c][0] + U[a][1] * V[c][1] + dd * U[a][2] * V[c][2];
  float Rf[3][3];
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) {
      Rf[a][c] = (float)R[a][c];
      R_out[a * 3 + c] = Rf[a][c];
    }
  // t = mean_Y - R mean_X  in fp32  (procrustes.py:43)
  for (int a = 0; a < 3; ++a) {
    const float mxf[3] = {(float)mx[0], (float)mx[1], (float)mx[2]};
    float acc = 0.f;
    for (int c = 0; c < 3; ++c) acc += Rf[a][c] * mxf[c];
    t_out[a] = (float)my[a] - acc;
  }
}

__device__ void finish_pose(const ProcrParams& p, int b, const float R[9], const float t[3], double cond) {
  const bool ok = cond < (double)p.max_condition_num;  // false for nan        (procrustes.py:87)
  for (int k = 0; k < 9; ++k) {
    p.R[b * 9 + k] = R[k];
    p.R_forwd[b * 9 + k] = ok ? R[k] : ((k % 4 == 0) ? 1.f : 0.f);
  }
  for (int k = 0; k < 3; ++k) {
    p.t[b * 3 + k] = t[k];
    p.t_forwd[b * 3 + k] = ok ? t[k] : 0.f;
  }
  p.condition[b] = cond;
  p.solution_mask[b] = ok ? 1 : 0;
}

// ---- 3. select + moments + Kabsch + warp: ONE cooperative kernel ---------------------------------------------------
// G CTAs per batch element.  Every CTA walks the candidate histogram and scans the candidate keys itself (the list is
// ~16-25 k keys, L2-resident), so every CTA knows the exact K_b-th largest key T without a kernel boundary or a broadcast:
//   * the candidates of the histogram bin where the count crosses K_b (10-70 of them) go to a short list that is ranked
//     by counting -- 64-bit keys (value, ~index) are distinct, so T is exact and identical on every CTA;
//   * heavily tied values (crossing bin > SEL_LIST entries) take the general six-level radix select on CTA 0, the other
//     CTAs wait for its T (release / acquire flag; the launch is cooperative, so they are co-resident);
//   * fewer candidates than K_b (the bound misled): all CTAs rebuild the candidate list from the whole matrix, one
//     grid barrier, then the general select -- slow, exact, practically never taken.
// In the same scan a CTA keeps the selected candidates of ITS rows (source index in [g N / G, (g + 1) N / G)), sorts
// them by flat index (bitonic, shared memory) and reduces them in that order: thread t takes sorted entries t, t + 256, ...
// and the partial sums meet in fixed shuffle / warp order, so the fp32 moments -- and the pose -- are bit-reproducible
// from run to run although the candidate list is appended with atomics.  As in the reference (procrustes.py:30-35) the
// points are centred before they are multiplied: each CTA centres on its own weighted mean, the last CTA to arrive
// combines the partials exactly in fp64 (fixed order over the CTAs), solves the 3x3 problem and warps the source points.
constexpr int PP_THREADS = 512;    // (one CTA per SM: the more warps, the better the L2 latency of the list scan is hidden)
constexpr int PP_MAX_G = 64;       // CTAs per batch element (at most)
constexpr int PM_SUMS = 17;        // W, W_abs, sum w x [3], sum w y [3], sum w y x^T [9]
constexpr int PM_PART = 24;        // doubles per CTA partial (PM_SUMS, padded)
constexpr int PP_RANK_SORT = 1024; // up to this many entries a CTA's list is sorted by rank counting, beyond by a bitonic sort
constexpr int SEL_LIST = 256;      // short list of the crossing histogram bin
constexpr unsigned long long PP_SENTINEL = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define WSYNC() asm volatile("bar.sync 1, %0;" ::"n"(PP_THREADS) : "memory")
struct SyncWorkers {
  __device__ __forceinline__ void operator()() const { WSYNC(); }
};
#define PSTAMPF(k) do { if (p.stamps && b == 0 && lane == 0) { p.stamps[(k)] = global_ns(); p.stamps[(k) + 100] = clock64(); } } while (0)
#define PSTAMP0(k) do { if (p.stamps && b == 0 && g == 0 && tid == 0) p.stamps[(k)] = global_ns(); } while (0)
#define PSTAMPL(k) do { if (p.stamps && b == 0 && tid == 0) { p.stamps[(k)] = global_ns(); p.stamps[(k) + 100] = clock64(); } } while (0)
__global__ void __launch_bounds__(PP_THREADS + 32) procr_pose_kernel(const ProcrParams p) {
  extern __shared__ __align__(16) unsigned long long mine_s[];  // [mine_cap] (flat index << 32 | value key) of this CTA's rows
  __shared__ SelectScratch sc;
  __shared__ unsigned long long list_s[SEL_LIST];
  __shared__ unsigned int list_n, mine_n, surv_n;
  __shared__ int bin_s;
  __shared__ unsigned int cum_s, hsel_s;
  __shared__ unsigned long long T_s;
  __shared__ float pose_s[12];
  __shared__ double wsum_s[PP_THREADS / 32][PM_SUMS + 1];
  const int b = blockIdx.y, g = blockIdx.x, G = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, M = p.M;
  const size_t total = (size_t)N * M;
  unsigned int* ckey = p.cand_key + (size_t)b * total;
  unsigned int* cidx = p.cand_idx + (size_t)b * total;
  unsigned int* sync = p.pose_sync + 4 * b;  // [0] fallback barrier arrivals, [1] T-ready flag, [2] moments arrivals (zero on entry)
  PSTAMP0(800);
  // ---- the finisher: warp PP_THREADS / 32 of CTA 0 (it exits at once in the other CTAs; the 512 working threads of every
  //      CTA synchronise among themselves on named barrier 1).  The 3x3 solve runs on ONE thread through fp64 division /
  //      square-root subroutines; inside a sampler step that streams hundreds of MB through L2 that code is cold in every
  //      cache and each first touch of a line is a DRAM round trip on the critical path (measured: 9.8 us cold, 5.0 us
  //      warm).  So this warp first runs the solve on a harmless matrix -- while the working warps scan the candidates --
  //      then waits for the G partial sums, combines them in fixed order, solves and publishes the pose.
  if (tid >= PP_THREADS) {
    if (g != 0) return;
    const int Kb = __ldcg(&p.state[b].Kb);
    const double* pg = p.partials + (size_t)b * PP_MAX_G * PM_PART;
    // Two trips through the SAME code: trip 0 is the warm-up (it combines whatever the partial records hold and solves a
    // harmless matrix, publishing nothing) while the working warps still scan the candidates; trip 1 waits for the G
    // arrivals and is the real thing, now with every instruction line in this SM's caches.
    for (int trip = 0; trip < 2; ++trip) {
      if (trip == 1) {
        if (lane == 0) {
          while (ld_acquire_gpu_u32(sync + 2) < (unsigned int)G) {
          }
        }
        __syncwarp();
        PSTAMPF(807);
      }
      // lane l adds the CTAs l, l + 32 (ascending), a fixed shuffle tree adds the lanes: bit-reproducible
      double tot[PM_SUMS];
#pragma unroll
      for (int k = 0; k < PM_SUMS; ++k) tot[k] = 0.0;
      for (int gg = lane; gg < G; gg += 32) {
#pragma unroll
        for (int k = 0; k < PM_SUMS; ++k) tot[k] += __ldcg(pg + (size_t)gg * PM_PART + k);
      }
#pragma unroll
      for (int k = 0; k < PM_SUMS; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot[k] += __shfl_xor_sync(0xffffffffu, tot[k], o);
      }
      if (lane == 0) {
        if (trip == 0) {   // harmless, well-conditioned stand-in sums (the records may hold anything)
#pragma unroll
          for (int k = 0; k < PM_SUMS; ++k) tot[k] = (k == 0 || k == 1) ? 1.0 : ((k == 8 || k == 12 || k == 16) ? 0.7 + 0.1 * k : 0.01 * k);
        }
        // w_norm = w / (sum|w| + eps): the normalised weights sum to slightly less than one  (procrustes.py:29-30);
        // S = sum w_norm (y - my)(x - mx)^T = inv * [sum w y x^T - (sum w y) mx^T - my (sum w x)^T + W my mx^T]
        const double inv = 1.0 / (tot[1] + 1e-4);
        const float invf = (float)inv;
        double mx[3], my[3], S[3][3];
        for (int a = 0; a < 3; ++a) {
          mx[a] = (Kb > 0) ? (double)(float)(tot[2 + a] * inv) : 0.0;  // the means are fp32 values, as in the reference
          my[a] = (Kb > 0) ? (double)(float)(tot[5 + a] * inv) : 0.0;
        }
        for (int a = 0; a < 3; ++a)
          for (int c = 0; c < 3; ++c) {
            const double sv = tot[8 + a * 3 + c] - tot[5 + a] * mx[c] - my[a] * tot[2 + c] + tot[0] * my[a] * mx[c];
            S[a][c] = (Kb > 0) ? sv * (double)invf : 0.0;
          }
        float R[9], t[3];
        double cond;
        if (trip == 1) PSTAMPF(808);
        kabsch_solve(S, mx, my, R, t, &cond);
        if (trip == 1) {
          PSTAMPF(809);
          finish_pose(p, b, R, t, cond);
          __threadfence();
          red_release_gpu_add_u32(sync + 3, 1u);   // the pose is public: every CTA warps its rows now
        } else if (cond < 0.0) {
          p.partials[0] = (double)(R[0] + t[0]);   // never true (a condition number is >= 1): keeps the warm-up alive
        }
      }
      __syncwarp();
    }
    if (p.sel_w) {   // (tests / diagnostics) unused slots of the selection outputs
      const unsigned int ne = min(__ldcg(&p.state[b].sel_count), (unsigned int)p.K_max);
      for (int k = (int)ne + lane; k < p.K_max; k += 32) {
        p.sel_w[(size_t)b * p.K_max + k] = 0.f;
        p.sel_src[(size_t)b * p.K_max + k] = 0;
        p.sel_tgt[(size_t)b * p.K_max + k] = 0;
      }
    }
    return;
  }
  // the histogram is requested before anything else (it does not depend on the state record read next)
  for (int q = tid; q < TK_BINS; q += PP_THREADS) sc.hist[q] = __ldcg(p.cand_hist + (size_t)b * TK_BINS + q);
  const ProcrState st = p.state[b];
  const int Kb = st.Kb;
  size_t n = st.n_cand;
  const int row_lo = (int)(((long long)N * g) / G), row_hi = (int)(((long long)N * (g + 1)) / G);
  if (tid == 0) {
    list_n = 0u;
    mine_n = 0u;
    surv_n = 0u;
    T_s = 0ull;
  }
  bool slow = false;
  if (Kb > 0 && n < (size_t)Kb) {
    // fallback: the bound left too few candidates; every CTA rebuilds its stripe of the list from the whole matrix
    for (size_t e = (size_t)g * PP_THREADS + tid; e < total; e += (size_t)G * PP_THREADS) {
      ckey[e] = float_to_ordered(conf_at(p, b, e));
      cidx[e] = (unsigned int)e;
    }
    n = total;
    slow = true;
    WSYNC();
    if (tid == 0) {
      red_release_gpu_add_u32(sync + 0, 1u);
      while (ld_acquire_gpu_u32(sync + 0) < (unsigned int)G) {
      }
    }
  }
  WSYNC();
  const bool select_some = Kb > 0 && (size_t)Kb < n;   // otherwise every candidate is used (T = 0)
  int bin = 0;
  unsigned int want = 0u, hsel = 0u;
  if (select_some && !slow) {
    if (tid < 32) {
      int bn;
      unsigned int cum, hs;
      warp_walk_hist(sc.hist, (unsigned int)Kb, bn, cum, hs);
      if (tid == 0) {
        bin_s = bn;
        cum_s = cum;
        hsel_s = hs;
      }
    }
    WSYNC();
    bin = bin_s;
    want = (unsigned int)Kb - cum_s;  // how many of the crossing bin's candidates are selected
    hsel = hsel_s;
    if (hsel > (unsigned int)SEL_LIST || want < 1u || want > hsel) slow = true;  // heavy ties: the general select
  }
  PSTAMP0(801);
  // ---- slow cases first: T from the general select on CTA 0, published through the state record.  `slow` is a function
  //      of the state and the histogram every CTA reads, so all CTAs of the batch element take this branch together.
  unsigned long long T = 0ull;
  if (select_some && slow) {
    if (g == 0) {
      T = block_select_kth<PP_THREADS>([&](size_t e) { return make_key64(ckey[e], cidx[e]); }, n, Kb, sc.hist, sc.ctl, SyncWorkers());
      if (tid == 0) {
        p.state[b].T = T;
        __threadfence();
        red_release_gpu_add_u32(sync + 1, 1u);
      }
    } else {
      if (tid == 0) {
        while (ld_acquire_gpu_u32(sync + 1) == 0u) {
        }
        T_s = *((volatile unsigned long long*)&p.state[b].T);
      }
      WSYNC();
      T = T_s;
    }
  }
  // ---- the candidate list (~16-25 k entries, L2-resident): crossing bin's short list + the selected candidates of my rows.
  //      Pass 1 reads the KEYS only (16-byte loads, eight in flight per thread) and keeps the positions of those in the
  //      crossing bin or above (one compare each: the bins are intervals of the key) in a shared-memory list, appended per
  //      warp (ballot + one atomic); pass 2 fetches key and index of those ~K_b survivors, all loads of a thread in flight
  //      together.  (Fetching an index inside pass 1 is a dependent load per candidate; fetching all of them doubles the bytes.)
  const unsigned int row_fi_lo = (unsigned int)row_lo * (unsigned int)M;
  const unsigned long long row_fi_hi = (unsigned long long)row_hi * (unsigned long long)M;
  if (Kb > 0) {
    const bool by_bin = select_some && !slow;   // bins above the crossing one are selected whole; T decides inside it
    auto take = [&](unsigned int k32, unsigned int fi) {
      if ((unsigned long long)fi >= (unsigned long long)row_fi_lo && (unsigned long long)fi < row_fi_hi) {
        const unsigned int pos = atomicAdd(&mine_n, 1u);   // <= K_b + SEL_LIST <= mine_cap entries
        if (pos < (unsigned int)p.mine_cap) mine_s[pos] = ((unsigned long long)fi << 32) | (unsigned long long)k32;
      }
    };
    if (!(select_some && slow)) {
      // key interval of the crossing bin: bin(k) = ((k - kmin) << sh) >> 21, clamped
      unsigned long long key_lo = 0ull, key_hi = 1ull << 32;
      if (by_bin) {
        const int sh = st.hist_sh;
        const unsigned long long one = (sh >= 64) ? 0ull : ((1ull << sh) - 1ull);
        if (bin > 0) key_lo = (unsigned long long)st.hist_kmin + ((((unsigned long long)bin << 21) + one) >> sh);
        if (bin < TK_BINS - 1) key_hi = (unsigned long long)st.hist_kmin + ((((unsigned long long)(bin + 1) << 21) + one) >> sh);
      }
      // The Sinkhorn tail leaves the list as one contiguous segment per producer CTA, each producer owning a row range:
      // then my rows' candidates sit in the few segments that overlap my rows and pass 1 only has to find the crossing
      // bin's handful of entries (whose indices the ranking needs).  Otherwise pass 1 keeps everything from the crossing
      // bin upwards and pass 2 sorts out the rows.
      const bool segmented = st.seg_G > 0 && st.seg_broken == 0u;
      const unsigned long long keep_hi = segmented ? key_hi : (1ull << 32);
      unsigned int* surv_s = reinterpret_cast<unsigned int*>(mine_s + p.mine_cap);  // [surv_cap] positions in the candidate list
      if (segmented) {
        const int Gp = st.seg_G;
        int gp_lo = (int)(((long long)row_lo * Gp) / N) - 1, gp_hi = (int)(((long long)row_hi * Gp) / N) + 1;
        gp_lo = max(gp_lo, 0);
        gp_hi = min(gp_hi, Gp - 1);
        for (int gp = gp_lo; gp <= gp_hi; ++gp) {
          const int r0 = (int)(((long long)N * gp) / Gp), r1 = (int)(((long long)N * (gp + 1)) / Gp);
          if (r1 <= row_lo || r0 >= row_hi) continue;   // (block-uniform)
          const uint2 seg = __ldcg(p.cand_seg + (size_t)b * NUM_SMS + gp);
          for (unsigned int q = tid; q < seg.y; q += PP_THREADS) {
            const unsigned int k32 = __ldcg(ckey + seg.x + q);
            const unsigned int fi = __ldcg(cidx + seg.x + q);
            if ((unsigned long long)k32 >= key_lo) take(k32, fi);
          }
        }
      }
      if (by_bin || !segmented) {
        auto keep = [&](const unsigned int (&k)[32], const unsigned int (&e)[32], int cntv) {
          // warp-collective append of this thread's passing keys: one scan + one atomic per warp and round
          unsigned int mask = 0u;
#pragma unroll
          for (int u = 0; u < 32; ++u)
            if (u < cntv && e[u] != 0xFFFFFFFFu && (unsigned long long)k[u] >= key_lo && (unsigned long long)k[u] < keep_hi) mask |= 1u << u;
          const int mine_c = __popc(mask);
          int incl = mine_c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int tmp = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += tmp;
          }
          const int wtot = __shfl_sync(0xffffffffu, incl, 31);
          if (wtot == 0) return;
          unsigned int base = 0u;
          if (lane == 31) base = atomicAdd(&surv_n, (unsigned int)wtot);
          base = __shfl_sync(0xffffffffu, base, 31);
          unsigned int pos = base + (unsigned int)(incl - mine_c);
#pragma unroll
          for (int u = 0; u < 32; ++u)
            if (mask & (1u << u)) {
              if (pos < (unsigned int)p.surv_cap) surv_s[pos] = e[u];
              ++pos;
            }
        };
        const bool vec4 = (((uintptr_t)ckey) & 15u) == 0;
        const size_t n4 = vec4 ? (n >> 2) : 0;
        const uint4* k4 = reinterpret_cast<const uint4*>(ckey);
        for (size_t q0 = 0; q0 < n4; q0 += (size_t)PP_THREADS * 8) {   // (block-uniform trip count: the appends are warp-collective)
          uint4 kk[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const size_t q = q0 + (size_t)u * PP_THREADS + tid;
            kk[u] = q < n4 ? __ldcg(k4 + q) : make_uint4(0u, 0u, 0u, 0u);
          }
          unsigned int k[32], e[32];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const size_t q = q0 + (size_t)u * PP_THREADS + tid;
            const bool v = q < n4;
            // (an out-of-range slot gets key 0 with key_lo > 0, or passes harmlessly when everything is kept: guard by e)
            k[4 * u + 0] = v ? kk[u].x : 0u; k[4 * u + 1] = v ? kk[u].y : 0u; k[4 * u + 2] = v ? kk[u].z : 0u; k[4 * u + 3] = v ? kk[u].w : 0u;
            e[4 * u + 0] = v ? (unsigned int)(q << 2) + 0u : 0xFFFFFFFFu; e[4 * u + 1] = v ? (unsigned int)(q << 2) + 1u : 0xFFFFFFFFu;
            e[4 * u + 2] = v ? (unsigned int)(q << 2) + 2u : 0xFFFFFFFFu; e[4 * u + 3] = v ? (unsigned int)(q << 2) + 3u : 0xFFFFFFFFu;
          }
          keep(k, e, 32);
        }
        for (size_t e0 = (n4 << 2); e0 < n; e0 += PP_THREADS) {
          const size_t ee = e0 + tid;
          unsigned int k[32], e[32];
#pragma unroll
          for (int u = 0; u < 32; ++u) {
            k[u] = 0u;
            e[u] = 0xFFFFFFFFu;
          }
          if (ee < n) {
            k[0] = __ldcg(ckey + ee);
            e[0] = (unsigned int)ee;
          }
          keep(k, e, 1);
        }
        WSYNC();
        const unsigned int ns = min(surv_n, (unsigned int)p.surv_cap);
        for (unsigned int q0 = 0; q0 < ns; q0 += PP_THREADS * 4) {
          unsigned int kk[4], fi[4];
          bool on[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const unsigned int q = q0 + u * PP_THREADS + tid;
            const unsigned int e = q < ns ? surv_s[q] : 0xFFFFFFFFu;
            on[u] = e != 0xFFFFFFFFu;
            kk[u] = on[u] ? __ldcg(ckey + e) : 0u;
            fi[u] = on[u] ? __ldcg(cidx + e) : 0u;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (on[u]) {
              if (by_bin && (unsigned long long)kk[u] < key_hi) {   // the crossing bin
                const unsigned int pos = atomicAdd(&list_n, 1u);
                if (pos < (unsigned int)SEL_LIST) list_s[pos] = make_key64(kk[u], fi[u]);
              }
              if (!segmented) take(kk[u], fi[u]);
            }
          }
        }
      }
    } else {
      // general-select case: T is known; a candidate's index is needed to compare its 64-bit key (rare and slow anyway)
      for (size_t e = tid; e < n; e += PP_THREADS) {
        const unsigned int k32 = __ldcg(ckey + e);
        if (k32 < (unsigned int)(T >> 32)) continue;
        const unsigned int fi = __ldcg(cidx + e);
        if (make_key64(k32, fi) >= T) take(k32, fi);
      }
    }
  }
  WSYNC();
  PSTAMP0(802);
  // ---- T of the normal case: rank the crossing bin's short list (64-bit keys are distinct: distinct indices)
  if (select_some && !slow) {
    const unsigned int L = min(list_n, (unsigned int)SEL_LIST);
    if ((unsigned int)tid < L) {
      const unsigned long long my = list_s[tid];
      unsigned int rank = 0u;
      for (unsigned int q = 0; q < L; ++q) rank += (list_s[q] > my) ? 1u : 0u;
      if (rank == want - 1u) T_s = my;
    }
    WSYNC();
    T = T_s;
  }
  if (g == 0 && tid == 0) {
    p.state[b].T = T;
    p.state[b].n_cand = (unsigned int)n;
    p.state[b].pad_ = (slow ? 0x80000000u : 0u) | (hsel & 0x7FFFFFFFu);  // diagnostics
  }
  PSTAMP0(803);
  // ---- my rows' selected candidates in ascending flat-index order (the appends above came in atomic order)
  const unsigned int cnt_raw = min(mine_n, (unsigned int)p.mine_cap);  // (mine_cap >= K_b + SEL_LIST: no overflow)
  WSYNC();
  if (tid == 0) mine_n = 0u;
  unsigned long long* sorted_s = mine_s;
  if (cnt_raw <= (unsigned int)PP_RANK_SORT) {
    // few entries (the normal case: ~K_b / G): rank every selected entry by counting, scatter into the upper half
    unsigned long long mv[PP_RANK_SORT / PP_THREADS];
    unsigned int rk[PP_RANK_SORT / PP_THREADS];
#pragma unroll
    for (int u = 0; u < PP_RANK_SORT / PP_THREADS; ++u) {
      const unsigned int q = tid + u * PP_THREADS;
      mv[u] = q < cnt_raw ? mine_s[q] : PP_SENTINEL;
      if (mv[u] != PP_SENTINEL && make_key64((unsigned int)mv[u], (unsigned int)(mv[u] >> 32)) < T) mv[u] = PP_SENTINEL;  // not among the K_b best
      rk[u] = 0u;
    }
    WSYNC();
#pragma unroll
    for (int u = 0; u < PP_RANK_SORT / PP_THREADS; ++u) {
      const unsigned int q = tid + u * PP_THREADS;
      if (q < cnt_raw) mine_s[q] = mv[u];
    }
    WSYNC();
    for (unsigned int q = 0; q < cnt_raw; ++q) {
      const unsigned long long o = mine_s[q];  // broadcast read; sentinels are larger than every real entry
#pragma unroll
      for (int u = 0; u < PP_RANK_SORT / PP_THREADS; ++u) rk[u] += (o < mv[u]) ? 1u : 0u;
    }
    sorted_s = mine_s + PP_RANK_SORT;   // mine_cap >= 2 * PP_RANK_SORT (host)
    unsigned int real = 0u;
#pragma unroll
    for (int u = 0; u < PP_RANK_SORT / PP_THREADS; ++u) {
      if (mv[u] != PP_SENTINEL) {
        sorted_s[rk[u]] = mv[u];   // distinct flat indices: a permutation of 0 .. count-1
        ++real;
      }
    }
    real = __reduce_add_sync(0xffffffffu, real);
    if (lane == 0 && real) atomicAdd(&mine_n, real);
    WSYNC();
  } else {
    // many entries in one CTA's rows (concentrated confidences): bitonic sort in place, sentinels sort to the end
    unsigned int npow = 32u;
    while (npow < cnt_raw) npow <<= 1;
    for (unsigned int q = tid; q < npow; q += PP_THREADS) {
      unsigned long long v = q < cnt_raw ? mine_s[q] : PP_SENTINEL;
      if (q < cnt_raw && make_key64((unsigned int)v, (unsigned int)(v >> 32)) < T) v = PP_SENTINEL;
      mine_s[q] = v;
    }
    WSYNC();
    for (unsigned int k = 2u; k <= npow; k <<= 1) {
      for (unsigned int j = k >> 1; j > 0u; j >>= 1) {
        for (unsigned int q = tid; q < npow; q += PP_THREADS) {
          const unsigned int partner = q ^ j;
          if (partner > q) {
            const unsigned long long a = mine_s[q], c = mine_s[partner];
            const bool up = (q & k) == 0u;
            if ((a > c) == up) {
              mine_s[q] = c;
              mine_s[partner] = a;
            }
          }
        }
        WSYNC();
      }
    }
    unsigned int c = 0u;
    for (unsigned int q = tid; q < npow; q += PP_THREADS) c += (mine_s[q] != PP_SENTINEL) ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0 && c) atomicAdd(&mine_n, c);
    WSYNC();
  }
  const unsigned int cnt = mine_n;
  PSTAMP0(804);
  const float* sp = p.src_pcd + (size_t)b * N * 3;
  const float* tp = p.tgt_pcd + (size_t)b * M * 3;
  // ---- weighted moments of my entries, ONE pass: sum w, sum |w|, sum w x, sum w y, sum w y x^T in fp64 (products of fp32
  //      values are exact there, so no centring pass is needed: the covariance about the weighted means is formed from
  //      these sums at the end).  ~K_b / G entries per CTA: the fp64 work is a few hundred DFMAs per SM.
  //      Thread t takes sorted entries t, t + 256, ...; lanes, warps and CTAs are added in fixed order: bit-reproducible.
  double acc[PM_SUMS];
#pragma unroll
  for (int k = 0; k < PM_SUMS; ++k) acc[k] = 0.0;
  for (unsigned int q = tid; q < cnt; q += PP_THREADS) {
    const unsigned long long v = sorted_s[q];
    const unsigned int fi = (unsigned int)(v >> 32);
    const float wf = ordered_to_float((unsigned int)v);
    const int i = (int)(fi / (unsigned int)M), j = (int)(fi - (unsigned int)i * (unsigned int)M);
    if (p.sel_w) {
      const unsigned int pos = atomicAdd(&p.state[b].sel_count, 1u);
      if (pos < (unsigned int)p.K_max) {
        p.sel_w[(size_t)b * p.K_max + pos] = wf;
        p.sel_src[(size_t)b * p.K_max + pos] = i;
        p.sel_tgt[(size_t)b * p.K_max + pos] = j;
      }
    }
    const double w = (double)wf;
    const double xs[3] = {(double)sp[i * 3 + 0], (double)sp[i * 3 + 1], (double)sp[i * 3 + 2]};
    const double ys[3] = {(double)tp[j * 3 + 0], (double)tp[j * 3 + 1], (double)tp[j * 3 + 2]};
    acc[0] += w;
    acc[1] += fabs(w);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      acc[2 + a] = fma(w, xs[a], acc[2 + a]);
      const double wy = w * ys[a];
      acc[5 + a] += wy;
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[8 + a * 3 + c] = fma(wy, xs[c], acc[8 + a * 3 + c]);
    }
  }
  PSTAMP0(805);
  // lanes (fixed shuffle tree), then warps (fixed order)
#pragma unroll
  for (int k = 0; k < PM_SUMS; ++k) {
    double x = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) wsum_s[warp][k] = x;
  }
  WSYNC();
  double* part = p.partials + ((size_t)b * PP_MAX_G + g) * PM_PART;
  if (tid < PM_SUMS) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < PP_THREADS / 32; ++w) t += wsum_s[w][tid];
    part[tid] = t;
  }
  WSYNC();
  if (tid == 0) {
    __threadfence();
    red_release_gpu_add_u32(sync + 2, 1u);   // this CTA's sums are public (the finisher waits for G arrivals)
    PSTAMP0(806);
    while (ld_acquire_gpu_u32(sync + 3) == 0u) {
    }
  }
  WSYNC();
  // ---- warp my rows' source points with the gated pose:  (R_forwd s + t_forwd)     pipeline.py:220
  if (p.src_warped) {
    if (tid < 12) pose_s[tid] = tid < 9 ? __ldcg(p.R_forwd + b * 9 + tid) : __ldcg(p.t_forwd + b * 3 + (tid - 9));
    WSYNC();
    float* o = p.src_warped + (size_t)b * N * 3;
    for (int i = row_lo + tid; i < row_hi; i += PP_THREADS) {
      const float x0 = sp[i * 3 + 0], x1 = sp[i * 3 + 1], x2 = sp[i * 3 + 2];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        // same association as a 3-term dot product followed by the translation add
        float acc = pose_s[a * 3 + 0] * x0;
        acc = fmaf(pose_s[a * 3 + 1], x1, acc);
        acc = fmaf(pose_s[a * 3 + 2], x2, acc);
        o[i * 3 + a] = acc + pose_s[9 + a];
      }
    }
  }
  PSTAMP0(810);
}
#undef PSTAMP0
#undef PSTAMPL
#undef PSTAMPF

// standalone weighted Kabsch on given correspondences: X, Y [B,K,3], w [B,K]
struct KabschParams {
  const float* X;
  const float* Y;
  const float* w;
  int B, K;
  float eps;
  float* R;
  float* t;
  double* condition;
};

__global__ void __launch_bounds__(256) kabsch_kernel(const KabschParams p) {
  __shared__ MomentScratch ms;
  const int b = blockIdx.x;
  const float* X = p.X + (size_t)b * p.K * 3;
  const float* Y = p.Y + (size_t)b * p.K * 3;
  const float* w = p.w + (size_t)b * p.K;
  float m1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
    const float wk = w[k];
    m1[0] += wk;
    m1[1] += fabsf(wk);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      m1[2 + a] = fmaf(wk, X[k * 3 + a], m1[2 + a]);
      m1[5 + a] = fmaf(wk, Y[k * 3 + a], m1[5 + a]);
    }
  }
  double s1[8];
  block_sum_f32<8>(m1, s1, ms);
  const double inv = 1.0 / (s1[1] + (double)p.eps);
  const float invf = (float)inv;
  const float mxf[3] = {(float)(s1[2] * inv), (float)(s1[3] * inv), (float)(s1[4] * inv)};
  const float myf[3] = {(float)(s1[5] * inv), (float)(s1[6] * inv), (float)(s1[7] * inv)};
  float m2[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
    const float wn = w[k] * invf;
    const float xc[3] = {X[k * 3 + 0] - mxf[0], X[k * 3 + 1] - mxf[1], X[k * 3 + 2] - mxf[2]};
    const float yc[3] = {Y[k * 3 + 0] - myf[0], Y[k * 3 + 1] - myf[1], Y[k * 3 + 2] - myf[2]};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) m2[a * 3 + c] = fmaf(wn * yc[a], xc[c], m2[a * 3 + c]);
  }
  double s2[9];
  block_sum_f32<9>(m2, s2, ms);
  if (threadIdx.x == 0) {
    float R[9], t[3];
    double cond;
    double S[3][3];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) S[a][c] = s2[a * 3 + c];
    const double mx[3] = {(double)mxf[0], (double)mxf[1], (double)mxf[2]}, my[3] = {(double)myf[0], (double)myf[1], (double)myf[2]};
    kabsch_solve(S, mx, my, R, t, &cond);
    for (int k = 0; k < 9; ++k) p.R[b * 9 + k] = R[k];
    for (int k = 0; k < 3; ++k) p.t[b * 3 + k] = t[k];
    p.condition[b] = cond;
  }
}


struct ProcrWorkspace {
  ProcrState* state;
  unsigned int* cand_key;
  unsigned int* cand_idx;
  unsigned int* sample_buf;
  unsigned int* sample_arrive;
  unsigned int* pose_sync;
  unsigned int* cand_hist;
  unsigned int* sample_hist;
  uint2* cand_seg;
  unsigned int* sample_hist2;
  float* sample_val;
  unsigned long long* sample_list;
  unsigned int* sample_list_n;
  double* partials;
  size_t total;
};

static ProcrWorkspace procr_carve(void* ws, int B, int N, int M) {
  ProcrWorkspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* ptr = ws ? (void*)((char*)ws + off) : nullptr;
    off += align_up(bytes, 256);
    return ptr;
  };
  w.state = (ProcrState*)take(sizeof(ProcrState) * B);
  w.cand_key = (unsigned int*)take(4ull * B * N * M);
  w.cand_idx = (unsigned int*)take(4ull * B * N * M);
  w.sample_buf = (unsigned int*)take(4ull * B * TS_SAMPLES);
  w.sample_arrive = (unsigned int*)take(4ull * 5 * B);  // sampling CTAs [B], then the pose kernel's words [B][4]: one memset
  w.pose_sync = w.sample_arrive ? w.sample_arrive + B : nullptr;
  w.cand_hist = (unsigned int*)take(4ull * B * TK_BINS);
  w.sample_hist = (unsigned int*)take(4ull * B * TK_BINS);  // fused Sinkhorn -> collect path: histogram of the sampled log2 confidences
  w.cand_seg = (uint2*)take(8ull * B * NUM_SMS);
  w.sample_hist2 = (unsigned int*)take(4ull * B * TK_BINS);
  w.sample_val = (float*)take(4ull * B * 16 * (((size_t)M + 1 + 3) & ~(size_t)3));   // [B][SH_ROWS = 16][pitch4(M + 1)]
  w.sample_list = (unsigned long long*)take(8ull * B * 16 * (((size_t)M + 1 + 3) & ~(size_t)3));
  w.sample_list_n = (unsigned int*)take(4ull * B);
  w.partials = (double*)take(8ull * B * PP_MAX_G * PM_PART);
  w.total = off;
  return w;
}

}  // namespace drg

using namespace drg;

extern "C" size_t drg_soft_procrustes_workspace_bytes(int B, int N, int M) {
  if (B < 1 || N < 1 || M < 1) return 0;
  return procr_carve(nullptr, B, N, M).total;
}

struct ProcrSource {  // potentials mode inputs (conf == NULL)
  const float* scores;
  const float* shift;
  int apply_mask;
  SkhViews views;
};

static int pow2_ceil(int x) {
  int v = 32;
  while (v < x) v <<= 1;
  return v;
}

// validate the arguments and describe the problem / workspace to the kernels
static int procr_setup(const drg_procrustes_args* a, bool potentials, void* workspace, size_t workspace_bytes, ProcrParams& p,
                       ProcrWorkspace& w) {
  DRG_CHECK_ARG(a != nullptr, "args is null");
  DRG_CHECK_ARG((a->conf != nullptr || potentials) && a->src_pcd && a->tgt_pcd, "conf/src_pcd/tgt_pcd must be non-null");
  DRG_CHECK_ARG(a->padded_lengths || (a->src_mask && a->tgt_mask), "masks must be non-null unless padded_lengths is set");
  DRG_CHECK_ARG(a->B >= 1 && a->B <= 1024 && a->N >= 1 && a->M >= 1, "need 1 <= B <= 1024 and N, M >= 1");
  DRG_CHECK_ARG((long long)a->N * a->M < (1ll << 32), "N*M must fit in 32 bits");
  DRG_CHECK_ARG(a->R && a->t && a->R_forwd && a->t_forwd && a->condition && a->solution_mask, "pose outputs must be non-null");
  DRG_CHECK_ARG(a->sample_rate > 0.f, "sample_rate must be > 0");
  DRG_CHECK_ARG((a->sel_w == nullptr) == (a->sel_src == nullptr) && (a->sel_w == nullptr) == (a->sel_tgt == nullptr),
                "sel_w/sel_src/sel_tgt must be given together");
  const int B = a->B, N = a->N, M = a->M;
  w = procr_carve(workspace, B, N, M);
  if (workspace == nullptr || workspace_bytes < w.total || ((uintptr_t)workspace & 255u)) {
    set_error("soft_procrustes: workspace missing, too small (%zu < %zu) or not 256-byte aligned", workspace_bytes, w.total);
    return DRG_ERR_WORKSPACE;
  }
  p = ProcrParams{};
  p.conf = potentials ? nullptr : a->conf;
  p.src_pcd = a->src_pcd;
  p.tgt_pcd = a->tgt_pcd;
  p.src_mask = a->src_mask;
  p.tgt_mask = a->tgt_mask;
  p.B = B;
  p.N = N;
  p.M = M;
  p.sample_rate = a->sample_rate;
  p.max_condition_num = a->max_condition_num;
  p.padded_lengths = a->padded_lengths;
  p.state = w.state;
  p.cand_key = w.cand_key;
  p.cand_idx = w.cand_idx;
  p.sample_buf = w.sample_buf;
  p.sample_arrive = w.sample_arrive;
  p.pose_sync = w.pose_sync;
  p.cand_hist = w.cand_hist;
  p.cand_seg = w.cand_seg;
  p.partials = w.partials;
  p.R = a->R;
  p.t = a->t;
  p.R_forwd = a->R_forwd;
  p.t_forwd = a->t_forwd;
  p.condition = a->condition;
  p.solution_mask = a->solution_mask;
  p.src_warped = a->src_warped;
  // K_b <= max(N, M) * sample_rate
  const long long kmax_ll = (long long)((double)(N > M ? N : M) * (double)a->sample_rate) + 1;
  p.K_max = a->sel_w ? a->K_max : (int)(kmax_ll < (long long)N * M ? kmax_ll : (long long)N * M);
  DRG_CHECK_ARG(p.K_max >= 1, "K_max must be >= 1");
  p.sel_w = a->sel_w;
  p.sel_src = a->sel_src;
  p.sel_tgt = a->sel_tgt;
  p.stamps = g_tuning_stamps;
  // the pose kernel keeps a CTA's selected candidates (at most K_b + the crossing bin's short list) in shared memory
  p.mine_cap = pow2_ceil((int)(kmax_ll < (long long)N * M ? kmax_ll : (long long)N * M) + SEL_LIST);
  if (p.mine_cap < 2 * PP_RANK_SORT) p.mine_cap = 2 * PP_RANK_SORT;  // the rank sort scatters into the upper half
  p.surv_cap = (int)(kmax_ll < (long long)N * M ? kmax_ll : (long long)N * M) + SEL_LIST + 64;
  if ((size_t)p.mine_cap * 8 + (size_t)p.surv_cap * 4 > 220 * 1024) {
    set_error("soft_procrustes: max(N, M) * sample_rate = %lld correspondences exceed the pose kernel's shared-memory lists (about 16000)",
              kmax_ll);
    return DRG_ERR_UNSUPPORTED;
  }
  return DRG_OK;
}

// stand-alone candidate search (a stored confidence matrix, or shapes the persistent Sinkhorn does not take)
static int procr_search(const ProcrParams& p, cudaStream_t st) {
  const int B = p.B, N = p.N, M = p.M;
  DRG_CUDA(cudaFuncSetAttribute(topk_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SAMPLES * 4));
  int ts_ctas = NUM_SMS / (2 * B);
  if (ts_ctas > 32) ts_ctas = 32;
  if (ts_ctas < 1) ts_ctas = 1;
  {
    ProfScope prof_scope(PROF_TOPK_THRESHOLD, st);
    topk_threshold_kernel<<<dim3(ts_ctas, B), TS_THREADS, TS_SAMPLES * 4, st>>>(p);
  }
  DRG_LAUNCH_CHECK();
  int gx = (NUM_SMS * 8) / B;
  if (gx < 1) gx = 1;
  const long long n4 = ((long long)N * M + 3) / 4;
  if ((long long)gx * 256 > n4) gx = (int)((n4 + 255) / 256);
  {
    ProfScope prof_scope(PROF_TOPK_COLLECT, st);
    if (!p.conf && (M % 4) == 0 && ((((uintptr_t)p.scores) | ((uintptr_t)p.pv) | ((uintptr_t)p.tgt_mask)) & 15u) == 0 &&
        ((p.ldv & 3) == 0)) {
      // two CTAs per SM: one resident wave with ~14 rows each measured best at 4096^2 (1 -> 43.7 us, 2 -> 28.9 us, 3 -> 35.1 us)
      int gr = (NUM_SMS * 2) / B;
      if (gr < 1) gr = 1;
      if (gr > N) gr = N;
      topk_collect_rows_kernel<<<dim3(gr, B), 256, 0, st>>>(p);
    } else {
      topk_collect_kernel<<<dim3(gx, B), 256, 0, st>>>(p);
    }
  }
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

static int procr_pose(const ProcrParams& p, cudaStream_t st) {
  int G = NUM_SMS / p.B;
  if (G > PP_MAX_G) G = PP_MAX_G;
  if (G > p.N) G = p.N;
  if (G < 1) G = 1;
  const size_t smem = (size_t)p.mine_cap * 8 + (size_t)p.surv_cap * 4;
  DRG_CUDA(cudaFuncSetAttribute(procr_pose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof_scope(PROF_PROCR_SOLVE, st);
  if ((long long)G * p.B > NUM_SMS || G == 1) {
    // one working CTA per batch element never waits for another CTA: an ordinary launch, any batch size
    procr_pose_kernel<<<dim3(1, p.B), PP_THREADS + 32, smem, st>>>(p);
    DRG_LAUNCH_CHECK();
    return DRG_OK;
  }
  // several CTAs per batch element wait for each other on the rare paths (general select, whole-matrix fallback): the
  // cooperative launch guarantees that they are co-resident
  ProcrParams pp = p;
  void* args[] = {(void*)&pp};
  DRG_CUDA(cudaLaunchCooperativeKernel((const void*)procr_pose_kernel, dim3(G, p.B), dim3(PP_THREADS + 32), args, smem, st));
  count_launch();
  return DRG_OK;
}

extern "C" int drg_soft_procrustes(const drg_procrustes_args* a, void* workspace, size_t workspace_bytes, void* stream) {
  DRG_CHECK_ARG(a != nullptr && a->conf != nullptr, "args / conf is null");
  ProcrParams p;
  ProcrWorkspace w;
  int rc = procr_setup(a, false, workspace, workspace_bytes, p, w);
  if (rc != DRG_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // the arrival counters / flags must be zero on entry (the workspace is caller memory of unknown content)
  DRG_CUDA(cudaMemsetAsync(w.sample_arrive, 0, 4ull * 5 * a->B, st));
  rc = procr_search(p, st);
  if (rc != DRG_OK) return rc;
  return procr_pose(p, st);
}

extern "C" int drg_sinkhorn_soft_procrustes(const drg_sinkhorn_args* s, const drg_procrustes_args* a, void* skh_workspace,
                                            size_t skh_workspace_bytes, void* procr_workspace, size_t procr_workspace_bytes,
                                            void* stream) {
  DRG_CHECK_ARG(s != nullptr && a != nullptr, "args are null");
  DRG_CHECK_ARG(s->out_mode == DRG_OUT_NONE, "the fused call takes out_mode DRG_OUT_NONE: the confidence matrix is never written");
  DRG_CHECK_ARG(s->B == a->B && s->N == a->N && s->M == a->M, "sinkhorn and procrustes shapes differ");
  DRG_CHECK_ARG(a->src_mask == s->src_mask && a->tgt_mask == s->tgt_mask, "the fused call uses one pair of masks");
  ProcrParams p;
  ProcrWorkspace w;
  int rc = procr_setup(a, true, procr_workspace, procr_workspace_bytes, p, w);
  if (rc != DRG_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  DRG_CUDA(cudaMemsetAsync(w.sample_arrive, 0, 4ull * 5 * a->B, st));
  // Sinkhorn; when the persistent kernel takes the shape, the candidate search runs as its last phase
  SkhCollect col{};
  col.state = w.state;
  col.cand_key = w.cand_key;
  col.cand_idx = w.cand_idx;
  col.cand_hist = w.cand_hist;
  col.sample_hist = w.sample_hist;
  col.cand_seg = w.cand_seg;
  col.sample_hist2 = w.sample_hist2;
  col.sample_val = w.sample_val;
  col.sample_list = w.sample_list;
  col.sample_list_n = w.sample_list_n;
  col.sample_rate = a->sample_rate;
  col.padded_lengths = a->padded_lengths;
  col.K_max = p.K_max;
  bool collected = false;
  SkhViews views{};
  rc = skh_run_with_views(s, skh_workspace, skh_workspace_bytes, stream, &views, &col, &collected);
  if (rc != DRG_OK) return rc;
  p.scores = s->scores;
  p.pu = views.u;
  p.pv = views.v;
  p.pbc = views.bc;
  p.pshift = s->shift;
  p.ldu = views.ldu;
  p.ldv = views.ldv;
  p.apply_mask = s->apply_mask;
  if (!collected) {
    rc = procr_search(p, st);
    if (rc != DRG_OK) return rc;
  }
  return procr_pose(p, st);
}

extern "C" int drg_weighted_procrustes(const float* X, const float* Y, const float* w, int B, int K, float eps, float* R, float* t,
                                       double* condition, void* stream) {
  DRG_CHECK_ARG(X && Y && w && R && t && condition, "X/Y/w/R/t/condition must be non-null");
  DRG_CHECK_ARG(B >= 1 && K >= 1, "B, K must be >= 1");
  KabschParams p{X, Y, w, B, K, eps, R, t, condition};
  kabsch_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(p);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}
