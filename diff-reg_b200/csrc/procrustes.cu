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

constexpr int TK_BINS = 2048;
constexpr int TS_THREADS = 512;
constexpr int TS_SAMPLES = 32768;  // sample keys held in shared memory (128 KB)
constexpr int TS_FAST_TARGET = 128;  // order statistics up to this rank use the thread-maxima short cut of the bound select
constexpr int TS_FAST_LIST = 256;    // capacity of its short list (TS_THREADS threads rank it)
constexpr int SOLVE_THREADS = 1024;
constexpr int SOLVE_SMEM_CAND = 24576;  // candidates staged in shared memory by the solve kernel (192 KB)

constexpr int PM_THREADS = 256;   // multi-CTA moments kernel
constexpr int PM_MAX_G = 64;      // its CTAs per batch element (at most)
constexpr int PM_PART = 24;       // doubles per CTA partial: W, W_abs, cx[3], cy[3], D[3], E[3], C[9]
constexpr int SEL_THREADS = 1024; // select kernel
constexpr int SEL_LIST = 256;     // short list of the select kernel (candidates of the crossing histogram bin)

struct ProcrState {  // per batch element
  int Kb;                         // number of correspondences to use
  unsigned int n_cand;            // candidates appended so far
  unsigned long long lower_key;   // candidates have 64-bit key >= lower_key
  unsigned long long T;           // written by the select kernel: the Kb best candidates are those with key >= T
  unsigned int hist_kmin;         // candidate histogram (filled by the collect kernels): bin = ((value key - kmin) << sh) >> 21,
  int hist_sh;                    //   clamped to the top bin -- monotone in the key, ~2048 bins over [bound, 2 x sample range]
  unsigned int sel_count;         // selected correspondences written to sel_* so far
  unsigned int pad_;
};

__device__ __forceinline__ unsigned int cand_bin(unsigned int k32, unsigned int kmin, int sh) {
  const unsigned long long d = ((unsigned long long)(k32 - kmin) << sh) >> 21;  // k32 >= kmin for every candidate
  return d > (unsigned long long)(TK_BINS - 1) ? (unsigned int)(TK_BINS - 1) : (unsigned int)d;
}

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
  unsigned int* sample_arrive;   // [B] arrival counters of the sampling CTAs (zero between calls)
  unsigned int* moments_arrive;  // [B] arrival counters of the moments CTAs (zero between calls)
  unsigned int* cand_hist;       // [B, TK_BINS] histogram of the candidates' value keys (zeroed by the threshold kernel)
  double* partials;              // [B, PM_MAX_G, PM_PART] per-CTA moment partials
  float4* pcd4;                  // [B][N + M] the points padded to 16 bytes (src, then tgt): one gather per point in the solve kernel
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
  long long* dbg_times;          // tuning only (DRG_PROCR_TIMES=1): clock64 stamps of batch element 0
};

__device__ __forceinline__ unsigned int hash_u32(unsigned int x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

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

__device__ __forceinline__ unsigned long long make_key64(unsigned int ordered_value, unsigned int flat_index) {
  return ((unsigned long long)ordered_value << 32) | (unsigned long long)(0xFFFFFFFFu - flat_index);
}

// Exact selection of the k-th largest of n distinct 64-bit keys by one CTA (1 <= k <= n).
// key_at(e) returns the key of element e.  Six radix levels (11,11,10,11,11,10 bits, MSB first); stops
// early once the remaining bucket is wanted whole.  Returns T such that exactly k keys are >= T.
// (Measured alternatives that were slower on B200: warp-aggregated histogram updates via match.any, and
// normalising the keys to their common range first.)
struct __align__(16) SelectScratch {
  unsigned int hist[TK_BINS];
  unsigned long long prefix;
  int krem;
  int done;
};

// One warp walks a TK_BINS-bin histogram (shared memory) from the top and finds the bin where the running count reaches
// krem: `bin`, the count `cum` in the bins above it and its own count `hsel`, returned to all lanes.  Two steps, nothing
// kept in registers and no serial scan (an earlier version held a lane's 64 bins in registers -- spilled at 64 registers
// per thread -- and let the owning lane step through them one by one).  Step 1: lane l sums the l-th chunk of 64 bins
// (descending); a warp scan finds the chunk of the crossing.  Step 2: the 32 lanes split that chunk two bins each and
// scan again.  If krem exceeds the total count the lowest bin is returned.
__device__ __forceinline__ void warp_walk_hist(const unsigned int* hist, unsigned int krem, int& bin, unsigned int& cum,
                                               unsigned int& hsel) {
  const int lane = threadIdx.x & 31;
  constexpr int chunk = TK_BINS / 32;                    // 64 bins per lane
  const int lo = TK_BINS - chunk * (lane + 1);           // lowest bin of this lane's chunk
  const uint4* h4 = reinterpret_cast<const uint4*>(&hist[lo]);
  unsigned int local = 0u;
#pragma unroll
  for (int q = 0; q < chunk / 4; ++q) {
    const uint4 v4 = h4[q];
    local += v4.x + v4.y + v4.z + v4.w;
  }
  unsigned int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int tmp = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += tmp;
  }
  const unsigned int crossing = __ballot_sync(0xffffffffu, incl >= krem);
  const int owner = crossing ? (__ffs(crossing) - 1) : 31;
  const unsigned int before = __shfl_sync(0xffffffffu, incl - local, owner);  // keys in the chunks above the owner's
  const int top = TK_BINS - chunk * owner - 1;           // highest bin of the owner's chunk
  const unsigned int h0 = hist[top - 2 * lane], h1 = hist[top - 2 * lane - 1];
  const unsigned int pair = h0 + h1;
  unsigned int incl2 = pair;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int tmp = __shfl_up_sync(0xffffffffu, incl2, o);
    if (lane >= o) incl2 += tmp;
  }
  const unsigned int crossing2 = __ballot_sync(0xffffffffu, before + incl2 >= krem);
  const int lane2 = crossing2 ? (__ffs(crossing2) - 1) : 31;  // no crossing (krem beyond the count): the lowest bins
  const unsigned int c0 = before + incl2 - pair;
  const bool first = crossing2 != 0u && c0 + h0 >= krem;
  const int my_bin = first ? (top - 2 * lane) : (top - 2 * lane - 1);
  const unsigned int my_cum = first ? c0 : (c0 + h0);
  const unsigned int my_h = first ? h0 : h1;
  bin = __shfl_sync(0xffffffffu, my_bin, lane2);
  cum = __shfl_sync(0xffffffffu, my_cum, lane2);
  hsel = __shfl_sync(0xffffffffu, my_h, lane2);
}

template <int NT, class KeyAt>
__device__ unsigned long long block_select_kth(KeyAt key_at, size_t n, int k, SelectScratch& sc) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    sc.prefix = 0ull;
    sc.krem = k;
    sc.done = 0;
  }
  __syncthreads();
  int shift = 64;
  const int widths[6] = {11, 11, 10, 11, 11, 10};
  for (int level = 0; level < 6; ++level) {
    const int wbits = widths[level];
    shift -= wbits;
    for (int q = tid; q < TK_BINS; q += NT) sc.hist[q] = 0u;
    __syncthreads();
    const unsigned long long prefix = sc.prefix;
    const int hi_shift = shift + wbits;  // bits above the current digit
    for (size_t e = tid; e < n; e += NT) {
      const unsigned long long key = key_at(e);
      const bool match = (hi_shift >= 64) ? true : ((key >> hi_shift) == prefix);
      if (match) atomicAdd(&sc.hist[(unsigned int)((key >> shift) & ((1u << wbits) - 1u))], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      int dbin;
      unsigned int cum, hsel;
      const unsigned int krem = (unsigned int)sc.krem;
      warp_walk_hist(sc.hist, krem, dbin, cum, hsel);
      if (tid == 0) {
        sc.prefix = (prefix << wbits) | (unsigned long long)dbin;
        sc.krem = (int)(krem - cum);
        if (hsel == krem - cum) sc.done = 1;  // the whole bucket is wanted
      }
    }
    __syncthreads();
    if (sc.done) break;
  }
  const unsigned long long T = sc.prefix << shift;
  __syncthreads();
  return T;
}

// ---- 1. K_b and the sample-based lower bound ---------------------------------------------------
__global__ void __launch_bounds__(TS_THREADS) topk_threshold_kernel(const ProcrParams p) {
  extern __shared__ __align__(16) unsigned int sample_key[];  // [TS_SAMPLES]
  __shared__ SelectScratch sc;
  __shared__ unsigned long long cnt_s;
  const int b = blockIdx.y, tid = threadIdx.x;
#define TSTAMP(k) do { if (p.dbg_times && b == 0 && tid == 0) p.dbg_times[(k)] = clock64(); } while (0)
  TSTAMP(20);
  TSTAMP(21);
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
  TSTAMP(22);
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
        const unsigned long long tb = block_select_kth<TS_THREADS>(
            [&](size_t e) { return make_key64(tmax_s[e], (unsigned int)e); }, (size_t)TS_THREADS, (int)target, sc);
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
        lower = block_select_kth<TS_THREADS>(
            [&](size_t e) { return make_key64(sample_key[e], sample_pos((unsigned int)e)); }, (size_t)n_s, (int)target, sc);
      }
    }
  }
  TSTAMP(23);
#undef TSTAMP
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

// [N,3] / [M,3] points -> 16-byte records (the solve kernel then needs one gather per point instead of three; its two
// moment passes are bound by the L1 wavefronts of those scattered loads)
__device__ __forceinline__ void pad_points(const ProcrParams& p, int b) {
  const int L = p.N + p.M;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < L; q += gridDim.x * blockDim.x) {
    const float* src = (q < p.N) ? (p.src_pcd + ((size_t)b * p.N + q) * 3) : (p.tgt_pcd + ((size_t)b * p.M + (q - p.N)) * 3);
    p.pcd4[(size_t)b * L + q] = make_float4(src[0], src[1], src[2], 0.f);
  }
}

__global__ void __launch_bounds__(256) topk_collect_kernel(const ProcrParams p) {
  const int b = blockIdx.y;
  pad_points(p, b);
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
  pad_points(p, b);
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
__device__ void kabsch_solve(const double S[3][3], const double mx[3], const double my[3], float R_out[9], float t_out[3],
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

__global__ void __launch_bounds__(SOLVE_THREADS) procr_solve_kernel(const ProcrParams p) {
  __shared__ SelectScratch sc;
  __shared__ MomentScratch ms;
  __shared__ unsigned int warp_cnt[SOLVE_THREADS / 32];
  __shared__ double mean_s[6], cov_s[9];
  __shared__ float pose_s[12];
  __shared__ unsigned int range_s[2];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) {
    range_s[0] = 0xFFFFFFFFu;
    range_s[1] = 0u;
  }
  __syncthreads();
#define PSTAMP(k) do { if (p.dbg_times && b == 0 && tid == 0) p.dbg_times[(k)] = clock64(); } while (0)
  PSTAMP(0);
  const size_t total = (size_t)p.N * p.M;
  const ProcrState st = p.state[b];
  const int Kb = st.Kb;
  unsigned int* ckey = p.cand_key + (size_t)b * total;
  unsigned int* cidx = p.cand_idx + (size_t)b * total;
  size_t n = st.n_cand;
  extern __shared__ unsigned int cand_s[];  // [2][SOLVE_SMEM_CAND]: keys, indices
  if (n < (size_t)Kb) {
    // fallback: the sample-based bound left too few candidates; take the whole matrix
    for (size_t e = tid; e < total; e += SOLVE_THREADS) {
      ckey[e] = float_to_ordered(conf_at(p, b, e));
      cidx[e] = (unsigned int)e;
    }
    n = total;
    __syncthreads();
  }

  // The candidate list normally fits in shared memory: stage it once (coalesced, 16 loads in flight per thread) so that
  // the radix levels and the emission below do not pay a global-memory round trip per 1024 candidates.  The staging
  // pass also finds the range [kmin, kmax] of the 32-bit value keys: the select then runs on keys normalised to that
  // range (value - kmin, left-aligned), so its first level spreads the candidates over all 2048 bins instead of the
  // handful of bins that share the confidences' exponent (19 k atomics into ~4 addresses), and two levels normally do.
  constexpr int QMAX = SOLVE_SMEM_CAND / SOLVE_THREADS;
  const bool in_smem = n <= (size_t)SOLVE_SMEM_CAND;
  unsigned int kmin = 0u;
  int nsh = 0;
  if (in_smem) {
    unsigned int lo = 0xFFFFFFFFu, hi = 0u;
    for (int q0 = 0; q0 < QMAX; q0 += 8) {
      if ((size_t)q0 * SOLVE_THREADS >= n) break;  // uniform
      unsigned int kk[8], ii[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const size_t e = (size_t)tid + (size_t)(q0 + u) * SOLVE_THREADS;
        const bool in = e < n;
        kk[u] = in ? ckey[e] : 0u;
        ii[u] = in ? cidx[e] : 0u;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const size_t e = (size_t)tid + (size_t)(q0 + u) * SOLVE_THREADS;
        if (e < n) {
          cand_s[e] = kk[u];
          cand_s[SOLVE_SMEM_CAND + e] = ii[u];
          lo = min(lo, kk[u]);
          hi = max(hi, kk[u]);
        }
      }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((tid & 31) == 0) {
      atomicMin(&range_s[0], lo);
      atomicMax(&range_s[1], hi);
    }
    __syncthreads();
    kmin = range_s[0];
    const unsigned int kmax = range_s[1];
    nsh = (kmax > kmin) ? __clz((int)(kmax - kmin)) : 32;
  }
  PSTAMP(1);
  const unsigned int* kp = in_smem ? cand_s : ckey;
  const unsigned int* ip = in_smem ? cand_s + SOLVE_SMEM_CAND : cidx;
  // order-preserving on the candidates (all values >= kmin, value range below 2^(32 - nsh)); identity when not staged
  auto nkey = [&](unsigned int k32, unsigned int fi) -> unsigned long long { return make_key64(k32 - kmin, fi) << nsh; };

  // ---- exact radix select of the Kb largest 64-bit keys (value << 32 | ~index): no ties
  unsigned long long T = 0ull;  // select keys >= T
  if (Kb > 0 && (size_t)Kb < n)
    T = block_select_kth<SOLVE_THREADS>([&](size_t e) { return nkey(kp[e], ip[e]); }, n, Kb, sc);

  PSTAMP(2);
  if (p.dbg_times && b == 0 && tid == 0) p.dbg_times[10] = (long long)n;
  // ---- two fp32 passes over the selected candidates: weighted means, then the centred covariance -- the reference's
  //      own order of operations (procrustes.py:29-34).  A thread first marks which of its (at most QMAX) staged
  //      candidates are selected, then visits them four at a time so that the eight point gathers of a batch are in
  //      flight together (one L2 round trip per selected candidate was what bounded these passes).
  //      (Measured: compacting the selection first and fp64 moments were both slower on B200.)
  const float* sp = p.src_pcd + (size_t)b * p.N * 3;
  const float4* sp4 = p.pcd4 + (size_t)b * (p.N + p.M);  // written by the collect kernel
  const float4* tp4 = sp4 + p.N;
  unsigned int selmask = 0u;
  if (in_smem && Kb > 0) {
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
      const size_t e = (size_t)tid + (size_t)q * SOLVE_THREADS;
      if (e < n && nkey(kp[e], ip[e]) >= T) selmask |= 1u << q;
    }
  }
  auto visit_selected = [&](auto&& f) {
    if (in_smem) {
      unsigned int m = selmask;
      while (m) {
        unsigned int kk[4];
        int ii[4], jj[4];
        float4 xx[4], yy[4];
        bool on[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          on[u] = m != 0u;
          if (on[u]) {
            const int q = __ffs((int)m) - 1;
            m &= m - 1u;
            const size_t e = (size_t)tid + (size_t)q * SOLVE_THREADS;
            kk[u] = kp[e];
            const unsigned int fi = ip[e];
            ii[u] = (int)(fi / (unsigned int)p.M);
            jj[u] = (int)(fi - (unsigned int)ii[u] * (unsigned int)p.M);
            xx[u] = sp4[ii[u]];
            yy[u] = tp4[jj[u]];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (on[u]) f(kk[u], ii[u], jj[u], xx[u], yy[u]);
      }
    } else {
      for (size_t e = tid; e < n; e += SOLVE_THREADS) {
        const unsigned int k32 = kp[e], fi = ip[e];
        if (nkey(k32, fi) >= T) {
          const int i = (int)(fi / (unsigned int)p.M), j = (int)(fi - (unsigned int)i * (unsigned int)p.M);
          f(k32, i, j, sp4[i], tp4[j]);
        }
      }
    }
  };
  unsigned int ne = 0;
  if (tid == 0) warp_cnt[0] = 0u;
  __syncthreads();
  if (Kb > 0) {
    float m1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // sum w, sum |w|, sum w x (3), sum w y (3)
    visit_selected([&](unsigned int k32, int i, int j, const float4& x4, const float4& y4) {
      const float wf = ordered_to_float(k32);
      if (p.sel_w) {
        const unsigned int pos = atomicAdd(&warp_cnt[0], 1u);
        if (pos < (unsigned int)p.K_max) {
          p.sel_w[(size_t)b * p.K_max + pos] = wf;
          p.sel_src[(size_t)b * p.K_max + pos] = i;
          p.sel_tgt[(size_t)b * p.K_max + pos] = j;
        }
      }
      const float xs[3] = {x4.x, x4.y, x4.z}, ys[3] = {y4.x, y4.y, y4.z};
      m1[0] += wf;
      m1[1] += fabsf(wf);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        m1[2 + a] = fmaf(wf, xs[a], m1[2 + a]);
        m1[5 + a] = fmaf(wf, ys[a], m1[5 + a]);
      }
    });
    double s1[8];
    block_sum_f32<8>(m1, s1, ms);
    PSTAMP(7);
    // w_norm = w / (sum|w| + eps): the normalised weights sum to slightly less than one  (procrustes.py:29-30)
    const double inv = 1.0 / (s1[1] + 1e-4);
    const float invf = (float)inv;
    const float mxf[3] = {(float)(s1[2] * inv), (float)(s1[3] * inv), (float)(s1[4] * inv)};
    const float myf[3] = {(float)(s1[5] * inv), (float)(s1[6] * inv), (float)(s1[7] * inv)};
    float m2[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    visit_selected([&](unsigned int k32, int, int, const float4& x4, const float4& y4) {
      const float wn = ordered_to_float(k32) * invf;
      const float xc[3] = {x4.x - mxf[0], x4.y - mxf[1], x4.z - mxf[2]};
      const float yc[3] = {y4.x - myf[0], y4.y - myf[1], y4.z - myf[2]};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) m2[a * 3 + c] = fmaf(wn * yc[a], xc[c], m2[a * 3 + c]);
    });
    double s2[9];
    block_sum_f32<9>(m2, s2, ms);
    PSTAMP(3);
    if (tid == 0) {
      for (int a = 0; a < 3; ++a) {
        mean_s[a] = (double)mxf[a];
        mean_s[3 + a] = (double)myf[a];
        for (int c = 0; c < 3; ++c) cov_s[a * 3 + c] = s2[a * 3 + c];
      }
    }
    ne = warp_cnt[0];
  } else if (tid == 0) {
    for (int k = 0; k < 6; ++k) mean_s[k] = 0.0;
    for (int k = 0; k < 9; ++k) cov_s[k] = 0.0;
  }
  if (p.sel_w) {
    __syncthreads();
    ne = warp_cnt[0];
    for (int k = (int)ne + tid; k < p.K_max; k += SOLVE_THREADS) {
      p.sel_w[(size_t)b * p.K_max + k] = 0.f;
      p.sel_src[(size_t)b * p.K_max + k] = 0;
      p.sel_tgt[(size_t)b * p.K_max + k] = 0;
    }
  }
  __syncthreads();
  PSTAMP(4);
  if (tid == 0) {
    float R[9], t[3];
    double cond;
    double S[3][3];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) S[a][c] = cov_s[a * 3 + c];
    kabsch_solve(S, mean_s, mean_s + 3, R, t, &cond);
    PSTAMP(5);
    finish_pose(p, b, R, t, cond);
    for (int k = 0; k < 9; ++k) pose_s[k] = p.R_forwd[b * 9 + k];
    for (int k = 0; k < 3; ++k) pose_s[9 + k] = p.t_forwd[b * 3 + k];
  }
  __syncthreads();
  // ---- warp the source points with the gated pose:  (R_forwd s + t_forwd)     pipeline.py:220
  if (p.src_warped) {
    float* o = p.src_warped + (size_t)b * p.N * 3;
    for (int i = tid; i < p.N; i += SOLVE_THREADS) {
      const float x0 = sp[i * 3 + 0], x1 = sp[i * 3 + 1], x2 = sp[i * 3 + 2];
      for (int a = 0; a < 3; ++a) {
        // same association as a 3-term dot product followed by the translation add
        float acc = pose_s[a * 3 + 0] * x0;
        acc = fmaf(pose_s[a * 3 + 1], x1, acc);
        acc = fmaf(pose_s[a * 3 + 2], x2, acc);
        o[i * 3 + a] = acc + pose_s[9 + a];
      }
    }
  }
  PSTAMP(6);
#undef PSTAMP
}

// ---- 5b. the pose step spread over several CTAs (the default) -----------------------------------------------
// The single-CTA kernel above is bound by things one SM does badly: ~19 k shared-memory atomics for the first radix
// level of the select (~10 k cycles) and ~8 k scattered 16-byte point gathers per moment pass, which one SM issues at
// well under one request per clock (~20 k cycles per pass) -- 75-80 k cycles in all.  Here instead:
//   * the collect kernels build the first-level histogram of the candidates as they append them (global atomics spread
//     over the whole GPU);
//   * procr_select_kernel (one CTA per batch element) walks that histogram, reads the candidate keys once, puts the
//     few candidates of the crossing bin on a short list and finds the exact K_b-th largest key T by rank counting;
//   * procr_moments_kernel (up to 64 CTAs per batch element) splits the candidates: every CTA gathers the points of its
//     selected candidates and reduces them to weighted moments about ITS OWN weighted mean (two passes, as the
//     reference centres before it multiplies); the last CTA to arrive combines the partials exactly (parallel-axis
//     terms in fp64), solves the 3x3 problem, applies the condition gate and warps the source points.
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__global__ void __launch_bounds__(SEL_THREADS) procr_select_kernel(const ProcrParams p) {
  __shared__ SelectScratch sc;
  __shared__ unsigned long long list_s[SEL_LIST];
  __shared__ unsigned int list_n;
  __shared__ int bin_s;
  __shared__ unsigned int cum_s, hsel_s;
  __shared__ unsigned long long T_s;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
#define SSTAMP(k) do { if (p.dbg_times && b == 0 && tid == 0) p.dbg_times[(k)] = global_ns(); } while (0)
  SSTAMP(30);
  // the histogram is requested before anything else (it does not depend on the state record read next)
  for (int q = tid; q < TK_BINS; q += SEL_THREADS) sc.hist[q] = __ldcg(p.cand_hist + (size_t)b * TK_BINS + q);
  const size_t total = (size_t)p.N * p.M;
  const ProcrState st = p.state[b];
  const int Kb = st.Kb;
  unsigned int* ckey = p.cand_key + (size_t)b * total;
  unsigned int* cidx = p.cand_idx + (size_t)b * total;
  size_t n = st.n_cand;
  bool slow = false;
  if (n < (size_t)Kb) {
    // fallback: the sample-based bound left too few candidates; take the whole matrix
    for (size_t e = tid; e < total; e += SEL_THREADS) {
      ckey[e] = float_to_ordered(conf_at(p, b, e));
      cidx[e] = (unsigned int)e;
    }
    n = total;
    slow = true;
    __syncthreads();
  }
  unsigned long long T = 0ull;  // Kb >= n: every candidate is used
  if (Kb > 0 && (size_t)Kb < n) {
    if (!slow) {
      if (tid == 0) {
        list_n = 0u;
        T_s = 0ull;
      }
      __syncthreads();
      if (tid < 32) {
        int bin;
        unsigned int cum, hsel;
        warp_walk_hist(sc.hist, (unsigned int)Kb, bin, cum, hsel);
        if (tid == 0) {
          bin_s = bin;
          cum_s = cum;
          hsel_s = hsel;
        }
      }
      __syncthreads();
      SSTAMP(31);
      const int bin = bin_s;
      const unsigned int want = (unsigned int)Kb - cum_s;  // how many of the crossing bin's candidates are selected
      const unsigned int hsel = hsel_s;
      if (hsel > (unsigned int)SEL_LIST || want < 1u || want > hsel) {
        slow = true;  // heavily tied values (or an inconsistent histogram): the general select
      } else {
        // one pass over the candidate keys: those of the crossing bin go to the short list
        // (16-byte loads, up to eight in flight per thread: ~19 k keys arrive in one round trip; ckey is 256-byte aligned)
        // (a batch element's list starts at b * N * M keys: 16-byte aligned unless N * M is odd-ish -- then no vector part)
        const bool vec4 = (((uintptr_t)ckey) & 15u) == 0;
        const size_t n4 = vec4 ? ((n + 3) >> 2) : 0;   // reading up to 3 keys past n stays inside the (padded) workspace
        const uint4* k4 = reinterpret_cast<const uint4*>(ckey);
        for (size_t q0 = 0; q0 < n4; q0 += (size_t)SEL_THREADS * 8) {
          uint4 kk[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const size_t q = q0 + (size_t)u * SEL_THREADS + tid;
            kk[u] = q < n4 ? k4[q] : make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const size_t q = q0 + (size_t)u * SEL_THREADS + tid;
            const unsigned int k1[4] = {kk[u].x, kk[u].y, kk[u].z, kk[u].w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const size_t e = (q << 2) + w;
              if (q < n4 && e < n && (int)cand_bin(k1[w], st.hist_kmin, st.hist_sh) == bin) {
                const unsigned int pos = atomicAdd(&list_n, 1u);
                if (pos < (unsigned int)SEL_LIST) list_s[pos] = make_key64(k1[w], cidx[e]);  // (fetching all indices with the keys: slower)
              }
            }
          }
        }
        for (size_t e = (n4 << 2) + tid; e < n; e += SEL_THREADS) {  // unaligned list: scalar loads
          const unsigned int k32 = ckey[e];
          if ((int)cand_bin(k32, st.hist_kmin, st.hist_sh) == bin) {
            const unsigned int pos = atomicAdd(&list_n, 1u);
            if (pos < (unsigned int)SEL_LIST) list_s[pos] = make_key64(k32, cidx[e]);
          }
        }
        __syncthreads();
        SSTAMP(32);
        const unsigned int L = list_n;
        if (L != hsel) {
          slow = true;  // cannot happen unless the histogram and the list disagree; stay exact
        } else {
          if ((unsigned int)tid < L) {
            const unsigned long long my = list_s[tid];
            unsigned int rank = 0u;
            for (unsigned int q = 0; q < L; ++q) rank += (list_s[q] > my) ? 1u : 0u;  // keys are distinct (distinct indices)
            if (rank == want - 1u) T_s = my;
          }
          __syncthreads();
          T = T_s;
        }
      }
    }
    if (slow) T = block_select_kth<SEL_THREADS>([&](size_t e) { return make_key64(ckey[e], cidx[e]); }, n, Kb, sc);
  }
  SSTAMP(33);
#undef SSTAMP
  if (tid == 0) {
    p.state[b].T = T;
    p.state[b].n_cand = (unsigned int)n;
    p.state[b].pad_ = (slow ? 0x80000000u : 0u) | (unsigned int)(Kb > 0 && (size_t)Kb < n ? hsel_s & 0x7FFFFFFFu : 0u);  // diagnostics
  }
}

__global__ void __launch_bounds__(PM_THREADS) procr_moments_kernel(const ProcrParams p) {
  __shared__ MomentScratch ms;
  __shared__ unsigned int ticket_s;
  __shared__ double comb_s[16];
  __shared__ double mean_s[6], cov_s[9];
  __shared__ float pose_s[12];
  const int b = blockIdx.y;
  const int G = gridDim.x;
  const int tid = threadIdx.x;
#define MSTAMP0(k) do { if (p.dbg_times && b == 0 && blockIdx.x == 0 && tid == 0) p.dbg_times[(k)] = global_ns(); } while (0)
#define MSTAMPL(k) do { if (p.dbg_times && b == 0 && tid == 0) p.dbg_times[(k)] = global_ns(); } while (0)
  MSTAMP0(40);
  const size_t total = (size_t)p.N * p.M;
  const ProcrState st = p.state[b];
  const int Kb = st.Kb;
  const unsigned long long T = st.T;
  const size_t n = st.n_cand;
  const unsigned int* ckey = p.cand_key + (size_t)b * total;
  const unsigned int* cidx = p.cand_idx + (size_t)b * total;
  const float4* sp4 = p.pcd4 + (size_t)b * (p.N + p.M);  // written by the collect kernel
  const float4* tp4 = sp4 + p.N;
  // this CTA's slice of the candidate list
  const size_t per = (n + G - 1) / G;
  const size_t e_lo = min(n, per * blockIdx.x), e_hi = min(n, e_lo + per);
  // visit the selected candidates of the slice, two per thread at a time (four gathers in flight)
  auto visit_selected = [&](auto&& f) {
    if (Kb <= 0) return;
    for (size_t e0 = e_lo; e0 < e_hi; e0 += (size_t)PM_THREADS * 2) {
      unsigned int kk[2], fi[2];
      int ii[2], jj[2];
      float4 xx[2], yy[2];
      bool on[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const size_t e = e0 + (size_t)u * PM_THREADS + tid;
        on[u] = e < e_hi;
        kk[u] = on[u] ? ckey[e] : 0u;
        fi[u] = on[u] ? cidx[e] : 0u;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        on[u] = on[u] && make_key64(kk[u], fi[u]) >= T;
        if (on[u]) {
          ii[u] = (int)(fi[u] / (unsigned int)p.M);
          jj[u] = (int)(fi[u] - (unsigned int)ii[u] * (unsigned int)p.M);
          xx[u] = sp4[ii[u]];
          yy[u] = tp4[jj[u]];
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (on[u]) f(kk[u], ii[u], jj[u], xx[u], yy[u]);
    }
  };
  // pass A: sum w, sum |w|, sum w x, sum w y  ->  this CTA's centre (its weighted mean)
  float m1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  visit_selected([&](unsigned int k32, int i, int j, const float4& x4, const float4& y4) {
    const float wf = ordered_to_float(k32);
    if (p.sel_w) {
      const unsigned int pos = atomicAdd(&p.state[b].sel_count, 1u);
      if (pos < (unsigned int)p.K_max) {
        p.sel_w[(size_t)b * p.K_max + pos] = wf;
        p.sel_src[(size_t)b * p.K_max + pos] = i;
        p.sel_tgt[(size_t)b * p.K_max + pos] = j;
      }
    }
    const float xs[3] = {x4.x, x4.y, x4.z}, ys[3] = {y4.x, y4.y, y4.z};
    m1[0] += wf;
    m1[1] += fabsf(wf);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      m1[2 + a] = fmaf(wf, xs[a], m1[2 + a]);
      m1[5 + a] = fmaf(wf, ys[a], m1[5 + a]);
    }
  });
  double s1[8];
  block_sum_f32<8>(m1, s1, ms);
  MSTAMP0(41);
  float cxf[3] = {0.f, 0.f, 0.f}, cyf[3] = {0.f, 0.f, 0.f};
  if (s1[0] != 0.0 && s1[1] > 0.0) {
    const double iw = 1.0 / s1[0];
    for (int a = 0; a < 3; ++a) {
      cxf[a] = (float)(s1[2 + a] * iw);
      cyf[a] = (float)(s1[5 + a] * iw);
      if (!(fabsf(cxf[a]) < INFINITY)) cxf[a] = 0.f;  // any finite centre is valid
      if (!(fabsf(cyf[a]) < INFINITY)) cyf[a] = 0.f;
    }
  }
  // pass B: moments about the centre: D = sum w (x - cx), E = sum w (y - cy), C = sum w (y - cy)(x - cx)^T
  float m2[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) m2[k] = 0.f;
  visit_selected([&](unsigned int k32, int, int, const float4& x4, const float4& y4) {
    const float wf = ordered_to_float(k32);
    const float xc[3] = {x4.x - cxf[0], x4.y - cxf[1], x4.z - cxf[2]};
    const float yc[3] = {y4.x - cyf[0], y4.y - cyf[1], y4.z - cyf[2]};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      m2[a] = fmaf(wf, xc[a], m2[a]);
      m2[3 + a] = fmaf(wf, yc[a], m2[3 + a]);
      const float wy = wf * yc[a];
#pragma unroll
      for (int c = 0; c < 3; ++c) m2[6 + a * 3 + c] = fmaf(wy, xc[c], m2[6 + a * 3 + c]);
    }
  });
  double s2[15];
  block_sum_f32<15>(m2, s2, ms);
  MSTAMP0(42);
  double* part = p.partials + ((size_t)b * PM_MAX_G + blockIdx.x) * PM_PART;
  if (tid == 0) {
    part[0] = s1[0];
    part[1] = s1[1];
    for (int a = 0; a < 3; ++a) {
      part[2 + a] = (double)cxf[a];
      part[5 + a] = (double)cyf[a];
    }
    for (int k = 0; k < 15; ++k) part[8 + k] = s2[k];
    __threadfence();
    ticket_s = atomicAdd(&p.moments_arrive[b], 1u);
  }
  __syncthreads();
  MSTAMP0(43);
  if (ticket_s != (unsigned int)(G - 1)) return;  // not the last CTA of this batch element
  MSTAMPL(44);
  if (tid == 0) p.moments_arrive[b] = 0u;         // self-reset for the next call
  __threadfence();
  // ---- combine (fixed order over the CTAs: reproducible).  With centres c_g:
  //   sum w x = sum_g (D_g + W_g cx_g),   S = inv * sum_g [ C_g + E_g (cx_g - mx)^T + (cy_g - my) D_g^T + W_g (cy_g - my)(cx_g - mx)^T ]
  // all partials into shared memory with one round trip (a thread walking them in global memory pays an L2 latency
  // per CTA: 64 x ~700 cycles)
  __shared__ double part_s[PM_MAX_G * PM_PART];
  {
    const double* pg = p.partials + (size_t)b * PM_MAX_G * PM_PART;
    for (int q = tid; q < G * PM_PART; q += PM_THREADS) part_s[q] = __ldcg(pg + q);
  }
  __syncthreads();
  const double* pb = part_s;
  // one warp per output value: lane l takes CTAs l and l + 32, then a fixed shuffle tree adds the lanes (a single thread
  // walking the 64 partials in fp64 cost ~3 us per phase: a dependent chain of ~100 cycles per CTA)
  const int lane = tid & 31, warp = tid >> 5;
  auto warp_sum_f64 = [&](double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
  };
  for (int out = warp; out < 7; out += PM_THREADS / 32) {
    double acc = 0.0;
    for (int g = lane; g < G; g += 32) {
      const double* q = pb + (size_t)g * PM_PART;
      if (out == 0) acc += q[1];
      else if (out < 4) acc += q[8 + (out - 1)] + q[0] * q[2 + (out - 1)];
      else acc += q[11 + (out - 4)] + q[0] * q[5 + (out - 4)];
    }
    acc = warp_sum_f64(acc);
    if (lane == 0) comb_s[out] = acc;
  }
  __syncthreads();
  // w_norm = w / (sum|w| + eps): the normalised weights sum to slightly less than one  (procrustes.py:29-30)
  const double inv = 1.0 / (comb_s[0] + 1e-4);
  const float invf = (float)inv;
  double mx[3], my[3];
  for (int a = 0; a < 3; ++a) {
    mx[a] = (double)(float)(comb_s[1 + a] * inv);  // the means are fp32 values, as in the reference
    my[a] = (double)(float)(comb_s[4 + a] * inv);
  }
  for (int out = warp; out < 9; out += PM_THREADS / 32) {
    const int a = out / 3, c = out - 3 * a;
    double acc = 0.0;
    for (int g = lane; g < G; g += 32) {
      const double* q = pb + (size_t)g * PM_PART;
      const double W = q[0], dx = q[2 + c] - mx[c], dy = q[5 + a] - my[a];
      acc += q[14 + a * 3 + c] + q[11 + a] * dx + dy * q[8 + c] + W * dy * dx;
    }
    acc = warp_sum_f64(acc);
    if (lane == 0) cov_s[out] = (Kb > 0) ? acc * (double)invf : 0.0;
  }
  if (tid == 0) {
    for (int a = 0; a < 3; ++a) {
      mean_s[a] = (Kb > 0) ? mx[a] : 0.0;
      mean_s[3 + a] = (Kb > 0) ? my[a] : 0.0;
    }
  }
  __syncthreads();
  MSTAMPL(45);
  if (tid == 0) {
    float R[9], t[3];
    double cond;
    double S[3][3];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) S[a][c] = cov_s[a * 3 + c];
    kabsch_solve(S, mean_s, mean_s + 3, R, t, &cond);
    MSTAMPL(46);
    finish_pose(p, b, R, t, cond);
    for (int k = 0; k < 9; ++k) pose_s[k] = p.R_forwd[b * 9 + k];
    for (int k = 0; k < 3; ++k) pose_s[9 + k] = p.t_forwd[b * 3 + k];
  }
  if (p.sel_w) {
    const unsigned int ne = min(__ldcg(&p.state[b].sel_count), (unsigned int)p.K_max);
    for (int k = (int)ne + tid; k < p.K_max; k += PM_THREADS) {
      p.sel_w[(size_t)b * p.K_max + k] = 0.f;
      p.sel_src[(size_t)b * p.K_max + k] = 0;
      p.sel_tgt[(size_t)b * p.K_max + k] = 0;
    }
  }
  __syncthreads();
  // ---- warp the source points with the gated pose:  (R_forwd s + t_forwd)     pipeline.py:220
  if (p.src_warped) {
    const float* sp = p.src_pcd + (size_t)b * p.N * 3;
    float* o = p.src_warped + (size_t)b * p.N * 3;
    for (int i0 = 0; i0 < p.N; i0 += PM_THREADS * 8) {  // eight points per thread and batch: 24 loads in flight
      float xin[8][3];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * PM_THREADS + tid;
#pragma unroll
        for (int a = 0; a < 3; ++a) xin[u][a] = i < p.N ? sp[i * 3 + a] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * PM_THREADS + tid;
        if (i < p.N) {
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            // same association as a 3-term dot product followed by the translation add
            float acc = pose_s[a * 3 + 0] * xin[u][0];
            acc = fmaf(pose_s[a * 3 + 1], xin[u][1], acc);
            acc = fmaf(pose_s[a * 3 + 2], xin[u][2], acc);
            o[i * 3 + a] = acc + pose_s[9 + a];
          }
        }
      }
    }
  }
  MSTAMPL(47);
#undef MSTAMP0
#undef MSTAMPL
}

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

static long long* g_procr_times = nullptr;  // tuning only

struct ProcrWorkspace {
  ProcrState* state;
  unsigned int* cand_key;
  unsigned int* cand_idx;
  unsigned int* sample_buf;
  unsigned int* sample_arrive;
  unsigned int* moments_arrive;
  unsigned int* cand_hist;
  double* partials;
  float4* pcd4;
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
  w.sample_arrive = (unsigned int*)take(4ull * 2 * B);  // sampling CTAs, then moments CTAs
  w.moments_arrive = w.sample_arrive ? w.sample_arrive + B : nullptr;
  w.cand_hist = (unsigned int*)take(4ull * B * TK_BINS);
  w.partials = (double*)take(8ull * B * PM_MAX_G * PM_PART);
  w.pcd4 = (float4*)take(16ull * B * ((size_t)N + M));
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

static int procr_run(const drg_procrustes_args* a, const ProcrSource* src, void* workspace, size_t workspace_bytes, void* stream) {
  DRG_CHECK_ARG(a != nullptr, "args is null");
  DRG_CHECK_ARG((a->conf != nullptr || src != nullptr) && a->src_pcd && a->tgt_pcd, "conf/src_pcd/tgt_pcd must be non-null");
  DRG_CHECK_ARG(a->padded_lengths || (a->src_mask && a->tgt_mask), "masks must be non-null unless padded_lengths is set");
  DRG_CHECK_ARG(a->B >= 1 && a->B <= 1024 && a->N >= 1 && a->M >= 1, "need 1 <= B <= 1024 and N, M >= 1");
  DRG_CHECK_ARG((long long)a->N * a->M < (1ll << 32), "N*M must fit in 32 bits");
  DRG_CHECK_ARG(a->R && a->t && a->R_forwd && a->t_forwd && a->condition && a->solution_mask, "pose outputs must be non-null");
  DRG_CHECK_ARG(a->sample_rate > 0.f, "sample_rate must be > 0");
  DRG_CHECK_ARG((a->sel_w == nullptr) == (a->sel_src == nullptr) && (a->sel_w == nullptr) == (a->sel_tgt == nullptr),
                "sel_w/sel_src/sel_tgt must be given together");
  const int B = a->B, N = a->N, M = a->M;
  ProcrWorkspace w = procr_carve(workspace, B, N, M);
  if (workspace == nullptr || workspace_bytes < w.total || ((uintptr_t)workspace & 255u)) {
    set_error("soft_procrustes: workspace missing, too small (%zu < %zu) or not 256-byte aligned", workspace_bytes, w.total);
    return DRG_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ProcrParams p{};
  p.conf = src ? nullptr : a->conf;
  if (src) {
    p.scores = src->scores;
    p.pu = src->views.u;
    p.pv = src->views.v;
    p.pbc = src->views.bc;
    p.pshift = src->shift;
    p.ldu = src->views.ldu;
    p.ldv = src->views.ldv;
    p.apply_mask = src->apply_mask;
  }
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
  p.moments_arrive = w.moments_arrive;
  p.cand_hist = w.cand_hist;
  p.partials = w.partials;
  p.pcd4 = w.pcd4;
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
  {
    static long long* tbuf = nullptr;
    static int want = -1;
    if (want < 0) want = getenv("DRG_PROCR_TIMES") ? 1 : 0;
    if (want && !tbuf) {
      cudaMalloc(&tbuf, 64 * sizeof(long long));
      cudaMemset(tbuf, 0, 64 * sizeof(long long));
      g_procr_times = tbuf;
    }
    p.dbg_times = want ? tbuf : nullptr;
  }

  static bool attr_set = false;
  if (!attr_set) {
    DRG_CUDA(cudaFuncSetAttribute(topk_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SAMPLES * 4));
    attr_set = true;
  }
  // the arrival counters must be zero on entry (the workspace is caller memory of unknown content)
  DRG_CUDA(cudaMemsetAsync(w.sample_arrive, 0, 4ull * 2 * B, st));
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
    static int per_sm = -1;  // tuning only: DRG_COLLECT_PER_SM
    if (per_sm < 0) {
      const char* e = getenv("DRG_COLLECT_PER_SM");
      per_sm = e ? atoi(e) : 2;
      if (per_sm < 1 || per_sm > 16) per_sm = 2;
    }
    // CTAs per SM.  80 registers x 256 threads allow three resident CTAs per SM; measured at 4096^2 inside the step
    // (tools/collect_sweep.sh): 1 -> 43.7 us, 2 -> 28.9 us, 3 -> 35.1 us, 4 -> 30.6 us, 5 -> 34.1 us (2960 / 2991 / 3005
    // steps/s for 5 / 4 / 2): one resident wave of 296 CTAs with ~14 rows each beats more, shorter CTAs.
    int gr = (NUM_SMS * per_sm) / B;
    if (gr < 1) gr = 1;
    if (gr > N) gr = N;
    topk_collect_rows_kernel<<<dim3(gr, B), 256, 0, st>>>(p);
  } else {
    topk_collect_kernel<<<dim3(gx, B), 256, 0, st>>>(p);
  }
  }
  DRG_LAUNCH_CHECK();
  static int single_cta = -1;  // DRG_PROCR_SINGLE=1: the single-CTA solve kernel (A/B comparisons)
  if (single_cta < 0) single_cta = getenv("DRG_PROCR_SINGLE") ? 1 : 0;
  if (single_cta) {
    static bool solve_attr_set = false;
    if (!solve_attr_set) {
      DRG_CUDA(cudaFuncSetAttribute(procr_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SOLVE_SMEM_CAND * 8));
      solve_attr_set = true;
    }
    ProfScope prof_scope(PROF_PROCR_SOLVE, st);
    procr_solve_kernel<<<B, SOLVE_THREADS, SOLVE_SMEM_CAND * 8, st>>>(p);
  } else {
    int G = NUM_SMS / B;
    if (G > PM_MAX_G) G = PM_MAX_G;
    if (G < 4) G = 4;
    {
      ProfScope prof_scope(PROF_PROCR_SELECT, st);
      procr_select_kernel<<<B, SEL_THREADS, 0, st>>>(p);
    }
    DRG_LAUNCH_CHECK();
    ProfScope prof_scope(PROF_PROCR_SOLVE, st);
    procr_moments_kernel<<<dim3(G, B), PM_THREADS, 0, st>>>(p);
  }
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_soft_procrustes(const drg_procrustes_args* a, void* workspace, size_t workspace_bytes, void* stream) {
  DRG_CHECK_ARG(a != nullptr && a->conf != nullptr, "args / conf is null");
  return procr_run(a, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int drg_sinkhorn_soft_procrustes(const drg_sinkhorn_args* s, const drg_procrustes_args* a, void* skh_workspace,
                                            size_t skh_workspace_bytes, void* procr_workspace, size_t procr_workspace_bytes,
                                            void* stream) {
  DRG_CHECK_ARG(s != nullptr && a != nullptr, "args are null");
  DRG_CHECK_ARG(s->out_mode == DRG_OUT_NONE, "the fused call takes out_mode DRG_OUT_NONE: the confidence matrix is never written");
  DRG_CHECK_ARG(s->B == a->B && s->N == a->N && s->M == a->M, "sinkhorn and procrustes shapes differ");
  DRG_CHECK_ARG(a->src_mask == s->src_mask && a->tgt_mask == s->tgt_mask, "the fused call uses one pair of masks");
  ProcrSource src{};
  int rc = skh_run_with_views(s, skh_workspace, skh_workspace_bytes, stream, &src.views);
  if (rc != DRG_OK) return rc;
  src.scores = s->scores;
  src.shift = s->shift;
  src.apply_mask = s->apply_mask;
  return procr_run(a, &src, procr_workspace, procr_workspace_bytes, stream);
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

extern "C" int drg_debug_read_procr_times(long long* host_out, int n) {
  if (!g_procr_times || n > 64) return DRG_ERR_UNSUPPORTED;
  DRG_CUDA(cudaMemcpy(host_out, g_procr_times, sizeof(long long) * n, cudaMemcpyDeviceToHost));
  return DRG_OK;
}
