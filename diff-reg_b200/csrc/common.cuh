// Shared device/host helpers for the diffreg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/diffreg_b200.h"

namespace drg {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float NEG_BIG = -1.0e30f;  // finite stand-in for -inf in running maxima (avoids inf-inf)
constexpr int NUM_SMS = 148;
// 16-bit split GEMM operands (features.cu / gemm.cu): a row holds two segments of kc = K rounded up to 64 columns (hi and lo
// halves of the row-scaled values) and an 8-column tail with the row's (1 / scale, Euclidean norm) as floats
__host__ __device__ inline int split16_kc(int K) { return (K + 63) & ~63; }
__host__ __device__ inline int split16_pitch(int K) { return 2 * split16_kc(K) + 8; }

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define DRG_CHECK_ARG(cond, msg)                  \
  do {                                            \
    if (!(cond)) {                                \
      ::drg::set_error("invalid argument: %s", msg); \
      return DRG_ERR_INVALID;                     \
    }                                             \
  } while (0)

#define DRG_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::drg::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DRG_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define DRG_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      ::drg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DRG_ERR_CUDA;                                                              \
    }                                                                                   \
    ::drg::count_launch();                                                              \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- shared between sinkhorn.cu and procrustes.cu ------------------------------------------
struct SkhConst {  // per batch element normalisation constants of the Sinkhorn (matching.py:24-27)
  float norm;        // -log(ms + ns)
  float log_mu_bin;  // log(ns) + norm
  float log_nu_bin;  // log(ms) + norm
  float pad;
};
// where a finished Sinkhorn left its potentials inside its workspace
struct SkhViews {
  const float* u;       // [B, ldu]
  const float* v;       // [B, ldv]
  const SkhConst* bc;   // [B]
  int ldu, ldv;
};

// ---- top-K candidate search shared by the Sinkhorn tail (sinkhorn.cu) and the SoftProcrustes kernels (procrustes.cu) ----
constexpr int TK_BINS = 2048;
struct ProcrState {  // per batch element
  int Kb;                         // number of correspondences to use
  unsigned int n_cand;            // candidates appended so far
  unsigned long long lower_key;   // stand-alone path: candidates have 64-bit key >= lower_key
  unsigned long long T;           // written by the pose kernel: the Kb best candidates are those with key >= T
  unsigned int hist_kmin;         // candidate histogram (filled by the collect pass): bin = ((value key - kmin) << sh) >> 21,
  int hist_sh;                    //   clamped to [0, TK_BINS) -- monotone in the key, ~2048 bins over [bound, 2 x sample range]
  unsigned int sel_count;         // selected correspondences written to sel_* so far
  unsigned int pad_;
  int seg_G;                      // > 0: the list was written by seg_G producer CTAs, producer g owning the rows
                                  //   [N g / seg_G, N (g + 1) / seg_G) and ONE contiguous segment of the list (cand_seg)
  unsigned int seg_broken;        // != 0: some producer also appended outside its segment (its shared list was full)
};
// What the persistent Sinkhorn needs to run the candidate search of SoftProcrustes as its own last phase (the pose step
// of get_warped_from_noising_matching, Diff-Reg-4dmatch/models/pipeline.py:207-223): the score slab is still L2-resident
// and the potentials are final, so neither a sampling kernel nor a separate pass over the matrix is needed.
struct SkhCollect {
  ProcrState* state;              // [B]
  unsigned int* cand_key;         // [B, N*M] order-preserving bits of the candidates' confidences
  unsigned int* cand_idx;         // [B, N*M] their flat indices
  unsigned int* cand_hist;        // [B, TK_BINS]
  unsigned int* sample_hist;      // [B, TK_BINS] histogram of the sampled log2 confidences (SH_PER_OCTAVE bins per octave)
  uint2* cand_seg;                // [B, NUM_SMS] (offset, count) of every producer CTA's segment of the candidate list
  unsigned int* sample_hist2;     // [B, TK_BINS] second-level histogram: the samples of a crowded first-level bin, re-binned
  float* sample_val;              // [B, SH_ROWS, ldv] the sampled log2 confidences themselves (each CTA re-reads its own)
  unsigned long long* sample_list; // [B, SH_ROWS * ldv] third level: 64-bit keys of the samples of a crowded sub-bin
  unsigned int* sample_list_n;    // [B] their number
  float sample_rate;              // SoftProcrustesLayer.sample_rate
  int padded_lengths;             // 3DMatch variant: lengths are N, M whatever the masks say
  int K_max;                      // K_b is clamped to this
};
// runs drg_sinkhorn (out_mode NONE allowed) and reports the views (sinkhorn.cu).  collect != NULL: if the persistent
// kernel handles the shape it also leaves the top-K candidate list / histogram / state for procr_pose_kernel and sets
// *collected; otherwise *collected = false and the caller runs the stand-alone threshold + collect kernels.
int skh_run_with_views(const drg_sinkhorn_args* a, void* workspace, size_t workspace_bytes, void* stream, SkhViews* views,
                       const SkhCollect* collect = nullptr, bool* collected = nullptr);

// ---- peer-to-peer exchange over NVLink (row-sharded Sinkhorn, p2p.cu / sinkhorn.cu) ---------------------
constexpr int P2P_MAX_RANKS = 8;
// One allocation per rank, exported to the peers with CUDA IPC:
//   inbox [2 slots][world senders][slot_elems] float2, flags [2 slots][nflags] u32, status (1 = a wait timed out)
struct P2PComm {
  int rank, world;
  size_t slot_elems;
  int nflags;
  void* base;                     // this rank's allocation
  void* peer_base[P2P_MAX_RANKS]; // every rank's allocation as mapped into this process (own: base)
  bool peer_open[P2P_MAX_RANKS];
  unsigned int epoch;             // exchanges performed so far (identical on every rank)
  size_t bytes;
};
struct P2PView {  // what an exchange kernel needs, by value
  float2* inbox[P2P_MAX_RANKS];
  unsigned int* flags[P2P_MAX_RANKS];
  int* status;
  size_t slot_elems;
  int world, rank, slot, nflags;
  unsigned int target;            // arrivals expected on a flag of this slot once every rank has sent
};
inline size_t p2p_inbox_bytes(const P2PComm& c) { return 2 * (size_t)c.world * c.slot_elems * sizeof(float2); }
inline P2PView p2p_view(P2PComm& c) {
  P2PView v{};
  for (int r = 0; r < c.world; ++r) {
    v.inbox[r] = reinterpret_cast<float2*>(c.peer_base[r]);
    v.flags[r] = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(c.peer_base[r]) + p2p_inbox_bytes(c));
  }
  v.status = reinterpret_cast<int*>(reinterpret_cast<char*>(c.base) + p2p_inbox_bytes(c) + 2 * (size_t)c.nflags * sizeof(unsigned int));
  v.slot_elems = c.slot_elems;
  v.world = c.world;
  v.rank = c.rank;
  v.slot = (int)(c.epoch & 1u);
  v.nflags = c.nflags;
  v.target = (unsigned int)c.world * (c.epoch / 2u + 1u);
  return v;
}

// ---- optional per-kernel timing (bench.py's roofline leg) ---------------------------------
// When enabled through drg_profile_enable(1), every launch of a slotted kernel is bracketed by two
// CUDA events on the launching stream; drg_profile_read() sums the elapsed times.  Disabled: no cost.
enum ProfSlot {
  PROF_SKH_ITER = 0,
  PROF_SKH_COL,
  PROF_SKH_FINAL,
  PROF_SKH_PREP,
  PROF_GEMM,
  PROF_PREP_OPERAND,
  PROF_ROWCOL_BEST,
  PROF_MATCH_ROWS,
  PROF_TOPK_COLLECT,
  PROF_PROCR_SOLVE,
  PROF_TOPK_THRESHOLD,
  PROF_PROCR_SELECT,
  PROF_SKH_FUSED,   // persistent Sinkhorn launches that carry the final pass (exp / DDIM) as their last phase
  PROF_NSLOTS
};
extern std::atomic<int> g_prof_enabled;
// tuning only: caller-owned device buffer (>= 1024 int64) for in-kernel phase stamps, NULL = off (drg_tuning_set_stamp_buffer)
extern long long* g_tuning_stamps;
void prof_begin(int slot, cudaStream_t st);
void prof_end(int slot, cudaStream_t st);
struct ProfScope {
  int slot;
  cudaStream_t st;
  bool on;
  ProfScope(int s, cudaStream_t stream) : slot(s), st(stream), on(g_prof_enabled.load(std::memory_order_relaxed) != 0) {
    if (on) prof_begin(slot, st);
  }
  ~ProfScope() {
    if (on) prof_end(slot, st);
  }
};

// ---- device helpers -------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mbarrier (shared::cta)
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// try_wait with a suspend-time hint: the hardware parks the thread for up to `ns` instead of re-issuing the probe
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns) {
  // probe; if the phase is not complete, back off with nanosleep so that waiting warps do not take issue slots
  // from the warps that are computing on the same scheduler
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 4-byte cp.async (LDGSTS) for rows that are not 16-byte aligned
__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// arrive on the mbarrier once all prior cp.async of this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ float warp_min(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fminf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// (max, sum) pair of a log2-domain log-sum-exp: value = m + log2(s)
struct LseAcc {
  float m, s;
};
__device__ __forceinline__ LseAcc lse_empty() { return LseAcc{NEG_BIG, 0.f}; }
__device__ __forceinline__ void lse_add_value(LseAcc& a, float x2) {  // exact online add of one value
  float mn = fmaxf(a.m, x2);
  a.s = a.s * ex2(a.m - mn) + ex2(x2 - mn);
  a.m = mn;
}
__device__ __forceinline__ void lse_merge(LseAcc& a, float m2, float s2) {
  float mn = fmaxf(a.m, m2);
  a.s = a.s * ex2(a.m - mn) + s2 * ex2(m2 - mn);
  a.m = mn;
}
__device__ __forceinline__ float lse_value(const LseAcc& a) { return a.m + lg2(a.s); }

// order-preserving float <-> uint encoding for atomicMin/Max on floats
__device__ __forceinline__ unsigned int float_to_ordered(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bin of a candidate's value key in the candidate histogram (monotone in the key)
__device__ __forceinline__ unsigned int cand_bin(unsigned int k32, unsigned int kmin, int sh) {
  if (k32 < kmin) return 0u;
  const unsigned long long d = ((unsigned long long)(k32 - kmin) << sh) >> 21;
  return d > (unsigned long long)(TK_BINS - 1) ? (unsigned int)(TK_BINS - 1) : (unsigned int)d;
}
// range of the candidate histogram: from the bound (key kmin) to twice the distance of the largest sample (key smax);
// larger keys share the top bin
__device__ __forceinline__ int cand_hist_shift(unsigned int kmin, unsigned int smax, bool have_bound) {
  unsigned int range = 0xFFFFFFFFu - kmin;
  if (have_bound && smax > kmin) {
    const unsigned long long r2 = 2ull * (unsigned long long)(smax - kmin) + 1ull;
    if (r2 < (unsigned long long)range) range = (unsigned int)r2;
  }
  return range ? __clz((int)range) : 32;
}
__device__ __forceinline__ unsigned int hash_u32(unsigned int x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

// One warp walks a TK_BINS-bin histogram (shared memory) from the top and finds the bin where the running count reaches
// krem: `bin`, the count `cum` in the bins above it and its own count `hsel`, returned to all lanes.  Two steps, nothing
// kept in registers and no serial scan.  Step 1: lane l sums the l-th chunk of 64 bins (descending); a warp scan finds
// the chunk of the crossing.  Step 2: the 32 lanes split that chunk two bins each and scan again.  If krem exceeds the
// total count the lowest bin is returned.
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

__device__ __forceinline__ unsigned long long make_key64(unsigned int ordered_value, unsigned int flat_index) {
  return ((unsigned long long)ordered_value << 32) | (unsigned long long)(0xFFFFFFFFu - flat_index);
}

// Exact selection of the k-th largest of n distinct 64-bit keys by one CTA (1 <= k <= n).
// key_at(e) returns the key of element e.  Six radix levels (11,11,10,11,11,10 bits, MSB first); stops
// early once the remaining bucket is wanted whole.  Returns T such that exactly k keys are >= T.
// (Measured alternatives that were slower on B200: warp-aggregated histogram updates via match.any, and
// normalising the keys to their common range first.)
struct SelectCtl {
  unsigned long long prefix;
  int krem;
  int done;
};
struct __align__(16) SelectScratch {
  unsigned int hist[TK_BINS];
  SelectCtl ctl;
};

struct SyncAll {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
template <int NT, class KeyAt, class Sync = SyncAll>
__device__ unsigned long long block_select_kth(KeyAt key_at, size_t n, int k, unsigned int* hist, SelectCtl& ctl, Sync sync = Sync()) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    ctl.prefix = 0ull;
    ctl.krem = k;
    ctl.done = 0;
  }
  sync();
  int shift = 64;
  const int widths[6] = {11, 11, 10, 11, 11, 10};
  for (int level = 0; level < 6; ++level) {
    const int wbits = widths[level];
    shift -= wbits;
    for (int q = tid; q < TK_BINS; q += NT) hist[q] = 0u;
    sync();
    const unsigned long long prefix = ctl.prefix;
    const int hi_shift = shift + wbits;  // bits above the current digit
    for (size_t e = tid; e < n; e += NT) {
      const unsigned long long key = key_at(e);
      const bool match = (hi_shift >= 64) ? true : ((key >> hi_shift) == prefix);
      if (match) atomicAdd(&hist[(unsigned int)((key >> shift) & ((1u << wbits) - 1u))], 1u);
    }
    sync();
    if (tid < 32) {
      int dbin;
      unsigned int cum, hsel;
      const unsigned int krem = (unsigned int)ctl.krem;
      warp_walk_hist(hist, krem, dbin, cum, hsel);
      if (tid == 0) {
        ctl.prefix = (prefix << wbits) | (unsigned long long)dbin;
        ctl.krem = (int)(krem - cum);
        if (hsel == krem - cum) ctl.done = 1;  // the whole bucket is wanted
      }
    }
    sync();
    if (ctl.done) break;
  }
  const unsigned long long T = ctl.prefix << shift;
  sync();
  return T;
}

#endif  // __CUDACC__

}  // namespace drg
