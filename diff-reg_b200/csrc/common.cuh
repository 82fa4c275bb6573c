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
// runs drg_sinkhorn (out_mode NONE allowed) and reports the views (sinkhorn.cu)
int skh_run_with_views(const drg_sinkhorn_args* a, void* workspace, size_t workspace_bytes, void* stream, SkhViews* views);

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
  PROF_NSLOTS
};
extern std::atomic<int> g_prof_enabled;
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
#endif  // __CUDACC__

}  // namespace drg
