// Correspondence extraction: mutual-nearest-neighbour / confidence-threshold matches.  sm_100a.
//
// Replaces
//   Matching.get_match / get_topk_match       Diff-Reg-4dmatch/models/matching.py:71-107
//   mutual_topk_select (k = 1)                Diff-Reg-2d3d/vision3d/ops/mutual_topk_select.py:7-60
//                                             (= Diff-Reg-3dmatch/models/matching.py:6-59)
//   mutual_topk_select / batch_mutual_topk_select for k > 1 (the 2D-3D fine matching, model.py:738-746: k = 2 on
//   [B, Kc, Kc] patch similarities)           Diff-Reg-2d3d/vision3d/ops/mutual_topk_select.py:7-133
//     -> topk_rows_kernel / topk_cols_kernel leave every row's / column's k-th best packed key; an entry is in the top-k
//        of its row iff its key >= that key (mode 2 of the count / write passes)
//
// Pass 1 (rowcol_best_kernel): one read of the matrix gives, for every row and every column, the
//   best value and the lowest index attaining it, as a packed 64-bit key merged with atomicMax:
//   key = ordered(value) << 32 | ~index.  Each thread owns a column quad and walks down a band
//   of rows: column bests stay in registers, row bests are a warp-shuffle arg-max.
// Pass 2 (count) / scan / pass 3 (write): hits per row -> exclusive scan -> ordered write, so
//   the output is in the row-major order torch.nonzero() produces, with no host round-trip
//   except the caller reading the total.
#include <algorithm>

#include "common.cuh"

namespace drg {

constexpr int MT_THREADS = 256;
constexpr int MT_STRIP = MT_THREADS * 4;  // columns per CTA
constexpr int MT_BAND = 32;               // rows per CTA

__device__ __forceinline__ unsigned long long pack_key(float v, unsigned int idx, bool largest) {
  unsigned int o = float_to_ordered(v);
  if (!largest) o = ~o;
  return ((unsigned long long)o << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float key_value(unsigned long long k, bool largest) {
  unsigned int o = (unsigned int)(k >> 32);
  if (!largest) o = ~o;
  return ordered_to_float(o);
}
__device__ __forceinline__ unsigned int key_index(unsigned long long k) { return 0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull); }

__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// grid (strips, bands, B)
template <bool VEC>
__global__ void __launch_bounds__(MT_THREADS) rowcol_best_kernel(const float* __restrict__ x, int N, int M, int largest,
                                                                 unsigned long long* __restrict__ rowbest,
                                                                 unsigned long long* __restrict__ colbest) {
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * MT_BAND;
  const int i1 = min(N, i0 + MT_BAND);
  const int lane = threadIdx.x & 31;
  const bool lg = largest != 0;
  const float* xb = x + (size_t)b * N * M;
  // columns owned by this thread
  int cols[4];
#pragma unroll
  for (int e = 0; e < 4; ++e)
    cols[e] = VEC ? (blockIdx.x * MT_STRIP + 4 * (int)threadIdx.x + e) : (blockIdx.x * MT_STRIP + (int)threadIdx.x + MT_THREADS * e);
  unsigned long long cbest[4] = {0ull, 0ull, 0ull, 0ull};
  for (int i = i0; i < i1; ++i) {
    float v[4];
    if constexpr (VEC) {
      if (cols[0] < M) {
        const float4 t = *reinterpret_cast<const float4*>(xb + (size_t)i * M + cols[0]);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (cols[e] < M) v[e] = xb[(size_t)i * M + cols[e]];
    }
    unsigned long long rb = 0ull;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (cols[e] < M) {
        const unsigned long long kr = pack_key(v[e], (unsigned int)cols[e], lg);
        rb = kr > rb ? kr : rb;
        const unsigned long long kc = pack_key(v[e], (unsigned int)i, lg);
        cbest[e] = kc > cbest[e] ? kc : cbest[e];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, rb, o);
      rb = other > rb ? other : rb;
    }
    if (lane == 0 && rb != 0ull) atomicMax(&rowbest[(size_t)b * N + i], rb);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (cols[e] < M && cbest[e] != 0ull) atomicMax(&colbest[(size_t)b * M + cols[e]], cbest[e]);
}

// k-th best packed key of every row (one warp per row) / of every column (one warp per column, lanes stride over the
// rows: uncoalesced, but k > 1 only occurs on the small fine-level patch matrices).  Each lane keeps its own k best keys
// sorted in registers; k rounds of a warp arg-max then pop the global best k times.  Fewer than k entries: key 0 (every
// entry is selected, as torch.topk would fail / select all).  k <= TOPK_MAX.
constexpr int TOPK_MAX = 8;
template <bool COLS>
__global__ void __launch_bounds__(256) topk_line_kernel(const float* __restrict__ x, int B, int N, int M, int k, int largest,
                                                        unsigned long long* __restrict__ kth) {
  const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const int nlines = COLS ? B * M : B * N;
  if (warp >= nlines) return;
  const int L = COLS ? M : N, len = COLS ? N : M;        // lines per batch element, elements per line
  const int b = warp / L, l = warp - b * L;
  const float* base = x + (size_t)b * N * M + (COLS ? (size_t)l : (size_t)l * M);
  const size_t stride = COLS ? (size_t)M : 1;
  const bool lg = largest != 0;
  unsigned long long best[TOPK_MAX];
#pragma unroll
  for (int q = 0; q < TOPK_MAX; ++q) best[q] = 0ull;
  for (int e = lane; e < len; e += 32) {
    unsigned long long key = pack_key(base[(size_t)e * stride], (unsigned int)e, lg);
#pragma unroll
    for (int q = 0; q < TOPK_MAX; ++q) {   // sorted insertion (descending)
      if (q < k && key > best[q]) {
        const unsigned long long t = best[q];
        best[q] = key;
        key = t;
      }
    }
  }
  unsigned long long kth_key = 0ull;
  for (int r = 0; r < k; ++r) {
    unsigned long long head = best[0], m = head;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
      m = other > m ? other : m;
    }
    kth_key = m;
    if (m == 0ull) break;                  // fewer than k entries in the line
    if (head == m) {                       // keys are distinct (distinct indices): exactly one lane pops
#pragma unroll
      for (int q = 0; q + 1 < TOPK_MAX; ++q) best[q] = best[q + 1];
      best[TOPK_MAX - 1] = 0ull;
    }
  }
  if (lane == 0) kth[warp] = (len < k) ? 0ull : kth_key;
}

struct MatchParams {
  const float* x;
  int B, N, M;
  int mode;     // 0: get_match (value equality), 1: top-1 select (index based), 2: top-k select (rowbest / colbest hold the k-th best keys)
  const unsigned char* row_mask;  // mode 2, optional [B,N]: hits in masked rows are dropped AFTER the selection
  const unsigned char* col_mask;  // mode 2, optional [B,M]
  int mutual;
  int has_thr;
  float thr;
  int largest;
  const unsigned long long* rowbest;
  const unsigned long long* colbest;
  int* counts;         // [B*N]
  const int* offsets;  // [B*N + 1]
  long long* index_out;  // [K,3]
  float* val_out;        // [K]
  unsigned char* mask_out;  // [B,N,M] or NULL
  long long capacity;       // slots in index_out / val_out; hits beyond it are dropped
};

__device__ __forceinline__ bool is_hit(const MatchParams& p, int b, int i, int j, float v, float rowv, unsigned int rowj) {
  bool pass_thr = true;
  if (p.has_thr) pass_thr = p.largest ? (v > p.thr) : (v < p.thr);
  if (p.mode == 0) {
    // conf > thr [& conf == row max & conf == column max]      matching.py:73-80
    if (!pass_thr) return false;
    if (!p.mutual) return true;
    if (v != rowv) return false;
    return v == key_value(p.colbest[(size_t)b * p.M + j], true);
  }
  if (p.mode == 2) {
    // within the top-k of the row / of the column (keys >= the k-th best key), AND / OR, threshold, masks
    //                                                             mutual_topk_select.py:30-57, 95-127
    const bool lg = p.largest != 0;
    const bool row_in = pack_key(v, (unsigned int)j, lg) >= p.rowbest[(size_t)b * p.N + i];
    const bool col_in = pack_key(v, (unsigned int)i, lg) >= p.colbest[(size_t)b * p.M + j];
    bool h = (p.mutual ? (row_in && col_in) : (row_in || col_in)) && pass_thr;
    if (h && p.row_mask) h = p.row_mask[(size_t)b * p.N + i] != 0;
    if (h && p.col_mask) h = p.col_mask[(size_t)b * p.M + j] != 0;
    return h;
  }
  // top-1 of the row / of the column, AND (mutual) or OR        mutual_topk_select.py:30-50
  const bool row_hit = ((unsigned int)j == rowj);
  const bool col_hit = (key_index(p.colbest[(size_t)b * p.M + j]) == (unsigned int)i);
  const bool hit = p.mutual ? (row_hit && col_hit) : (row_hit || col_hit);
  return hit && pass_thr;
}

// one warp per row; WRITE = false counts, WRITE = true emits in column order
template <bool WRITE>
__global__ void __launch_bounds__(256) match_rows_kernel(const MatchParams p) {
  const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const int nrows = p.B * p.N;
  if (warp >= nrows) return;
  const int b = warp / p.N, i = warp - b * p.N;
  const float* row = p.x + ((size_t)b * p.N + i) * p.M;
  const unsigned long long rk = p.rowbest[(size_t)b * p.N + i];
  const float rowv = key_value(rk, p.largest != 0);
  const unsigned int rowj = key_index(rk);
  // Fast exit for the mutual cases: a hit needs the row best, whose column is known (mode 1) or
  // whose value must also be a column best (mode 0; ties need the full scan, so only mode 1 skips).
  int base = WRITE ? p.offsets[warp] : 0;
  int count = 0;
  if (p.mode == 1 && p.mutual && !(WRITE && p.mask_out)) {
    if (lane == 0) {
      const int j = (int)rowj;
      if (rk != 0ull && is_hit(p, b, i, j, row[j], rowv, rowj)) {
        if (WRITE && base < p.capacity) {
          long long* o = p.index_out + (size_t)base * 3;
          o[0] = b; o[1] = i; o[2] = j;
          p.val_out[base] = row[j];
        }
        count = 1;
      }
      if (!WRITE) p.counts[warp] = count;
    }
    return;
  }
  for (int j0 = 0; j0 < p.M; j0 += 32) {
    const int j = j0 + lane;
    bool hit = false;
    float v = 0.f;
    if (j < p.M) {
      v = row[j];
      hit = is_hit(p, b, i, j, v, rowv, rowj);
      if (WRITE && p.mask_out) p.mask_out[((size_t)b * p.N + i) * p.M + j] = hit ? 1 : 0;
    }
    const unsigned int bal = __ballot_sync(0xffffffffu, hit);
    if (WRITE && hit) {
      const int pos = base + count + __popc(bal & ((1u << lane) - 1u));
      if (pos < p.capacity) {
        long long* o = p.index_out + (size_t)pos * 3;
        o[0] = b; o[1] = i; o[2] = j;
        p.val_out[pos] = v;
      }
    }
    count += __popc(bal);
  }
  if (!WRITE && lane == 0) p.counts[warp] = count;
}

// single-CTA exclusive scan: offsets[0..n] (offsets[n] = total), total also to *total_out
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int* __restrict__ counts, int n, int* __restrict__ offsets,
                                                            int* __restrict__ total_out) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int idx = base + tid;
    const int c = idx < n ? counts[idx] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    const int before = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + incl - c;
    if (idx < n) offsets[idx] = before;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
  if (tid == 0) {
    offsets[n] = carry_s;
    *total_out = carry_s;
  }
}

// mutual top-1 matches from packed bests: one CTA, ordered compaction over the B*N rows
__global__ void __launch_bounds__(1024) match_from_best_kernel(const unsigned long long* __restrict__ rowbest,
                                                               const unsigned long long* __restrict__ colbest, int B, int N, int M,
                                                               int has_thr, float thr, long long* __restrict__ index_out,
                                                               float* __restrict__ val_out, long long capacity,
                                                               int* __restrict__ total_out) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  const int nrows = B * N;
  for (int base = 0; base < nrows; base += 1024) {
    const int r = base + tid;
    bool hit = false;
    int b = 0, i = 0, j = 0;
    float v = 0.f;
    if (r < nrows) {
      b = r / N;
      i = r - b * N;
      const unsigned long long rk = rowbest[r];
      if (rk != 0ull) {
        j = (int)key_index(rk);
        v = key_value(rk, true);
        const unsigned long long ck = colbest[(size_t)b * M + j];
        hit = (ck != 0ull) && (key_index(ck) == (unsigned int)i) && (!has_thr || v > thr);
      }
    }
    const unsigned int bal = __ballot_sync(0xffffffffu, hit);
    const int wcount = __popc(bal);
    if (lane == 0) warp_sums[warp] = wcount;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    if (hit) {
      const long long pos = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + __popc(bal & ((1u << lane) - 1u));
      if (pos < capacity) {
        long long* o = index_out + pos * 3;
        o[0] = b; o[1] = i; o[2] = j;
        val_out[pos] = v;
      }
    }
    __syncthreads();
    if (tid == 0) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
  if (tid == 0) *total_out = carry_s;
}

struct MatchWorkspace {
  unsigned long long* rowbest;
  unsigned long long* colbest;
  int* counts;
  int* offsets;
  size_t total;
};

static MatchWorkspace match_carve(void* ws, int B, int N, int M) {
  MatchWorkspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* ptr = ws ? (void*)((char*)ws + off) : nullptr;
    off += align_up(bytes, 256);
    return ptr;
  };
  w.rowbest = (unsigned long long*)take(8ull * B * N);
  w.colbest = (unsigned long long*)take(8ull * B * M);
  w.counts = (int*)take(4ull * B * N);
  w.offsets = (int*)take(4ull * ((size_t)B * N + 1));
  w.total = off;
  return w;
}

static MatchParams match_params(const float* x, int B, int N, int M, int mode, int mutual, int has_thr, float thr, int largest,
                                const MatchWorkspace& w) {
  MatchParams p{};
  p.x = x;
  p.B = B;
  p.N = N;
  p.M = M;
  p.mode = mode;
  p.mutual = mutual;
  p.has_thr = has_thr;
  p.thr = thr;
  p.largest = largest;
  p.rowbest = w.rowbest;
  p.colbest = w.colbest;
  p.counts = w.counts;
  p.offsets = w.offsets;
  return p;
}

}  // namespace drg

using namespace drg;

extern "C" size_t drg_match_workspace_bytes(int B, int N, int M) {
  if (B < 1 || N < 1 || M < 1) return 0;
  return match_carve(nullptr, B, N, M).total;
}

static int match_check(const float* x, int B, int N, int M, int mode, void* ws, size_t ws_bytes) {
  DRG_CHECK_ARG(x != nullptr, "matrix is null");
  DRG_CHECK_ARG(B >= 1 && N >= 1 && M >= 1, "B, N, M must be >= 1");
  DRG_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 (get_match) or 1 (top-1 select)");
  DRG_CHECK_ARG((long long)B * N < (1ll << 31) && (long long)N * M < (1ll << 32), "matrix too large");
  if (ws == nullptr || ws_bytes < match_carve(nullptr, B, N, M).total || ((uintptr_t)ws & 255u)) {
    set_error("match: workspace missing, too small or not 256-byte aligned");
    return DRG_ERR_WORKSPACE;
  }
  return DRG_OK;
}

extern "C" int drg_match_count(const float* x, int B, int N, int M, int mode, int mutual, int has_thr, float thr, int largest,
                               void* ws, size_t ws_bytes, int* total_out, void* stream) {
  int rc = match_check(x, B, N, M, mode, ws, ws_bytes);
  if (rc != DRG_OK) return rc;
  DRG_CHECK_ARG(total_out != nullptr, "total_out is null");
  DRG_CHECK_ARG(mode == 1 || largest, "get_match is defined for largest only");
  cudaStream_t st = (cudaStream_t)stream;
  MatchWorkspace w = match_carve(ws, B, N, M);
  const bool need_best = (mode == 1) || mutual;
  if (need_best) {
    const size_t nfill = (size_t)(((char*)w.counts - (char*)w.rowbest) / 8);  // rowbest + colbest (contiguous, padded)
    fill_u64_kernel<<<(int)std::min<size_t>((nfill + 255) / 256, 1024), 256, 0, st>>>(w.rowbest, nfill, 0ull);
    DRG_LAUNCH_CHECK();
    dim3 grid((M + MT_STRIP - 1) / MT_STRIP, (N + MT_BAND - 1) / MT_BAND, B);
    const bool vec = (M % 4 == 0) && (((uintptr_t)x & 15u) == 0);
    if (vec)
      {
        ProfScope prof_scope(PROF_ROWCOL_BEST, st);
        rowcol_best_kernel<true><<<grid, MT_THREADS, 0, st>>>(x, N, M, largest, w.rowbest, w.colbest);
      }
    else
      {
        ProfScope prof_scope(PROF_ROWCOL_BEST, st);
        rowcol_best_kernel<false><<<grid, MT_THREADS, 0, st>>>(x, N, M, largest, w.rowbest, w.colbest);
      }
    DRG_LAUNCH_CHECK();
  }
  MatchParams p = match_params(x, B, N, M, mode, mutual, has_thr, thr, largest, w);
  const int nrows = B * N;
  {
    ProfScope prof_scope(PROF_MATCH_ROWS, st);
    match_rows_kernel<false><<<(nrows + 7) / 8, 256, 0, st>>>(p);
  }
  DRG_LAUNCH_CHECK();
  scan_counts_kernel<<<1, 1024, 0, st>>>(w.counts, nrows, w.offsets, total_out);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_match_write(const float* x, int B, int N, int M, int mode, int mutual, int has_thr, float thr, int largest,
                               void* ws, size_t ws_bytes, long long* index_out, float* val_out, long long capacity,
                               unsigned char* mask_out, void* stream) {
  int rc = match_check(x, B, N, M, mode, ws, ws_bytes);
  if (rc != DRG_OK) return rc;
  DRG_CHECK_ARG(index_out != nullptr && val_out != nullptr, "index_out/val_out are null (allocate at least one element)");
  DRG_CHECK_ARG(capacity >= 1, "capacity must be >= 1");
  MatchWorkspace w = match_carve(ws, B, N, M);
  MatchParams p = match_params(x, B, N, M, mode, mutual, has_thr, thr, largest, w);
  p.index_out = index_out;
  p.val_out = val_out;
  p.mask_out = mask_out;
  p.capacity = capacity;
  const int nrows = B * N;
  match_rows_kernel<true><<<(nrows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(p);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

/* top-k (k >= 1) selection: see include/diffreg_b200.h */
static int topk_prepare(const float* x, int B, int N, int M, int k, int largest, const MatchWorkspace& w, cudaStream_t st) {
  topk_line_kernel<false><<<(B * N + 7) / 8, 256, 0, st>>>(x, B, N, M, k, largest, w.rowbest);
  DRG_LAUNCH_CHECK();
  topk_line_kernel<true><<<(B * M + 7) / 8, 256, 0, st>>>(x, B, N, M, k, largest, w.colbest);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_topk_match_count(const float* x, int B, int N, int M, int k, int mutual, int has_thr, float thr, int largest,
                                    const unsigned char* row_mask, const unsigned char* col_mask, void* ws, size_t ws_bytes,
                                    int* total_out, void* stream) {
  int rc = match_check(x, B, N, M, 1, ws, ws_bytes);
  if (rc != DRG_OK) return rc;
  DRG_CHECK_ARG(total_out != nullptr, "total_out is null");
  DRG_CHECK_ARG(k >= 1 && k <= TOPK_MAX, "k must be in 1..8");
  cudaStream_t st = (cudaStream_t)stream;
  MatchWorkspace w = match_carve(ws, B, N, M);
  rc = topk_prepare(x, B, N, M, k, largest, w, st);
  if (rc != DRG_OK) return rc;
  MatchParams p = match_params(x, B, N, M, 2, mutual, has_thr, thr, largest, w);
  p.row_mask = row_mask;
  p.col_mask = col_mask;
  const int nrows = B * N;
  match_rows_kernel<false><<<(nrows + 7) / 8, 256, 0, st>>>(p);
  DRG_LAUNCH_CHECK();
  scan_counts_kernel<<<1, 1024, 0, st>>>(w.counts, nrows, w.offsets, total_out);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_topk_match_write(const float* x, int B, int N, int M, int k, int mutual, int has_thr, float thr, int largest,
                                    const unsigned char* row_mask, const unsigned char* col_mask, void* ws, size_t ws_bytes,
                                    long long* index_out, float* val_out, long long capacity, unsigned char* mask_out, void* stream) {
  int rc = match_check(x, B, N, M, 1, ws, ws_bytes);
  if (rc != DRG_OK) return rc;
  DRG_CHECK_ARG(index_out != nullptr && val_out != nullptr, "index_out/val_out are null (allocate at least one element)");
  DRG_CHECK_ARG(capacity >= 1 && k >= 1 && k <= TOPK_MAX, "capacity must be >= 1 and k in 1..8");
  MatchWorkspace w = match_carve(ws, B, N, M);
  MatchParams p = match_params(x, B, N, M, 2, mutual, has_thr, thr, largest, w);
  p.row_mask = row_mask;
  p.col_mask = col_mask;
  p.index_out = index_out;
  p.val_out = val_out;
  p.mask_out = mask_out;
  p.capacity = capacity;
  match_rows_kernel<true><<<(B * N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(p);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_match_from_best(const unsigned long long* rowbest, const unsigned long long* colbest, int B, int N, int M,
                                   int has_thr, float thr, long long* index_out, float* val_out, long long capacity, int* total_out,
                                   void* stream) {
  DRG_CHECK_ARG(rowbest && colbest && index_out && val_out && total_out, "rowbest/colbest/index_out/val_out/total_out must be non-null");
  DRG_CHECK_ARG(B >= 1 && N >= 1 && M >= 1 && capacity >= 1, "B, N, M, capacity must be >= 1");
  DRG_CHECK_ARG((long long)B * N < (1ll << 31), "too many rows");
  {
    ProfScope prof_scope(PROF_MATCH_ROWS, (cudaStream_t)stream);
    match_from_best_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rowbest, colbest, B, N, M, has_thr, thr, index_out, val_out, capacity,
                                                                 total_out);
  }
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}
