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
//   1. topk_threshold_kernel  one CTA per batch element: mask counts -> K_b; 32768 hashed samples of the
//                             matrix into shared memory; EXACT radix select of the t-th largest sample
//                             (64-bit key = ordered value << 32 | ~flat index, so ties are ordered too),
//                             t ~ 2x the expected number of top-K_b entries in the sample -> lower bound L
//   2. topk_collect_kernel    ONE pass over the matrix: entries with key >= L are appended to a candidate
//                             list (warp-aggregated atomics); everything else is never touched again
//   3. procr_solve_kernel     one CTA per batch element: exact radix select of the K_b largest
//                             candidates, fp64 weighted moments, closed-form 3x3 SVD (one-sided Jacobi,
//                             fp64), reflection fix, condition-number gate, and the src-point warp.
// Because L is an order statistic of the sample (not a histogram bin edge) the candidate list holds
// ~2 K_b + 32 N M / 32768 entries whatever the value distribution (flat, tied or all-zero matrices
// included).  If the sample still misleads (fewer than K_b candidates) the solve kernel falls back to
// collecting the whole matrix itself: slow, but exact.
#include "common.cuh"

namespace drg {

constexpr int TK_BINS = 2048;
constexpr int TS_THREADS = 512;
constexpr int TS_SAMPLES = 32768;  // sample keys held in shared memory (128 KB)
constexpr int SOLVE_THREADS = 1024;

struct ProcrState {  // per batch element
  int Kb;                         // number of correspondences to use
  unsigned int n_cand;            // candidates appended so far
  unsigned long long lower_key;   // candidates have 64-bit key >= lower_key
};

struct ProcrParams {
  const float* conf;             // [B,N,M]
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

__device__ __forceinline__ unsigned int hash_u32(unsigned int x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

__device__ __forceinline__ unsigned long long make_key64(unsigned int ordered_value, unsigned int flat_index) {
  return ((unsigned long long)ordered_value << 32) | (unsigned long long)(0xFFFFFFFFu - flat_index);
}

// Exact selection of the k-th largest of n distinct 64-bit keys by one CTA (1 <= k <= n).
// key_at(e) returns the key of element e.  Six radix levels (11,11,10,11,11,10 bits, MSB first); stops
// early once the remaining bucket is wanted whole.  Returns T such that exactly k keys are >= T.
// scratch: hist[TK_BINS] plus three words of shared memory.
struct SelectScratch {
  unsigned int hist[TK_BINS];
  unsigned long long prefix;
  int krem;
  int done;
};

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
      // warp 0 walks the histogram from the top: lane l owns the l-th chunk of bins (descending)
      const int nb = 1 << wbits, chunk = nb >> 5;
      const int hi = nb - 1 - chunk * tid;  // highest bin of this lane's chunk
      unsigned int local = 0u;
      for (int d = hi; d > hi - chunk; --d) local += sc.hist[d];
      unsigned int incl = local;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int tmp = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += tmp;
      }
      const unsigned int krem = (unsigned int)sc.krem;
      const unsigned int crossing = __ballot_sync(0xffffffffu, incl >= krem);
      const int owner = crossing ? (__ffs(crossing) - 1) : 31;
      if (tid == owner) {
        unsigned int cum = incl - local;
        int d = hi;
        for (; d > hi - chunk + 1; --d) {
          if (cum + sc.hist[d] >= krem) break;
          cum += sc.hist[d];
        }
        sc.prefix = (prefix << wbits) | (unsigned long long)d;
        sc.krem = (int)(krem - cum);
        if (sc.hist[d] == krem - cum) sc.done = 1;  // the whole bucket is wanted
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
  extern __shared__ unsigned int sample_key[];  // [TS_SAMPLES]
  __shared__ SelectScratch sc;
  __shared__ unsigned long long cnt_s;
  const int b = blockIdx.x, tid = threadIdx.x;
  // ---- mask counts of every batch element (K is a mean over the batch)      procrustes.py:61-65
  float cap_sum = 0.f;
  int my_cap = 0;
  for (int bb = 0; bb < p.B; ++bb) {
    int ns = p.N, nt = p.M;
    if (!p.padded_lengths) {
      if (tid == 0) cnt_s = 0ull;
      __syncthreads();
      unsigned long long c = 0ull;
      for (int i = tid; i < p.N; i += TS_THREADS) c += p.src_mask[(size_t)bb * p.N + i] ? (1ull << 32) : 0ull;
      for (int j = tid; j < p.M; j += TS_THREADS) c += p.tgt_mask[(size_t)bb * p.M + j] ? 1ull : 0ull;
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

  // ---- sample the matrix
  const size_t total = (size_t)p.N * p.M;
  const float* x = p.conf + (size_t)b * total;
  const bool all = total <= (size_t)TS_SAMPLES;
  const unsigned int n_s = all ? (unsigned int)total : (unsigned int)TS_SAMPLES;
  const unsigned int salt = 0x9e3779b9u * (unsigned int)(b + 1);
  auto sample_pos = [&](unsigned int q) -> unsigned int {
    return all ? q : (unsigned int)(((unsigned long long)hash_u32(q + salt) * (unsigned long long)total) >> 32);
  };
  for (unsigned int q = tid; q < n_s; q += TS_THREADS) sample_key[q] = float_to_ordered(x[sample_pos(q)]);
  __syncthreads();

  // ---- the t-th largest sample key is the lower bound
  unsigned long long lower = 0ull;
  if (Kb > 0) {
    long long target;
    if (all) {
      target = Kb;  // the sample IS the matrix: the bound is the exact K_b-th largest
    } else {
      // twice the expected number of top-K_b entries inside the sample, plus slack for small counts
      const double expect = (double)Kb * ((double)n_s / (double)total);
      target = (long long)(2.0 * expect + 32.0);
    }
    if (target < (long long)n_s) {
      lower = block_select_kth<TS_THREADS>(
          [&](size_t e) { return make_key64(sample_key[e], sample_pos((unsigned int)e)); }, (size_t)n_s, (int)target, sc);
    }
  }
  if (tid == 0) {
    ProcrState s;
    s.Kb = Kb;
    s.n_cand = 0u;
    s.lower_key = lower;
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
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (take[e]) {
      p.cand_key[(size_t)b * total + pos] = key[e];
      p.cand_idx[(size_t)b * total + pos] = flat0 + e;
      ++pos;
    }
  }
}

__global__ void __launch_bounds__(256) topk_collect_kernel(const ProcrParams p) {
  const int b = blockIdx.y;
  const size_t total = (size_t)p.N * p.M;
  const float* x = p.conf + (size_t)b * total;
  const unsigned long long lower = p.state[b].lower_key;
  const bool vec = ((total & 3) == 0) && ((((uintptr_t)x) & 15u) == 0);
  const size_t n4 = vec ? (total >> 2) : ((total + 3) >> 2);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
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
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (q * 4 + e < total) ? x[q * 4 + e] : -INFINITY;
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

// ---- 5. select + Kabsch + warp ----------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* red /*[32]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  const int nw = (blockDim.x + 31) >> 5;
  for (int w = 0; w < nw; ++w) s += red[w];  // same order in every thread: bitwise identical
  return s;
}

// One-sided Jacobi SVD of a 3x3 matrix (fp64): A = U diag(s) V^T, singular values descending.
__device__ void svd3x3(const double A[3][3], double U[3][3], double s[3], double V[3][3]) {
  double W[3][3];  // working copy, columns get orthogonalised: W = A V
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      W[i][j] = A[i][j];
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int pcol = 0; pcol < 2; ++pcol) {
      for (int qcol = pcol + 1; qcol < 3; ++qcol) {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int i = 0; i < 3; ++i) {
          alpha += W[i][pcol] * W[i][pcol];
          beta += W[i][qcol] * W[i][qcol];
          gamma += W[i][pcol] * W[i][qcol];
        }
        if (gamma == 0.0) continue;
        off = fmax(off, fabs(gamma) / sqrt(fmax(alpha * beta, 1e-300)));
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + tt * tt), sn = c * tt;
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
    if (off < 1e-15) break;
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

// From the weighted raw moments to (R, t, condition); thread 0 only.
//   sw = sum |w|, sx = sum w x, sy = sum w y, sxy[a][c] = sum w y_a x_c  (all fp64)
__device__ void kabsch_from_moments(double sw_abs, double sw, const double sx[3], const double sy[3], const double syx[3][3],
                                    double eps, float R_out[9], float t_out[3], double* cond_out) {
  // w_norm = w / (sum|w| + eps): the normalised weights sum to slightly less than one  (procrustes.py:29-30)
  const double inv = 1.0 / (sw_abs + eps);
  const double sn = sw * inv;
  double mx[3], my[3];
  for (int a = 0; a < 3; ++a) {
    mx[a] = sx[a] * inv;
    my[a] = sy[a] * inv;
  }
  // Sxy = sum w_norm (y - my)(x - mx)^T = sum w_norm y x^T - (2 - sn) my mx^T      (procrustes.py:34)
  double S[3][3];
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) S[a][c] = syx[a][c] * inv - (2.0 - sn) * my[a] * mx[c];
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
  __shared__ double red[32];
  __shared__ unsigned int n_emit;
  __shared__ float pose_s[12];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const size_t total = (size_t)p.N * p.M;
  const ProcrState st = p.state[b];
  const int Kb = st.Kb;
  unsigned int* ckey = p.cand_key + (size_t)b * total;
  unsigned int* cidx = p.cand_idx + (size_t)b * total;
  size_t n = st.n_cand;
  if (n < (size_t)Kb) {
    // fallback: the sample-based bound left too few candidates; take the whole matrix
    const float* x = p.conf + (size_t)b * total;
    for (size_t e = tid; e < total; e += SOLVE_THREADS) {
      ckey[e] = float_to_ordered(x[e]);
      cidx[e] = (unsigned int)e;
    }
    n = total;
    __syncthreads();
  }

  // ---- exact radix select of the Kb largest 64-bit keys (value << 32 | ~index): no ties
  unsigned long long T = 0ull;  // select keys >= T
  if (Kb > 0 && (size_t)Kb < n)
    T = block_select_kth<SOLVE_THREADS>([&](size_t e) { return make_key64(ckey[e], cidx[e]); }, n, Kb, sc);

  // ---- emit the selection and accumulate the weighted moments in fp64
  if (tid == 0) n_emit = 0u;
  __syncthreads();
  const float* sp = p.src_pcd + (size_t)b * p.N * 3;
  const float* tp = p.tgt_pcd + (size_t)b * p.M * 3;
  double a_w = 0.0, a_wabs = 0.0, a_x[3] = {0, 0, 0}, a_y[3] = {0, 0, 0}, a_yx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  if (Kb > 0) {
    for (size_t e = tid; e < n; e += SOLVE_THREADS) {
      const unsigned int k32 = ckey[e], fi = cidx[e];
      if (make_key64(k32, fi) >= T) {
        const float wf = ordered_to_float(k32);
        const int i = (int)(fi / (unsigned int)p.M), j = (int)(fi - (unsigned int)i * (unsigned int)p.M);
        if (p.sel_w) {
          const unsigned int pos = atomicAdd(&n_emit, 1u);
          if (pos < (unsigned int)p.K_max) {
            p.sel_w[(size_t)b * p.K_max + pos] = wf;
            p.sel_src[(size_t)b * p.K_max + pos] = i;
            p.sel_tgt[(size_t)b * p.K_max + pos] = j;
          }
        }
        const double w = (double)wf;
        const double x[3] = {(double)sp[i * 3 + 0], (double)sp[i * 3 + 1], (double)sp[i * 3 + 2]};
        const double y[3] = {(double)tp[j * 3 + 0], (double)tp[j * 3 + 1], (double)tp[j * 3 + 2]};
        a_w += w;
        a_wabs += fabs(w);
        for (int a = 0; a < 3; ++a) {
          a_x[a] += w * x[a];
          a_y[a] += w * y[a];
          for (int c = 0; c < 3; ++c) a_yx[a][c] += w * y[a] * x[c];
        }
      }
    }
  }
  double m_w = block_sum(a_w, red), m_wabs = block_sum(a_wabs, red);
  double m_x[3], m_y[3], m_yx[3][3];
  for (int a = 0; a < 3; ++a) {
    m_x[a] = block_sum(a_x[a], red);
    m_y[a] = block_sum(a_y[a], red);
    for (int c = 0; c < 3; ++c) m_yx[a][c] = block_sum(a_yx[a][c], red);
  }
  if (p.sel_w) {
    __syncthreads();
    for (int k = (int)n_emit + tid; k < p.K_max; k += SOLVE_THREADS) {
      p.sel_w[(size_t)b * p.K_max + k] = 0.f;
      p.sel_src[(size_t)b * p.K_max + k] = 0;
      p.sel_tgt[(size_t)b * p.K_max + k] = 0;
    }
  }
  if (tid == 0) {
    float R[9], t[3];
    double cond;
    kabsch_from_moments(m_wabs, m_w, m_x, m_y, m_yx, 1e-4, R, t, &cond);
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
  __shared__ double red[32];
  const int b = blockIdx.x;
  const float* X = p.X + (size_t)b * p.K * 3;
  const float* Y = p.Y + (size_t)b * p.K * 3;
  const float* w = p.w + (size_t)b * p.K;
  double a_w = 0.0, a_wabs = 0.0, a_x[3] = {0, 0, 0}, a_y[3] = {0, 0, 0}, a_yx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
    const double wk = (double)w[k];
    a_w += wk;
    a_wabs += fabs(wk);
    for (int a = 0; a < 3; ++a) {
      const double xa = (double)X[k * 3 + a], ya = (double)Y[k * 3 + a];
      a_x[a] += wk * xa;
      a_y[a] += wk * ya;
      for (int c = 0; c < 3; ++c) a_yx[a][c] += wk * ya * (double)X[k * 3 + c];
    }
  }
  double m_w = block_sum(a_w, red), m_wabs = block_sum(a_wabs, red);
  double m_x[3], m_y[3], m_yx[3][3];
  for (int a = 0; a < 3; ++a) {
    m_x[a] = block_sum(a_x[a], red);
    m_y[a] = block_sum(a_y[a], red);
    for (int c = 0; c < 3; ++c) m_yx[a][c] = block_sum(a_yx[a][c], red);
  }
  if (threadIdx.x == 0) {
    float R[9], t[3];
    double cond;
    kabsch_from_moments(m_wabs, m_w, m_x, m_y, m_yx, (double)p.eps, R, t, &cond);
    for (int k = 0; k < 9; ++k) p.R[b * 9 + k] = R[k];
    for (int k = 0; k < 3; ++k) p.t[b * 3 + k] = t[k];
    p.condition[b] = cond;
  }
}

struct ProcrWorkspace {
  ProcrState* state;
  unsigned int* cand_key;
  unsigned int* cand_idx;
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
  w.total = off;
  return w;
}

}  // namespace drg

using namespace drg;

extern "C" size_t drg_soft_procrustes_workspace_bytes(int B, int N, int M) {
  if (B < 1 || N < 1 || M < 1) return 0;
  return procr_carve(nullptr, B, N, M).total;
}

extern "C" int drg_soft_procrustes(const drg_procrustes_args* a, void* workspace, size_t workspace_bytes, void* stream) {
  DRG_CHECK_ARG(a != nullptr, "args is null");
  DRG_CHECK_ARG(a->conf && a->src_pcd && a->tgt_pcd, "conf/src_pcd/tgt_pcd must be non-null");
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
  p.conf = a->conf;
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

  static bool attr_set = false;
  if (!attr_set) {
    DRG_CUDA(cudaFuncSetAttribute(topk_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SAMPLES * 4));
    attr_set = true;
  }
  {
    ProfScope prof_scope(PROF_TOPK_THRESHOLD, st);
    topk_threshold_kernel<<<B, TS_THREADS, TS_SAMPLES * 4, st>>>(p);
  }
  DRG_LAUNCH_CHECK();
  int gx = (NUM_SMS * 8) / B;
  if (gx < 1) gx = 1;
  const long long n4 = ((long long)N * M + 3) / 4;
  if ((long long)gx * 256 > n4) gx = (int)((n4 + 255) / 256);
  {
    ProfScope prof_scope(PROF_TOPK_COLLECT, st);
    topk_collect_kernel<<<dim3(gx, B), 256, 0, st>>>(p);
  }
  DRG_LAUNCH_CHECK();
  {
    ProfScope prof_scope(PROF_PROCR_SOLVE, st);
    procr_solve_kernel<<<B, SOLVE_THREADS, 0, st>>>(p);
  }
  DRG_LAUNCH_CHECK();
  return DRG_OK;
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
