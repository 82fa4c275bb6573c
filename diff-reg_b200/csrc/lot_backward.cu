// Backward pass of the log-domain Sinkhorn (training path, SURVEY.md section 8f rank 3: autograd through
// log_optimal_transport, Diff-Reg-4dmatch/models/matching.py:6-38 as torch records it for `loss.backward()`).  sm_100a.
//
// Forward (matching.py:30-36), Z the (N+1) x (M+1) score matrix with the dustbin row / column holding alpha:
//     u_t = log_mu - LSE_j(Z + v_{t-1}),   v_t = log_nu - LSE_i(Z + u_t),   t = 1..I,  v_0 = 0;   out = Z + u_I + v_I - norm.
// With G = dL/d out the chain rule unrolls into 2 I "weighted exp" mat-vecs over Z and one final pass:
//     gv_I = colsum(G);  gu_I = rowsum(G) - sum_j gv_I[j] Pc_I[i,j]
//     for t = I .. 1:    gv_{t-1}[j] = - sum_i gu_t[i] Pr_t[i,j];     gu_{t-1}[i] = - sum_j gv_{t-1}[j] Pc_{t-1}[i,j]
//     dL/dZ = G - sum_t ( gv_t[j] Pc_t[i,j] + gu_t[i] Pr_t[i,j] )
// where Pc_t = exp(Z + u_t + v_t - log_nu) (the column softmax of the v_t step) and Pr_t = exp(Z + v_{t-1} + u_t - log_mu) (the row
// softmax of the u_t step) -- entries <= 1, so nothing overflows; -inf scores give P = 0.  dL/d scores = dL/dZ[:N, :M],
// dL/d alpha = the sum of dL/dZ over the dustbin row and column.  Derivation checked against torch's autograd of the unmodified
// reference (oracle.log_optimal_transport_backward, tests/test_oracle_golden.py).  Z is never materialised; every pass is one read of
// the scores (E' bytes): 2 I + 1 reads + 1 write against the ~60 passes autograd makes.  Deterministic (no float atomics).
#include "common.cuh"

namespace drg {

struct LotbParams {
  const float* scores;   // [B, N, M], may hold -inf
  const float* alpha;    // device scalar
  const float* G;        // [B, N+1, M+1] gradient of the output
  const float* u_all;    // [I, B, N+1]: u_t, t = 1..I
  const float* v_all;    // [I, B, M+1]: v_t, t = 1..I (v_0 = 0)
  const float* consts;   // [B, 4]: norm, log_mu of the dustbin row, log_nu of the dustbin column
  float* gu_all;         // [I, B, N+1]
  float* gv_all;         // [I, B, M+1]
  float* colpart;        // [B, nslab, M+1]
  float* gscores;        // [B, N, M]
  float* galpha;         // [B]
  int B, N, M, I, nslab, slab_rows;
};

constexpr int LOTB_SLAB = 64;   // rows per CTA of the column passes / the final pass

__device__ __forceinline__ float lotb_z(const LotbParams& p, int b, int i, int j, float alpha) {
  return (i < p.N && j < p.M) ? p.scores[((size_t)b * p.N + i) * p.M + j] : alpha;
}
__device__ __forceinline__ float lotb_log_mu(const LotbParams& p, int b, int i) { return i < p.N ? p.consts[4 * b] : p.consts[4 * b + 1]; }
__device__ __forceinline__ float lotb_log_nu(const LotbParams& p, int b, int j) { return j < p.M ? p.consts[4 * b] : p.consts[4 * b + 2]; }
__device__ __forceinline__ const float* lotb_u(const LotbParams& p, int t, int b) { return p.u_all + ((size_t)(t - 1) * p.B + b) * (p.N + 1); }
__device__ __forceinline__ const float* lotb_v(const LotbParams& p, int t, int b) { return p.v_all + ((size_t)(t - 1) * p.B + b) * (p.M + 1); }
__device__ __forceinline__ float* lotb_gu(const LotbParams& p, int t, int b) { return p.gu_all + ((size_t)(t - 1) * p.B + b) * (p.N + 1); }
__device__ __forceinline__ float* lotb_gv(const LotbParams& p, int t, int b) { return p.gv_all + ((size_t)(t - 1) * p.B + b) * (p.M + 1); }

// norm = -log(ms + ns), log_mu[N] = log(ns) + norm, log_nu[M] = log(ms) + norm   (matching.py:14-15, 24-27)
__global__ void __launch_bounds__(256) lotb_consts_kernel(const uint8_t* __restrict__ src_mask, const uint8_t* __restrict__ tgt_mask, int N, int M,
                                                         float* __restrict__ consts) {
  __shared__ int red[2][8];
  const int b = blockIdx.x;
  int ms = 0, ns = 0;
  for (int i = threadIdx.x; i < N; i += 256) ms += src_mask[(size_t)b * N + i] != 0;
  for (int j = threadIdx.x; j < M; j += 256) ns += tgt_mask[(size_t)b * M + j] != 0;
  ms = __reduce_add_sync(0xffffffffu, ms);
  ns = __reduce_add_sync(0xffffffffu, ns);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = ms;
    red[1][threadIdx.x >> 5] = ns;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ms = ns = 0;
    for (int w = 0; w < 8; ++w) {
      ms += red[0][w];
      ns += red[1][w];
    }
    const float norm = -logf((float)(ms + ns));
    consts[4 * b] = norm;
    consts[4 * b + 1] = logf((float)ns) + norm;
    consts[4 * b + 2] = logf((float)ms) + norm;
    consts[4 * b + 3] = 0.f;
  }
}

// Row pass of step t (one warp per row of Z): gu_t[i] = base_i - sum_j gv_t[j] exp(Z_ij + u_t[i] + v_t[j] - log_nu[j]),
// base_i = rowsum(G)[i] for t == I (with_G), else 0.
__global__ void __launch_bounds__(256) lotb_row_kernel(const LotbParams p, int t, int with_G) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i > p.N) return;
  const float alpha = *p.alpha;
  const float ui = lotb_u(p, t, b)[i];
  const float* vt = lotb_v(p, t, b);
  const float* gv = lotb_gv(p, t, b);
  const float* Gr = p.G + ((size_t)b * (p.N + 1) + i) * (p.M + 1);
  float acc = 0.f, base = 0.f;
  int j_begin = 0;
  if (i < p.N && (p.M & 3) == 0 && ((uintptr_t)p.scores & 15u) == 0) {
    // interior columns four at a time (the score row is 16-byte aligned); log_nu = norm there
    const float4* zr = reinterpret_cast<const float4*>(p.scores + ((size_t)b * p.N + i) * p.M);
    const float c0 = ui - p.consts[4 * b];
    for (int q = lane; q < (p.M >> 2); q += 32) {
      const float4 z = zr[q];
      const int j = q << 2;
      acc = fmaf(gv[j], expf(z.x + c0 + vt[j]), acc);
      acc = fmaf(gv[j + 1], expf(z.y + c0 + vt[j + 1]), acc);
      acc = fmaf(gv[j + 2], expf(z.z + c0 + vt[j + 2]), acc);
      acc = fmaf(gv[j + 3], expf(z.w + c0 + vt[j + 3]), acc);
      if (with_G) base += (Gr[j] + Gr[j + 1]) + (Gr[j + 2] + Gr[j + 3]);
    }
    j_begin = p.M;      // the dustbin column below
  }
  for (int j = j_begin + lane; j <= p.M; j += 32) {
    const float z = lotb_z(p, b, i, j, alpha);
    acc = fmaf(gv[j], expf(z + ui + (vt[j] - lotb_log_nu(p, b, j))), acc);
    if (with_G) base += Gr[j];
  }
  acc = warp_sum(acc);
  base = warp_sum(base);
  if (lane == 0) lotb_gu(p, t, b)[i] = base - acc;
}

// Column pass: partial sums over a slab of rows, one thread per column.  init: colpart = sum_i G_ij (-> gv_I); else
// colpart = sum_i gu_t[i] exp(Z_ij + u_t[i] - log_mu[i] + v_{t-1}[j])  (-> gv_{t-1} = -sum).
__global__ void __launch_bounds__(256) lotb_col_kernel(const LotbParams p, int t, int init) {
  const int b = blockIdx.z, j = blockIdx.x * 256 + threadIdx.x;
  const int i0 = blockIdx.y * p.slab_rows, i1 = min(i0 + p.slab_rows, p.N + 1);
  if (j > p.M) return;
  const float alpha = *p.alpha;
  float acc = 0.f;
  if (init) {
    for (int i = i0; i < i1; ++i) acc += p.G[((size_t)b * (p.N + 1) + i) * (p.M + 1) + j];
  } else {
    const float* ut = lotb_u(p, t, b);
    const float* gu = lotb_gu(p, t, b);
    const float vp = t > 1 ? lotb_v(p, t - 1, b)[j] : 0.f;
    int i = i0;
    if (j < p.M) {
      const float nrm = p.consts[4 * b];
      const float* zc = p.scores + (size_t)b * p.N * p.M + j;
      for (; i + 4 <= min(i1, p.N); i += 4) {          // four rows in flight (interior rows: log_mu = norm)
        const float z0 = zc[(size_t)i * p.M], z1 = zc[(size_t)(i + 1) * p.M], z2 = zc[(size_t)(i + 2) * p.M], z3 = zc[(size_t)(i + 3) * p.M];
        acc = fmaf(gu[i], expf(z0 + (ut[i] - nrm) + vp), acc);
        acc = fmaf(gu[i + 1], expf(z1 + (ut[i + 1] - nrm) + vp), acc);
        acc = fmaf(gu[i + 2], expf(z2 + (ut[i + 2] - nrm) + vp), acc);
        acc = fmaf(gu[i + 3], expf(z3 + (ut[i + 3] - nrm) + vp), acc);
      }
    }
    for (; i < i1; ++i)
      acc = fmaf(gu[i], expf(lotb_z(p, b, i, j, alpha) + (ut[i] - lotb_log_mu(p, b, i)) + vp), acc);
  }
  p.colpart[((size_t)b * p.nslab + blockIdx.y) * (p.M + 1) + j] = acc;
}
// gv_dst[j] = sign * sum over the slabs (fixed order: deterministic)
__global__ void __launch_bounds__(256) lotb_colreduce_kernel(const LotbParams p, int t_dst, float sign) {
  const int b = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
  if (j > p.M) return;
  float acc = 0.f;
  for (int s = 0; s < p.nslab; ++s) acc += p.colpart[((size_t)b * p.nslab + s) * (p.M + 1) + j];
  lotb_gv(p, t_dst, b)[j] = sign * acc;
}

// dL/dZ at (i, j)
__device__ __forceinline__ float lotb_gz(const LotbParams& p, int b, int i, int j, float alpha) {
  const float z = lotb_z(p, b, i, j, alpha);
  const float lmu = lotb_log_mu(p, b, i), lnu = lotb_log_nu(p, b, j);
  float g = p.G[((size_t)b * (p.N + 1) + i) * (p.M + 1) + j];
  for (int t = 1; t <= p.I; ++t) {
    const float ut = lotb_u(p, t, b)[i], vt = lotb_v(p, t, b)[j], vp = t > 1 ? lotb_v(p, t - 1, b)[j] : 0.f;
    g -= lotb_gv(p, t, b)[j] * expf(z + ut + (vt - lnu));
    g -= lotb_gu(p, t, b)[i] * expf(z + (ut - lmu) + vp);
  }
  return g;
}

// Final pass over the scores: thread = column, slab of rows; the per-column terms of every step stay in registers (I <= 4)
template <int IT>
__global__ void __launch_bounds__(256) lotb_final_kernel(const LotbParams p) {
  const int b = blockIdx.z, j = blockIdx.x * 256 + threadIdx.x;
  const int i0 = blockIdx.y * p.slab_rows, i1 = min(i0 + p.slab_rows, p.N);
  if (j >= p.M) return;
  if (IT == 0) {   // any iteration count: everything re-read per element (cached)
    const float alpha = *p.alpha;
    for (int i = i0; i < i1; ++i) p.gscores[((size_t)b * p.N + i) * p.M + j] = lotb_gz(p, b, i, j, alpha);
    return;
  }
  const float lnu = p.consts[4 * b], lmu = p.consts[4 * b];   // (i < N, j < M: both are norm)
  float gvj[IT > 0 ? IT : 1], cb[IT > 0 ? IT : 1], vp[IT > 0 ? IT : 1];
#pragma unroll
  for (int t = 1; t <= IT; ++t) {
    gvj[t - 1] = lotb_gv(p, t, b)[j];
    cb[t - 1] = lotb_v(p, t, b)[j] - lnu;
    vp[t - 1] = t > 1 ? lotb_v(p, t - 1, b)[j] : 0.f;
  }
  for (int i = i0; i < i1; ++i) {
    const float z = p.scores[((size_t)b * p.N + i) * p.M + j];
    float g = p.G[((size_t)b * (p.N + 1) + i) * (p.M + 1) + j];
#pragma unroll
    for (int t = 1; t <= IT; ++t) {
      const float ut = lotb_u(p, t, b)[i];
      g -= gvj[t - 1] * expf(z + ut + cb[t - 1]);
      g -= lotb_gu(p, t, b)[i] * expf(z + (ut - lmu) + vp[t - 1]);
    }
    p.gscores[((size_t)b * p.N + i) * p.M + j] = g;
  }
}

// dL/d alpha of one batch element: dL/dZ summed over the dustbin row and column -- one entry per thread, block sums into
// alphapart[b][block], then a fixed-order sum (a single CTA per batch element needed 72 us for the 8193 entries of a 4096^2 problem)
__global__ void __launch_bounds__(256) lotb_alpha_kernel(const LotbParams p, float* __restrict__ alphapart) {
  __shared__ float red[8];
  const int b = blockIdx.y, e = blockIdx.x * 256 + threadIdx.x;
  const float alpha = *p.alpha;
  float acc = 0.f;
  if (e <= p.M) acc = lotb_gz(p, b, p.N, e, alpha);                    // the dustbin row (incl. the corner)
  else if (e <= p.M + p.N) acc = lotb_gz(p, b, e - p.M - 1, p.M, alpha);  // the dustbin column
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    alphapart[(size_t)b * gridDim.x + blockIdx.x] = s;
  }
}
__global__ void lotb_alpha_reduce_kernel(const float* __restrict__ alphapart, int nblk, float* __restrict__ galpha) {
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < nblk; ++k) s += alphapart[(size_t)blockIdx.x * nblk + k];
    galpha[blockIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Backward of the dual-softmax branch (matching.py:147-157): conf = A * B, A = softmax over the SRC axis of sim / T with the invalid
// src rows at -inf, B = softmax over the TGT axis with the invalid tgt columns at -inf.  With P = A B, c_j = sum_i G_ij P_ij and
// r_i = sum_j G_ij P_ij:   dL/d sim_ij = (2 P_ij G_ij - A_ij c_j - B_ij r_i) / T.
// ---------------------------------------------------------------------------------------------------------------------------
struct DsbParams {
  const float* sim;       // [B, N, M]
  const uint8_t* src_mask;
  const uint8_t* tgt_mask;
  const float* G;         // [B, N, M]
  float* rstat;           // [B, N, 2] (max, sum) of the rows over the valid tgt columns
  float* cstat;           // [B, M, 2] (max, sum) of the columns over the valid src rows
  float* r;               // [B, N]
  float* c;               // [B, M]
  float* colpart;         // [B, nslab, M, 2]
  float* gsim;            // [B, N, M]
  int B, N, M, nslab;
  float inv_T;
};
__device__ __forceinline__ void dsb_merge(float& m, float& sum, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  if (mn == __int_as_float(0xff800000)) return;       // both empty
  sum = sum * expf(m - mn) + s2 * expf(m2 - mn);
  m = mn;
}
// A_ij, B_ij from the statistics
__device__ __forceinline__ void dsb_ab(const DsbParams& p, int b, int i, int j, float x, bool sv, bool tv, float& A, float& Bv) {
  const float2 cs = reinterpret_cast<const float2*>(p.cstat)[(size_t)b * p.M + j];
  const float2 rs = reinterpret_cast<const float2*>(p.rstat)[(size_t)b * p.N + i];
  A = sv ? expf(x - cs.x) / cs.y : 0.f;
  Bv = tv ? expf(x - rs.x) / rs.y : 0.f;
}
// mode 0: row statistics; mode 1: r_i = sum_j G P.  One warp per row.
__global__ void __launch_bounds__(256) dsb_row_kernel(const DsbParams p, int mode) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= p.N) return;
  const float* xr = p.sim + ((size_t)b * p.N + i) * p.M;
  const uint8_t* tm = p.tgt_mask + (size_t)b * p.M;
  if (mode == 0) {
    float m = __int_as_float(0xff800000), sum = 0.f;
    for (int j = lane; j < p.M; j += 32)
      if (tm[j]) dsb_merge(m, sum, xr[j] * p.inv_T, 1.f);
    for (int o = 16; o > 0; o >>= 1) dsb_merge(m, sum, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, sum, o));
    if (lane == 0) reinterpret_cast<float2*>(p.rstat)[(size_t)b * p.N + i] = make_float2(m, sum);
    return;
  }
  const bool sv = p.src_mask[(size_t)b * p.N + i] != 0;
  const float* gr = p.G + ((size_t)b * p.N + i) * p.M;
  float acc = 0.f;
  for (int j = lane; j < p.M; j += 32) {
    float A, Bv;
    dsb_ab(p, b, i, j, xr[j] * p.inv_T, sv, tm[j] != 0, A, Bv);
    acc = fmaf(gr[j], A * Bv, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) p.r[(size_t)b * p.N + i] = acc;
}
// mode 0: column statistics; mode 1: c_j = sum_i G P.  One thread per column over a slab of rows, then a fixed-order reduce.
__global__ void __launch_bounds__(256) dsb_col_kernel(const DsbParams p, int mode) {
  const int b = blockIdx.z, j = blockIdx.x * 256 + threadIdx.x;
  const int i0 = blockIdx.y * LOTB_SLAB, i1 = min(i0 + LOTB_SLAB, p.N);
  if (j >= p.M) return;
  const uint8_t* sm = p.src_mask + (size_t)b * p.N;
  float2* out = reinterpret_cast<float2*>(p.colpart) + ((size_t)b * p.nslab + blockIdx.y) * p.M + j;
  if (mode == 0) {
    float m = __int_as_float(0xff800000), sum = 0.f;
    for (int i = i0; i < i1; ++i)
      if (sm[i]) dsb_merge(m, sum, p.sim[((size_t)b * p.N + i) * p.M + j] * p.inv_T, 1.f);
    *out = make_float2(m, sum);
    return;
  }
  const bool tv = p.tgt_mask[(size_t)b * p.M + j] != 0;
  float acc = 0.f;
  for (int i = i0; i < i1; ++i) {
    float A, Bv;
    dsb_ab(p, b, i, j, p.sim[((size_t)b * p.N + i) * p.M + j] * p.inv_T, sm[i] != 0, tv, A, Bv);
    acc = fmaf(p.G[((size_t)b * p.N + i) * p.M + j], A * Bv, acc);
  }
  *out = make_float2(acc, 0.f);
}
__global__ void __launch_bounds__(256) dsb_colreduce_kernel(const DsbParams p, int mode) {
  const int b = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
  if (j >= p.M) return;
  const float2* part = reinterpret_cast<const float2*>(p.colpart) + (size_t)b * p.nslab * p.M + j;
  if (mode == 0) {
    float m = __int_as_float(0xff800000), sum = 0.f;
    for (int s2 = 0; s2 < p.nslab; ++s2) dsb_merge(m, sum, part[(size_t)s2 * p.M].x, part[(size_t)s2 * p.M].y);
    reinterpret_cast<float2*>(p.cstat)[(size_t)b * p.M + j] = make_float2(m, sum);
  } else {
    float acc = 0.f;
    for (int s2 = 0; s2 < p.nslab; ++s2) acc += part[(size_t)s2 * p.M].x;
    p.c[(size_t)b * p.M + j] = acc;
  }
}
__global__ void __launch_bounds__(256) dsb_final_kernel(const DsbParams p) {
  const int b = blockIdx.z, j = blockIdx.x * 256 + threadIdx.x;
  const int i0 = blockIdx.y * LOTB_SLAB, i1 = min(i0 + LOTB_SLAB, p.N);
  if (j >= p.M) return;
  const bool tv = p.tgt_mask[(size_t)b * p.M + j] != 0;
  const float cj = p.c[(size_t)b * p.M + j];
  for (int i = i0; i < i1; ++i) {
    const size_t e = ((size_t)b * p.N + i) * p.M + j;
    float A, Bv;
    dsb_ab(p, b, i, j, p.sim[e] * p.inv_T, p.src_mask[(size_t)b * p.N + i] != 0, tv, A, Bv);
    p.gsim[e] = (2.f * A * Bv * p.G[e] - A * cj - Bv * p.r[(size_t)b * p.N + i]) * p.inv_T;
  }
}

}  // namespace drg

using namespace drg;

static int lotb_nslab(int N) { return (N + 1 + LOTB_SLAB - 1) / LOTB_SLAB; }

extern "C" size_t drg_sinkhorn_backward_workspace_bytes(int B, int N, int M, int iters) {
  if (B < 1 || N < 1 || M < 1 || iters < 1) return 0;
  const size_t f = (size_t)4 * B + (size_t)iters * B * (N + 1) + (size_t)iters * B * (M + 1) + (size_t)B * lotb_nslab(N) * (M + 1) +
                   (size_t)B * ((N + M + 1 + 255) / 256);
  return align_up(f * sizeof(float), 256);
}

extern "C" int drg_sinkhorn_backward(const float* scores, const float* alpha, const uint8_t* src_mask, const uint8_t* tgt_mask, int B, int N,
                                     int M, int iters, const float* u_all, const float* v_all, const float* grad_out, float* grad_scores,
                                     float* grad_alpha, void* workspace, size_t workspace_bytes, void* stream) {
  DRG_CHECK_ARG(scores && alpha && src_mask && tgt_mask && u_all && v_all && grad_out && grad_scores && grad_alpha && workspace,
                "all pointers must be non-null");
  DRG_CHECK_ARG(B >= 1 && N >= 1 && M >= 1 && iters >= 1 && B <= 65535, "B, N, M, iters must be >= 1");
  DRG_CHECK_ARG(workspace_bytes >= drg_sinkhorn_backward_workspace_bytes(B, N, M, iters), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  LotbParams p{};
  p.scores = scores; p.alpha = alpha; p.G = grad_out; p.u_all = u_all; p.v_all = v_all;
  p.B = B; p.N = N; p.M = M; p.I = iters;
  p.nslab = lotb_nslab(N);
  p.slab_rows = LOTB_SLAB;
  float* w = reinterpret_cast<float*>(workspace);
  float* consts = w;
  p.consts = consts;
  p.gu_all = w + (size_t)4 * B;
  p.gv_all = p.gu_all + (size_t)iters * B * (N + 1);
  p.colpart = p.gv_all + (size_t)iters * B * (M + 1);
  p.gscores = grad_scores;
  p.galpha = grad_alpha;
  lotb_consts_kernel<<<B, 256, 0, st>>>(src_mask, tgt_mask, N, M, consts);
  DRG_LAUNCH_CHECK();
  const dim3 gcol((unsigned)((M + 1 + 255) / 256), (unsigned)p.nslab, (unsigned)B), gred((unsigned)((M + 1 + 255) / 256), (unsigned)B);
  const dim3 grow((unsigned)((N + 1 + 7) / 8), (unsigned)B);
  lotb_col_kernel<<<gcol, 256, 0, st>>>(p, iters, 1);          // gv_I = colsum(G)
  DRG_LAUNCH_CHECK();
  lotb_colreduce_kernel<<<gred, 256, 0, st>>>(p, iters, 1.f);
  DRG_LAUNCH_CHECK();
  for (int t = iters; t >= 1; --t) {
    lotb_row_kernel<<<grow, 256, 0, st>>>(p, t, t == iters ? 1 : 0);   // gu_t
    DRG_LAUNCH_CHECK();
    if (t > 1) {
      lotb_col_kernel<<<gcol, 256, 0, st>>>(p, t, 0);                 // gv_{t-1}
      DRG_LAUNCH_CHECK();
      lotb_colreduce_kernel<<<gred, 256, 0, st>>>(p, t - 1, -1.f);
      DRG_LAUNCH_CHECK();
    }
  }
  const dim3 gfin((unsigned)((M + 255) / 256), (unsigned)((N + LOTB_SLAB - 1) / LOTB_SLAB), (unsigned)B);
  switch (iters) {
    case 1: lotb_final_kernel<1><<<gfin, 256, 0, st>>>(p); break;
    case 2: lotb_final_kernel<2><<<gfin, 256, 0, st>>>(p); break;
    case 3: lotb_final_kernel<3><<<gfin, 256, 0, st>>>(p); break;
    case 4: lotb_final_kernel<4><<<gfin, 256, 0, st>>>(p); break;
    default: lotb_final_kernel<0><<<gfin, 256, 0, st>>>(p); break;
  }
  DRG_LAUNCH_CHECK();
  const int nblk = (N + M + 1 + 255) / 256;
  float* alphapart = p.colpart + (size_t)B * p.nslab * (M + 1);
  lotb_alpha_kernel<<<dim3((unsigned)nblk, (unsigned)B), 256, 0, st>>>(p, alphapart);
  DRG_LAUNCH_CHECK();
  lotb_alpha_reduce_kernel<<<B, 32, 0, st>>>(alphapart, nblk, p.galpha);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" size_t drg_dual_softmax_backward_workspace_bytes(int B, int N, int M) {
  if (B < 1 || N < 1 || M < 1) return 0;
  const size_t nslab = (size_t)(N + LOTB_SLAB - 1) / LOTB_SLAB;
  return align_up(((size_t)B * N * 3 + (size_t)B * M * 3 + (size_t)B * nslab * M * 2) * sizeof(float), 256);
}

extern "C" int drg_dual_softmax_backward(const float* sim, const uint8_t* src_mask, const uint8_t* tgt_mask, int B, int N, int M,
                                         float temperature, const float* grad_conf, float* grad_sim, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  DRG_CHECK_ARG(sim && src_mask && tgt_mask && grad_conf && grad_sim && workspace, "all pointers must be non-null");
  DRG_CHECK_ARG(B >= 1 && N >= 1 && M >= 1 && B <= 65535 && temperature > 0.f, "B, N, M >= 1, temperature > 0");
  DRG_CHECK_ARG(workspace_bytes >= drg_dual_softmax_backward_workspace_bytes(B, N, M), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  DsbParams p{};
  p.sim = sim; p.src_mask = src_mask; p.tgt_mask = tgt_mask; p.G = grad_conf; p.gsim = grad_sim;
  p.B = B; p.N = N; p.M = M;
  p.nslab = (N + LOTB_SLAB - 1) / LOTB_SLAB;
  p.inv_T = 1.f / temperature;
  float* w = reinterpret_cast<float*>(workspace);
  p.rstat = w;
  p.cstat = p.rstat + (size_t)2 * B * N;
  p.r = p.cstat + (size_t)2 * B * M;
  p.c = p.r + (size_t)B * N;
  p.colpart = p.c + (size_t)B * M;
  const dim3 grow((unsigned)((N + 7) / 8), (unsigned)B), gcol((unsigned)((M + 255) / 256), (unsigned)p.nslab, (unsigned)B),
      gred((unsigned)((M + 255) / 256), (unsigned)B);
  for (int mode = 0; mode < 2; ++mode) {     // statistics, then the weighted sums that need them
    dsb_row_kernel<<<grow, 256, 0, st>>>(p, mode);
    DRG_LAUNCH_CHECK();
    dsb_col_kernel<<<gcol, 256, 0, st>>>(p, mode);
    DRG_LAUNCH_CHECK();
    dsb_colreduce_kernel<<<gred, 256, 0, st>>>(p, mode);
    DRG_LAUNCH_CHECK();
  }
  dsb_final_kernel<<<gcol, 256, 0, st>>>(p);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}
