// Backward pass of the weighted Kabsch solve (training path, SURVEY.md section 8f rank 3: the reference's motion loss
// differentiates R, t of SoftProcrustesLayer.batch_weighted_procrustes, Diff-Reg-4dmatch/models/procrustes.py:18-44, with respect to the
// correspondence weights -- through its host SVD).  sm_100a.
//
// Forward: W1 = sum |w|, wt = w / (W1 + eps), muX = sum wt X, muY = sum wt Y, M = sum wt (Y - muY)(X - muX)^T = U D V^T,
// R = U diag(1, 1, det U det V) V^T, t = muY - R muX.  No SVD is needed backwards: H = R^T M is symmetric, and a perturbation
// dM turns R by R [omega]x with (tr(H) I - H) omega = vee(R^T dM - dM^T R); hence, with gR' = gR - gt muX^T and
// e = (tr(H) I - H)^-1 vee(skew(R^T gR')):
//     dL/dM = 2 R [e]x,   dL/d muY = gt - dL/dM b,   dL/d muX = -R^T gt - dL/dM^T a      (a = (1 - sum wt) muY, b = (1 - sum wt) muX)
//     dL/d wt_k = (Y_k - muY)^T dL/dM (X_k - muX) + X_k . dL/d muX + Y_k . dL/d muY
//     dL/d w_k  = dL/d wt_k / (W1 + eps) - sign(w_k) sum_j (dL/d wt_j) w_j / (W1 + eps)^2
// Checked against torch's autograd through the unmodified reference function (oracle.weighted_procrustes_backward, golden
// ``procrb_*``).  One CTA per batch element, fp64 sums in a fixed order (deterministic).
#include "common.cuh"

namespace drg {

constexpr int PB_THREADS = 256;

// block-wide sum of NV doubles per thread, fixed order; result broadcast in out[]
template <int NV>
__device__ void pb_block_sum(double (&v)[NV], double* sm, double (&out)[NV]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sm[warp * NV + k] = x;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = 0.0;
    for (int w = 0; w < PB_THREADS / 32; ++w) x += sm[w * NV + k];
    out[k] = x;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(PB_THREADS) procr_wbackward_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                                     const float* __restrict__ w, const float* __restrict__ R,
                                                                     const float* __restrict__ gR, const float* __restrict__ gt, int K, float eps,
                                                                     float* __restrict__ gw) {
  __shared__ double sm[(PB_THREADS / 32) * 17];
  __shared__ double bc[32];     // gM (9), gmuX (3), gmuY (3), muX (3), muY (3), denom
  const int b = blockIdx.x;
  const float* Xb = X + (size_t)b * K * 3;
  const float* Yb = Y + (size_t)b * K * 3;
  const float* wb = w + (size_t)b * K;
  float* gwb = gw + (size_t)b * K;
  // ---- sums: W1, sum w, sum w X, sum w Y, sum w Y X^T
  double v[17];
#pragma unroll
  for (int k = 0; k < 17; ++k) v[k] = 0.0;
  for (int k = threadIdx.x; k < K; k += PB_THREADS) {
    const double wk = wb[k];
    const double x0 = Xb[3 * k], x1 = Xb[3 * k + 1], x2 = Xb[3 * k + 2], y0 = Yb[3 * k], y1 = Yb[3 * k + 1], y2 = Yb[3 * k + 2];
    v[0] += fabs(wk);
    v[1] += wk;
    v[2] += wk * x0; v[3] += wk * x1; v[4] += wk * x2;
    v[5] += wk * y0; v[6] += wk * y1; v[7] += wk * y2;
    v[8] += wk * y0 * x0; v[9] += wk * y0 * x1; v[10] += wk * y0 * x2;
    v[11] += wk * y1 * x0; v[12] += wk * y1 * x1; v[13] += wk * y1 * x2;
    v[14] += wk * y2 * x0; v[15] += wk * y2 * x1; v[16] += wk * y2 * x2;
  }
  double s[17];
  pb_block_sum<17>(v, sm, s);
  if (threadIdx.x == 0) {
    const double den = s[0] + (double)eps, sw = s[1] / den;
    double muX[3] = {s[2] / den, s[3] / den, s[4] / den}, muY[3] = {s[5] / den, s[6] / den, s[7] / den};
    double M[3][3], Rm[3][3], G[3][3], g_t[3];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) {
        M[a][c] = s[8 + 3 * a + c] / den - (2.0 - sw) * muY[a] * muX[c];
        Rm[a][c] = R[(size_t)b * 9 + 3 * a + c];
      }
    for (int a = 0; a < 3; ++a) g_t[a] = gt[(size_t)b * 3 + a];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) G[a][c] = (double)gR[(size_t)b * 9 + 3 * a + c] - g_t[a] * muX[c];      // gR' = gR - gt muX^T
    // H = R^T M (symmetric), Kmat = tr(H) I - H, C = R^T gR', c = vee(skew(C))
    double H[3][3], C[3][3];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) {
        double h = 0.0, cc = 0.0;
        for (int k = 0; k < 3; ++k) {
          h += Rm[k][a] * M[k][c];
          cc += Rm[k][a] * G[k][c];
        }
        H[a][c] = h;
        C[a][c] = cc;
      }
    const double tr = H[0][0] + H[1][1] + H[2][2];
    double Km[3][3];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) Km[a][c] = (a == c ? tr : 0.0) - 0.5 * (H[a][c] + H[c][a]);
    const double cv[3] = {0.5 * (C[2][1] - C[1][2]), 0.5 * (C[0][2] - C[2][0]), 0.5 * (C[1][0] - C[0][1])};
    // e = Km^-1 cv (adjugate)
    const double c00 = Km[1][1] * Km[2][2] - Km[1][2] * Km[2][1], c01 = Km[1][2] * Km[2][0] - Km[1][0] * Km[2][2],
                 c02 = Km[1][0] * Km[2][1] - Km[1][1] * Km[2][0];
    const double det = Km[0][0] * c00 + Km[0][1] * c01 + Km[0][2] * c02;
    const double inv[3][3] = {{c00 / det, (Km[0][2] * Km[2][1] - Km[0][1] * Km[2][2]) / det, (Km[0][1] * Km[1][2] - Km[0][2] * Km[1][1]) / det},
                              {c01 / det, (Km[0][0] * Km[2][2] - Km[0][2] * Km[2][0]) / det, (Km[0][2] * Km[1][0] - Km[0][0] * Km[1][2]) / det},
                              {c02 / det, (Km[0][1] * Km[2][0] - Km[0][0] * Km[2][1]) / det, (Km[0][0] * Km[1][1] - Km[0][1] * Km[1][0]) / det}};
    double e[3];
    for (int a = 0; a < 3; ++a) e[a] = inv[a][0] * cv[0] + inv[a][1] * cv[1] + inv[a][2] * cv[2];
    const double ex[3][3] = {{0.0, -e[2], e[1]}, {e[2], 0.0, -e[0]}, {-e[1], e[0], 0.0}};
    double gM[3][3];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) gM[a][c] = 2.0 * (Rm[a][0] * ex[0][c] + Rm[a][1] * ex[1][c] + Rm[a][2] * ex[2][c]);
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) bc[3 * a + c] = gM[a][c];
    for (int c = 0; c < 3; ++c) {
      double gx = 0.0, gy = g_t[c];
      for (int k = 0; k < 3; ++k) {
        gx -= Rm[k][c] * g_t[k];                              // -R^T gt
        gx -= gM[k][c] * (1.0 - sw) * muY[k];                 // -gM^T a
        gy -= gM[c][k] * (1.0 - sw) * muX[k];                 // -gM b
      }
      bc[9 + c] = gx;
      bc[12 + c] = gy;
      bc[15 + c] = muX[c];
      bc[18 + c] = muY[c];
    }
    bc[21] = den;
  }
  __syncthreads();
  // ---- dL/d wt_k, and sum_k (dL/d wt_k) w_k
  double part[1] = {0.0};
  for (int k = threadIdx.x; k < K; k += PB_THREADS) {
    const double x[3] = {Xb[3 * k], Xb[3 * k + 1], Xb[3 * k + 2]}, y[3] = {Yb[3 * k], Yb[3 * k + 1], Yb[3 * k + 2]};
    double g = 0.0;
    for (int a = 0; a < 3; ++a) {
      double row = 0.0;
      for (int c = 0; c < 3; ++c) row += bc[3 * a + c] * (x[c] - bc[15 + c]);
      g += (y[a] - bc[18 + a]) * row + x[a] * bc[9 + a] + y[a] * bc[12 + a];
    }
    gwb[k] = (float)g;
    part[0] += g * (double)wb[k];
  }
  double tot[1];
  pb_block_sum<1>(part, sm, tot);
  const double den = bc[21];
  for (int k = threadIdx.x; k < K; k += PB_THREADS) {
    const float wk = wb[k];
    const double sgn = wk > 0.f ? 1.0 : (wk < 0.f ? -1.0 : 0.0);
    gwb[k] = (float)((double)gwb[k] / den - sgn * tot[0] / (den * den));
  }
}

}  // namespace drg

using namespace drg;

extern "C" int drg_weighted_procrustes_backward(const float* X, const float* Y, const float* w, const float* R, const float* grad_R,
                                                const float* grad_t, int B, int K, float eps, float* grad_w, void* stream) {
  DRG_CHECK_ARG(X && Y && w && R && grad_R && grad_t && grad_w, "all pointers must be non-null");
  DRG_CHECK_ARG(B >= 1 && K >= 1, "B, K must be >= 1");
  procr_wbackward_kernel<<<B, PB_THREADS, 0, (cudaStream_t)stream>>>(X, Y, w, R, grad_R, grad_t, K, eps, grad_w);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}
