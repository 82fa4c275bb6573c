// Correspondence RANSAC on the device: the consumer of get_match's correspondences in the
// 3DMatch / 4DMatch evaluation,
//   Diff-Reg-4dmatch/models/loss.py:13-24   ransac_pose_estimation -> open3d registration_ransac_based_on_correspondence
//   Diff-Reg-4dmatch/models/loss.py:366-398 MatchMotionLoss.ransac_regist_coarse (per batch element, < 3 matches -> identity)
// Open3D (pinned 0.13.0 in eccv24_4d_env.yml, absent from this image) runs max_iteration = 50000 trials one after the
// other on the host: draw ransac_n correspondences (with replacement), fit a rigid transform to them
// (TransformationEstimationPointToPoint(False): Umeyama without scale), count the correspondences within
// max_correspondence_distance of their partner under that transform (fitness = inliers / C, inlier_rmse), keep the trial
// with the higher fitness, the lower rmse on equal fitness.  Here every trial is one THREAD: the trials are independent, the
// correspondences are staged once per CTA in shared memory (every thread of a warp reads the same correspondence: a
// broadcast), and the best trial is a fixed-order reduction -- the result is a function of (inputs, seed) only.
// The draws come from a counter-based generator (splitmix64 of (seed, batch element, trial, draw)), so the oracle draws the
// same samples; Open3D seeds a Mersenne twister from random_device, so no run of the reference is reproducible either
// (its tester repeats the evaluation "to combat ransac nondeterministic", Diff-Reg-3dmatch/lib/tester.py:25).
#include "common.cuh"

namespace drg {
namespace {

constexpr int RS_MIN_THREADS = 64;   // trials per CTA: chosen per call (ransac_block) so that the CTAs fill the SMs evenly
constexpr int RS_MAX_THREADS = 384;
constexpr int RS_TILE = 1024;     // correspondences staged per shared-memory tile (24 KB)
constexpr int RS_MAX_N = 8;       // ransac_n <= 8

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// draw j of trial h of batch element b: uniform in [0, C) (multiply-shift of the generator's high 32 bits)
__device__ __forceinline__ int draw_index(unsigned long long seed, int b, int h, int j, int C) {
  const unsigned long long ctr = ((unsigned long long)(unsigned)b << 40) ^ ((unsigned long long)(unsigned)h << 4) ^ (unsigned long long)j;
  const unsigned long long r = splitmix64(seed ^ splitmix64(ctr));
  return (int)(((r >> 32) * (unsigned long long)C) >> 32);
}

// Eigenvectors of a symmetric 3x3 matrix (cyclic Jacobi, fp64), eigenvalues descending.
__device__ void eig3_sym(double A[3][3], double V[3][3], double lam[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 24; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double dg = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-32 * dg || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double d = A[q][q] - A[p][p];
        // fp64 sqrt and division are long software sequences and 50 000 threads run ~15 rotations each: one rsqrt for the
        // root (x rsqrt(x), x > 0 here), one reciprocal, one rsqrt for the cosine
        const double rad = d * d + 4.0 * apq * apq;
        const double t = (d >= 0.0 ? 2.0 : -2.0) * apq * __drcp_rn(fabs(d) + rad * rsqrt(rad));
        const double c = rsqrt(1.0 + t * t), s = c * t;
        const int r = 3 - p - q;  // the third index
        const double arp = A[r][p], arq = A[r][q];
        A[p][p] -= t * apq;
        A[q][q] += t * apq;
        A[p][q] = A[q][p] = 0.0;
        A[r][p] = A[p][r] = c * arp - s * arq;
        A[r][q] = A[q][r] = s * arp + c * arq;
        for (int i = 0; i < 3; ++i) {
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
      }
  }
  for (int j = 0; j < 3; ++j) lam[j] = A[j][j];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = a + 1; b < 3; ++b)
      if (lam[b] > lam[a]) {
        double tmp = lam[a]; lam[a] = lam[b]; lam[b] = tmp;
        for (int i = 0; i < 3; ++i) { tmp = V[i][a]; V[i][a] = V[i][b]; V[i][b] = tmp; }
      }
}

// Rigid fit y ~ R x + t to n point pairs (Kabsch / Umeyama without scale).  With H = sum (y - my)(x - mx)^T = U S V^T the
// answer U diag(1, 1, det U det V) V^T equals u0 v0^T + u1 v1^T + (u0 x u1)(v0 x v1)^T whatever the sign of the third pair,
// so only the two leading singular pairs are needed -- three points always give a rank-2 H.  false: the sample is
// degenerate (coincident or collinear points: second singular value below 1e-7 of the first).
// Three pairs: closed form.  Both centred triangles lie in planes; with right-handed orthonormal frames (e1, e2, ne) of the source
// plane and (f1, f2, nf) of the target plane H = F h E^T with the 2 x 2 matrix h = sum beta_k alpha_k^T of the in-plane
// coordinates, so the two leading singular pairs of H are those of h and  U diag(1, 1, det U det V) V^T = F q E^T + det(q) nf ne^T
// with q the orthogonal polar factor of h: the rotation (c, -s; s, c), (c, s) ~ (h00 + h11, h10 - h01), when det h > 0, the
// reflection (a, b; b, -a), (a, b) ~ (h00 - h11, h01 + h10), when det h < 0.  No iteration, five rsqrt.
__device__ __forceinline__ bool rigid_fit3(const float* xs, const float* ys, float R[9], float t[3]) {
  double x[3][3], y[3][3], mx[3], my[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { x[k][a] = (double)xs[k * 3 + a]; y[k][a] = (double)ys[k * 3 + a]; }
    mx[a] = (x[0][a] + x[1][a] + x[2][a]) * (1.0 / 3.0);
    my[a] = (y[0][a] + y[1][a] + y[2][a]) * (1.0 / 3.0);
  }
  double fr[2][3][3];  // [source | target][e1, e2, n][xyz]
  bool ok = true;
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const double(*p)[3] = w ? y : x;
    const double d1[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
    const double d2[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
    const double n[3] = {d1[1] * d2[2] - d1[2] * d2[1], d1[2] * d2[0] - d1[0] * d2[2], d1[0] * d2[1] - d1[1] * d2[0]};
    const double l1 = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2], ln = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    ok = ok && l1 > 0.0 && ln > 0.0;
    const double i1 = rsqrt(l1), in = rsqrt(ln);
#pragma unroll
    for (int a = 0; a < 3; ++a) { fr[w][0][a] = d1[a] * i1; fr[w][2][a] = n[a] * in; }
    fr[w][1][0] = fr[w][2][1] * fr[w][0][2] - fr[w][2][2] * fr[w][0][1];  // e2 = n x e1
    fr[w][1][1] = fr[w][2][2] * fr[w][0][0] - fr[w][2][0] * fr[w][0][2];
    fr[w][1][2] = fr[w][2][0] * fr[w][0][1] - fr[w][2][1] * fr[w][0][0];
  }
  if (!ok) return false;
  double h[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double al[2], be[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      al[i] = fr[0][i][0] * (x[k][0] - mx[0]) + fr[0][i][1] * (x[k][1] - mx[1]) + fr[0][i][2] * (x[k][2] - mx[2]);
      be[i] = fr[1][i][0] * (y[k][0] - my[0]) + fr[1][i][1] * (y[k][1] - my[1]) + fr[1][i][2] * (y[k][2] - my[2]);
    }
    h[0][0] += be[0] * al[0]; h[0][1] += be[0] * al[1];
    h[1][0] += be[1] * al[0]; h[1][1] += be[1] * al[1];
  }
  const double det = h[0][0] * h[1][1] - h[0][1] * h[1][0];
  const double fro = h[0][0] * h[0][0] + h[0][1] * h[0][1] + h[1][0] * h[1][0] + h[1][1] * h[1][1];
  if (!(fabs(det) > 1e-7 * fro)) return false;  // second singular value below ~1e-7 of the first
  double q[2][2], dq;
  if (det > 0.0) {
    const double c = h[0][0] + h[1][1], sn = h[1][0] - h[0][1], r = rsqrt(c * c + sn * sn);
    q[0][0] = c * r; q[0][1] = -sn * r; q[1][0] = sn * r; q[1][1] = c * r;
    dq = 1.0;
  } else {
    const double a = h[0][0] - h[1][1], b = h[0][1] + h[1][0], r = rsqrt(a * a + b * b);
    q[0][0] = a * r; q[0][1] = b * r; q[1][0] = b * r; q[1][1] = -a * r;
    dq = -1.0;
  }
  double Rd[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      Rd[a][c] = fr[1][0][a] * (q[0][0] * fr[0][0][c] + q[0][1] * fr[0][1][c]) + fr[1][1][a] * (q[1][0] * fr[0][0][c] + q[1][1] * fr[0][1][c]) +
                 dq * fr[1][2][a] * fr[0][2][c];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    t[a] = (float)(my[a] - (Rd[a][0] * mx[0] + Rd[a][1] * mx[1] + Rd[a][2] * mx[2]));
#pragma unroll
    for (int c = 0; c < 3; ++c) R[a * 3 + c] = (float)Rd[a][c];
  }
  return true;
}

// NS: ransac_n at compile time (sample arrays in registers), 0 = run time.
template <int NS>
__device__ __forceinline__ bool rigid_fit(const float* xs, const float* ys, int n_rt, float R[9], float t[3]) {
  if (NS == 3) return rigid_fit3(xs, ys, R, t);
  const int n = NS ? NS : n_rt;
  double mx[3] = {0, 0, 0}, my[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < n; ++k)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mx[a] += (double)xs[k * 3 + a];
      my[a] += (double)ys[k * 3 + a];
    }
  const double inv = 1.0 / (double)n;
  for (int a = 0; a < 3; ++a) { mx[a] *= inv; my[a] *= inv; }
  double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int k = 0; k < n; ++k)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) H[a][c] += ((double)ys[k * 3 + a] - my[a]) * ((double)xs[k * 3 + c] - mx[c]);
  double A[3][3], V[3][3], lam[3];
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) A[a][c] = H[0][a] * H[0][c] + H[1][a] * H[1][c] + H[2][a] * H[2][c];  // H^T H
  eig3_sym(A, V, lam);
  if (!(lam[0] > 0.0) || !(lam[1] > 1e-14 * lam[0])) return false;
  double u0[3], u1[3];
  for (int a = 0; a < 3; ++a) {
    u0[a] = H[a][0] * V[0][0] + H[a][1] * V[1][0] + H[a][2] * V[2][0];
    u1[a] = H[a][0] * V[0][1] + H[a][1] * V[1][1] + H[a][2] * V[2][1];
  }
  const double n0 = rsqrt(u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]);
  for (int a = 0; a < 3; ++a) u0[a] *= n0;
  const double dot = u0[0] * u1[0] + u0[1] * u1[1] + u0[2] * u1[2];
  for (int a = 0; a < 3; ++a) u1[a] -= dot * u0[a];
  const double n1 = rsqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  for (int a = 0; a < 3; ++a) u1[a] *= n1;
  const double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
  const double v2[3] = {V[1][0] * V[2][1] - V[2][0] * V[1][1], V[2][0] * V[0][1] - V[0][0] * V[2][1],
                        V[0][0] * V[1][1] - V[1][0] * V[0][1]};
  double Rd[3][3];
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) Rd[a][c] = u0[a] * V[c][0] + u1[a] * V[c][1] + u2[a] * v2[c];
  for (int a = 0; a < 3; ++a) {
    t[a] = (float)(my[a] - (Rd[a][0] * mx[0] + Rd[a][1] * mx[1] + Rd[a][2] * mx[2]));
    for (int c = 0; c < 3; ++c) R[a * 3 + c] = (float)Rd[a][c];
  }
  return true;
}

struct RansacParams {
  const float* src;          // [B, N, 3]
  const float* tgt;          // [B, M, 3]
  const long long* match;    // [C_total, 3] rows (b, i, j), grouped by b
  const int* offsets;        // [B + 1] first row of every batch element (in the workspace: written by ransac_prep_kernel)
  unsigned int* tickets;     // [B] CTAs of the batch element that have finished (zeroed by ransac_prep_kernel)
  int B, N, M, n, T;
  float thr2;
  unsigned long long seed;
  unsigned long long* cta_best;  // [B, ctas] packed (count, ~err bits) of the CTA's best trial
  int* cta_best_h;               // [B, ctas]
  float* cta_pose;               // [B, ctas, 12] R (row-major) and t of that trial
  int* hyp_count;                // optional [B, T]
  float* hyp_err2;               // optional [B, T]
  float* pose;                   // [B, 4, 4]
  float* fitness;                // [B]
  float* rmse;                   // [B]
  int* best;                     // [B] winning trial (-1: fewer than 3 correspondences or no valid trial)
  int* inliers;                  // [B]
};

// a point number outside [0, n) in the match list (the reference would raise an IndexError on the host; a launch cannot) is
// clamped so that no read leaves the point clouds
__device__ __forceinline__ long long clamp_index(long long i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

template <int NS>
__device__ __forceinline__ bool fit_trial(const RansacParams& p, int b, int h, int c0, int C, float R[9], float t[3]) {
  float xs[(NS ? NS : RS_MAX_N) * 3], ys[(NS ? NS : RS_MAX_N) * 3];
  const int n = NS ? NS : p.n;
#pragma unroll
  for (int j = 0; j < n; ++j) {
    const long long* row = p.match + (long long)(c0 + draw_index(p.seed, b, h, j, C)) * 3;
    const float* s = p.src + ((long long)b * p.N + clamp_index(row[1], p.N)) * 3;
    const float* g = p.tgt + ((long long)b * p.M + clamp_index(row[2], p.M)) * 3;
#pragma unroll
    for (int a = 0; a < 3; ++a) { xs[j * 3 + a] = s[a]; ys[j * 3 + a] = g[a]; }
  }
  return rigid_fit<NS>(xs, ys, p.n, R, t);
}

// (count, err2) -> 64-bit key, larger is better: more inliers first, then the smaller squared-error sum (at equal count
// the rmse order is the err2 order); err2 >= 0 so its float bits are monotone.
__device__ __forceinline__ unsigned long long trial_key(int count, float err2) {
  return ((unsigned long long)(unsigned)count << 32) | (unsigned long long)(0xFFFFFFFFu - __float_as_uint(err2));
}

// grid (ctas, B), blockDim.x (a multiple of 32, <= RS_MAX_THREADS) threads: thread = trial.  Correspondences pass through shared memory in tiles.  The last CTA
// of a batch element to finish reduces the CTA bests and writes the pose.
// MAXT: the launch bound (ptxas allocates 96 registers under a bound of 128 and 80, with a slower inner loop, under 384: the
// multi-wave launches keep the small bound).
template <int NS, int MAXT>
__global__ void __launch_bounds__(MAXT) ransac_trials_kernel(RansacParams p) {
  __shared__ float sm[RS_TILE * 6];
  __shared__ unsigned long long wkey[RS_MAX_THREADS / 32];
  __shared__ int wh[RS_MAX_THREADS / 32];
  const int nthr = blockDim.x, nwarp = nthr >> 5;
  const int b = blockIdx.y;
  const int c0 = p.offsets[b], C = p.offsets[b + 1] - c0;
  const int h = blockIdx.x * nthr + threadIdx.x;
  float R[9], t[3];
  bool valid = false;
  if (C >= 3 && h < p.T) valid = fit_trial<NS>(p, b, h, c0, C, R, t);
  int count = 0;
  float err2 = 0.f;
  const float thr2 = p.thr2;
  for (int base = 0; base < C; base += RS_TILE) {
    const int len = min(RS_TILE, C - base);
    __syncthreads();
    for (int k = threadIdx.x; k < len; k += nthr) {
      const long long* row = p.match + (long long)(c0 + base + k) * 3;
      const float* s = p.src + ((long long)b * p.N + clamp_index(row[1], p.N)) * 3;
      const float* g = p.tgt + ((long long)b * p.M + clamp_index(row[2], p.M)) * 3;
      sm[k * 6 + 0] = s[0]; sm[k * 6 + 1] = s[1]; sm[k * 6 + 2] = s[2];
      sm[k * 6 + 3] = g[0]; sm[k * 6 + 4] = g[1]; sm[k * 6 + 5] = g[2];
    }
    __syncthreads();
    if (valid) {
#pragma unroll 4
      for (int k = 0; k < len; ++k) {
        const float2 a0 = *reinterpret_cast<const float2*>(&sm[k * 6]);
        const float2 a1 = *reinterpret_cast<const float2*>(&sm[k * 6 + 2]);
        const float2 a2 = *reinterpret_cast<const float2*>(&sm[k * 6 + 4]);
        const float x = a0.x, y = a0.y, z = a1.x;
        const float dx = fmaf(R[0], x, fmaf(R[1], y, fmaf(R[2], z, t[0]))) - a1.y;
        const float dy = fmaf(R[3], x, fmaf(R[4], y, fmaf(R[5], z, t[1]))) - a2.x;
        const float dz = fmaf(R[6], x, fmaf(R[7], y, fmaf(R[8], z, t[2]))) - a2.y;
        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        // open3d: dis < max_correspondence_distance.  Predicated adds written out: the compiler's own choice for the count was an
        // add plus a predicated IMAD.MOV, which sits on the FMA pipe -- the pipe this loop is bound by.
        asm("{\n\t.reg .pred q;\n\tsetp.lt.f32 q, %2, %3;\n\t@q add.s32 %0, %0, 1;\n\t@q add.f32 %1, %1, %2;\n\t}" : "+r"(count), "+f"(err2) : "f"(d2), "f"(thr2));
      }
    }
  }
  if (h < p.T && p.hyp_count) {
    p.hyp_count[(long long)b * p.T + h] = valid ? count : -1;
    p.hyp_err2[(long long)b * p.T + h] = err2;
  }
  // best trial of the CTA: larger key, then the lower trial number (the sequential loop keeps the first of equals)
  unsigned long long key = (valid && count > 0) ? trial_key(count, err2) : 0ull;
  int bh = h;
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long k2 = __shfl_xor_sync(0xFFFFFFFFu, key, o);
    const int h2 = __shfl_xor_sync(0xFFFFFFFFu, bh, o);
    if (k2 > key || (k2 == key && h2 < bh)) { key = k2; bh = h2; }
  }
  if ((threadIdx.x & 31) == 0) { wkey[threadIdx.x >> 5] = key; wh[threadIdx.x >> 5] = bh; }
  __syncthreads();
  __shared__ bool last;
  __shared__ int cta_h;
  if (threadIdx.x == 0) {
    for (int w = 1; w < nwarp; ++w)
      if (wkey[w] > key || (wkey[w] == key && wh[w] < bh)) { key = wkey[w]; bh = wh[w]; }
    p.cta_best[(long long)b * gridDim.x + blockIdx.x] = key;
    p.cta_best_h[(long long)b * gridDim.x + blockIdx.x] = bh;
    cta_h = key ? bh : -1;
  }
  __syncthreads();
  if (h == cta_h) {  // the CTA's best trial leaves its fit: the finisher copies it (a single-thread fp64 refit took 10 us)
    float* q = p.cta_pose + ((long long)b * gridDim.x + blockIdx.x) * 12;
#pragma unroll
    for (int k = 0; k < 9; ++k) q[k] = R[k];
    q[9] = t[0]; q[10] = t[1]; q[11] = t[2];
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(&p.tickets[b], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  // The last CTA of the batch element to finish picks the best CTA record (fixed order: the result does not depend on which
  // CTA is last) and publishes its pose.
  __threadfence();
  const int ctas = gridDim.x;
  key = 0ull;
  bh = 0x7FFFFFFF;
  for (int k = threadIdx.x; k < ctas; k += nthr) {
    const unsigned long long k2 = __ldcg(&p.cta_best[(long long)b * ctas + k]);
    const int h2 = __ldcg(&p.cta_best_h[(long long)b * ctas + k]);
    if (k2 > key || (k2 == key && h2 < bh)) { key = k2; bh = h2; }
  }
  // a trial belongs to exactly one CTA (h / blockDim.x), so (key, h) carries the record number along
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long k2 = __shfl_xor_sync(0xFFFFFFFFu, key, o);
    const int h2 = __shfl_xor_sync(0xFFFFFFFFu, bh, o);
    if (k2 > key || (k2 == key && h2 < bh)) { key = k2; bh = h2; }
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { wkey[threadIdx.x >> 5] = key; wh[threadIdx.x >> 5] = bh; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int w = 1; w < nwarp; ++w)
    if (wkey[w] > key || (wkey[w] == key && wh[w] < bh)) { key = wkey[w]; bh = wh[w]; }
  p.tickets[b] = 0u;
  const int cbest = (int)(key >> 32);
  const bool found = C >= 3 && cbest > 0;
  if (found) {
    const float* q = p.cta_pose + ((long long)b * ctas + bh / nthr) * 12;
    for (int k = 0; k < 9; ++k) R[k] = __ldcg(q + k);
    for (int k = 0; k < 3; ++k) t[k] = __ldcg(q + 9 + k);
  } else {  // loss.py:384-387: identity; open3d returns the identity with fitness 0 when no trial had an inlier
    const float I9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; ++k) R[k] = I9[k];
    t[0] = t[1] = t[2] = 0.f;
  }
  float* P = p.pose + (long long)b * 16;
  for (int a = 0; a < 3; ++a) {
    for (int c = 0; c < 3; ++c) P[a * 4 + c] = R[a * 3 + c];
    P[a * 4 + 3] = t[a];
  }
  P[12] = P[13] = P[14] = 0.f;
  P[15] = 1.f;
  const float ebest = __uint_as_float(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
  p.fitness[b] = found ? (float)cbest / (float)C : 0.f;
  p.rmse[b] = found ? sqrtf(ebest / (float)cbest) : 0.f;
  p.best[b] = found ? bh : -1;
  p.inliers[b] = found ? cbest : 0;
}

// one CTA: first row of every batch element in the (b, i, j) list (rows grouped by ascending b) by binary search -- or a copy
// of the caller's offsets -- and the tickets cleared.
__global__ void ransac_prep_kernel(const long long* match, long long rows, const int* offsets_in, int B, int* offsets,
                                   unsigned int* tickets) {
  for (int b = threadIdx.x; b <= B; b += blockDim.x) {
    if (offsets_in) {
      offsets[b] = offsets_in[b];
    } else {
      long long lo = 0, hi = rows;  // first row with batch number >= b
      while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (match[mid * 3] < (long long)b) lo = mid + 1; else hi = mid;
      }
      offsets[b] = (int)lo;
    }
    if (b < B) tickets[b] = 0u;
  }
}

// upper bound of the CTAs per batch element (workspace layout)
inline int ransac_ctas(int T) { return (T + RS_MIN_THREADS - 1) / RS_MIN_THREADS; }
// Threads per CTA.  The kernel is issue-bound and one trial is one thread, so what matters is that every SM gets the same
// number of warps: 50 000 trials in CTAs of 128 are 391 CTAs = 2 or 3 per SM (measured: SMs active 73 % of the kernel's
// duration), in CTAs of 352 they are 143 CTAs = one per SM.  Pick the size with the best (wave fill) x (lane fill).
inline int ransac_block(int T, int B, int regs_per_thread) {
  int best = 128;
  double best_eff = -1.0;
  // several waves of CTAs: the block scheduler evens the SMs out by itself and small CTAs leave the shorter tail (measured
  // with 8 x 50 000 trials: 0.70 ms in CTAs of 128, 0.80 ms with the size this search picks)
  if ((long long)((T + 127) / 128) * B > (long long)NUM_SMS * (65536 / (regs_per_thread * 128))) return 128;
  for (int tpb = RS_MIN_THREADS; tpb <= RS_MAX_THREADS; tpb += 32) {
    const long long ctas = (long long)((T + tpb - 1) / tpb) * B;
    int resident = 65536 / (regs_per_thread * tpb);
    if (resident > 8) resident = 8;   // 24.6 KB of static shared memory per CTA
    if (resident < 1) continue;
    const double waves = (double)ctas / (double)(NUM_SMS * resident);
    double full = waves;
    if (full != (double)(long long)full) full = (double)((long long)full + 1);
    const double eff = waves / full * (double)T / (double)(((T + tpb - 1) / tpb) * (long long)tpb);
    if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && tpb > best)) {  // ties: the larger CTA stages the list fewer times
      best_eff = eff > best_eff ? eff : best_eff;
      best = tpb;
    }
  }
  return best;
}

}  // namespace
}  // namespace drg

using namespace drg;

extern "C" size_t drg_ransac_workspace_bytes(int B, int max_iteration) {
  if (B < 1 || max_iteration < 1) return 0;
  const size_t ctas = (size_t)ransac_ctas(max_iteration);
  return align_up((size_t)B * ctas * sizeof(unsigned long long), 256) + align_up((size_t)B * ctas * sizeof(int), 256) +
         align_up((size_t)B * ctas * 12 * sizeof(float), 256) + align_up((size_t)(B + 1) * sizeof(int), 256) +
         align_up((size_t)B * sizeof(unsigned int), 256);
}

extern "C" int drg_ransac_correspondence(const float* src, const float* tgt, int B, int N, int M, const long long* match,
                                         long long num_match, const int* offsets, float max_correspondence_distance, int ransac_n,
                                         int max_iteration, unsigned long long seed, float* pose, float* fitness,
                                         float* inlier_rmse, int* best_trial, int* inlier_count, int* trial_count,
                                         float* trial_err2, void* workspace, size_t workspace_bytes, void* stream) {
  DRG_CHECK_ARG(src && tgt && pose && fitness && inlier_rmse && best_trial && inlier_count && workspace,
                "src/tgt/outputs/workspace must be non-null");
  DRG_CHECK_ARG(B >= 1 && N >= 1 && M >= 1, "B, N, M must be >= 1");
  DRG_CHECK_ARG(B <= 65535, "B must be <= 65535");
  DRG_CHECK_ARG(num_match >= 0 && num_match < (1ll << 31) && (match || num_match == 0), "match must hold num_match (< 2^31) rows");
  DRG_CHECK_ARG(ransac_n >= 3 && ransac_n <= RS_MAX_N, "ransac_n must be in [3, 8]");
  DRG_CHECK_ARG(max_iteration >= 1 && max_iteration <= (1 << 27), "max_iteration must be in [1, 2^27]");
  DRG_CHECK_ARG(max_correspondence_distance > 0.f, "max_correspondence_distance must be > 0");
  DRG_CHECK_ARG((trial_count == nullptr) == (trial_err2 == nullptr), "trial_count and trial_err2 go together");
  if (workspace_bytes < drg_ransac_workspace_bytes(B, max_iteration)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, drg_ransac_workspace_bytes(B, max_iteration));
    return DRG_ERR_WORKSPACE;
  }
  const int ctas = ransac_ctas(max_iteration);
  char* w = (char*)workspace;
  RansacParams p;
  p.cta_best = (unsigned long long*)w;
  w += align_up((size_t)B * ctas * sizeof(unsigned long long), 256);
  p.cta_best_h = (int*)w;
  w += align_up((size_t)B * ctas * sizeof(int), 256);
  p.cta_pose = (float*)w;
  w += align_up((size_t)B * ctas * 12 * sizeof(float), 256);
  int* offs = (int*)w;
  w += align_up((size_t)(B + 1) * sizeof(int), 256);
  p.tickets = (unsigned int*)w;
  p.src = src; p.tgt = tgt; p.match = match; p.offsets = offs;
  p.B = B; p.N = N; p.M = M; p.n = ransac_n; p.T = max_iteration;
  p.thr2 = max_correspondence_distance * max_correspondence_distance;
  p.seed = seed;
  p.hyp_count = trial_count; p.hyp_err2 = trial_err2;
  p.pose = pose; p.fitness = fitness; p.rmse = inlier_rmse; p.best = best_trial; p.inliers = inlier_count;
  cudaStream_t st = (cudaStream_t)stream;
  ransac_prep_kernel<<<1, 256, 0, st>>>(match, num_match, offsets, B, offs, p.tickets);
  DRG_LAUNCH_CHECK();
  const int tpb = ransac_block(max_iteration, B, 96);
  const dim3 grid((max_iteration + tpb - 1) / tpb, B);
  if (tpb <= 128) {
    if (ransac_n == 3)
      ransac_trials_kernel<3, 128><<<grid, tpb, 0, st>>>(p);
    else if (ransac_n == 4)
      ransac_trials_kernel<4, 128><<<grid, tpb, 0, st>>>(p);
    else
      ransac_trials_kernel<0, 128><<<grid, tpb, 0, st>>>(p);
  } else {
    if (ransac_n == 3)
      ransac_trials_kernel<3, RS_MAX_THREADS><<<grid, tpb, 0, st>>>(p);
    else if (ransac_n == 4)
      ransac_trials_kernel<4, RS_MAX_THREADS><<<grid, tpb, 0, st>>>(p);
    else
      ransac_trials_kernel<0, RS_MAX_THREADS><<<grid, tpb, 0, st>>>(p);
  }
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}
