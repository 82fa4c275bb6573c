// Elementwise / row kernels of the geometry attention layer (the denoising transformer's building block):
//   masked, scaled row softmax of the attention logits, written straight as the split operand of the P.V GEMM
//     replaces   a.masked_fill_(q_mask & ~kv_mask, -inf); a = a / sqrt(d); a = softmax(a, dim=2)
//                Diff-Reg-4dmatch/models/transformer.py:80-84
//   LayerNorm (+ optional residual)
//     replaces   self.norm1(message) / x + self.norm2(message)          transformer.py:88,92-94
// The matrix products of the layer (q / k / v / merge / MLP projections, Q.K^T, P.V) run on the tcgen05 GEMM of gemm.cu
// with the fp16 split operands of features.cu; nothing here needs the tensor cores.  sm_100a.
#include "common.cuh"

namespace drg {

constexpr int ATT_THREADS = 256;

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < ATT_THREADS / 32; ++w) r = fmaxf(r, red[w]);
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < ATT_THREADS / 32; ++w) r += red[w];   // fixed order: deterministic
  return r;
}

// One CTA per row (b, h, l) of the logits [B*H, L, S].  The logit of a VALID query against an INVALID key is -inf (the
// reference's mask expression q_mask * ~kv_mask: rows of invalid queries are left alone), then scaled by 1 / sqrt(d), then
// softmax over the keys.  A valid query with no valid key yields NaN, as in the reference.
//   P   (optional) fp32 probabilities [B*H, L, S] (may alias the logits)
//   P16 (optional) the same row as the LEFT split operand of the P.V GEMM: [lo | hi | tail], row pitch 2 kc + 8 (features.cu)
__global__ void __launch_bounds__(ATT_THREADS) attn_softmax_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ q_mask,
                                                                   const uint8_t* __restrict__ kv_mask, int B, int H, int L, int S,
                                                                   float scale, float* P, unsigned short* __restrict__ P16) {
  __shared__ float red[ATT_THREADS / 32];
  const long long rows = (long long)B * H * L;
  const int kc = split16_kc(S), pitch = split16_pitch(S);
  const float scale2 = scale * LOG2E;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const long long bh = row / L;
    const int l = (int)(row - bh * L);
    const int b = (int)(bh / H);
    const float* a = logits + row * S;
    const bool mask_keys = kv_mask != nullptr && (q_mask == nullptr || q_mask[(size_t)b * L + l]);
    const uint8_t* km = kv_mask ? kv_mask + (size_t)b * S : nullptr;
    // pass 1: maximum of the scaled, masked logits (log2 domain)
    float m = -INFINITY;
    for (int s = threadIdx.x; s < S; s += ATT_THREADS) {
      float x = a[s];
      if (mask_keys && !km[s]) x = -INFINITY;
      m = fmaxf(m, x * scale2);
    }
    m = block_reduce_max(m, red);
    // pass 2: sum of the exponentials
    float sum = 0.f;
    for (int s = threadIdx.x; s < S; s += ATT_THREADS) {
      float x = a[s];
      if (mask_keys && !km[s]) x = -INFINITY;
      sum += ex2(x * scale2 - m);            // all keys masked: -inf - -inf = NaN, like the reference
    }
    sum = block_reduce_sum(sum, red);
    const float inv_sum = 1.f / sum;
    // the row maximum of P is 1 / sum: the split operand's power-of-two scale puts it into [2^14, 2^15)
    int e = 0;
    if (inv_sum > 0.f && inv_sum <= 3.0e38f) e = min(max(14 - ilogbf(inv_sum), -126), 126);
    const float sc = __int_as_float((e + 127) << 23);
    unsigned short* o = P16 ? P16 + row * pitch : nullptr;
    float ss = 0.f;
    for (int s = threadIdx.x; s < kc; s += ATT_THREADS) {
      float pr = 0.f;
      if (s < S) {
        float x = a[s];
        if (mask_keys && !km[s]) x = -INFINITY;
        pr = ex2(x * scale2 - m) * inv_sum;
        if (P) P[row * S + s] = pr;
      }
      if (o) {
        const float y = pr * sc;
        ss = fmaf(y, y, ss);
        unsigned short h, lo;
        asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(y));
        float hf;
        asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h));
        asm("cvt.rn.f16.f32 %0, %1;" : "=h"(lo) : "f"(y - hf));
        o[s] = lo;          // pattern 0 (left operand): [lo | hi]
        o[kc + s] = h;
      }
    }
    if (o) {
      ss = block_reduce_sum(ss, red);
      if (threadIdx.x == 0) {
        const float inv = __int_as_float((127 - e) << 23);
        *reinterpret_cast<float4*>(o + 2 * kc) = make_float4(inv, sqrtf(ss) * inv, 0.f, 0.f);
      }
    }
    __syncthreads();   // `red` is reused by the next row
  }
}

// The same for rows of up to 4096 keys with S % 4 == 0 (16-byte aligned rows): the row is read ONCE into registers (four
// quads per thread), and the operand halves leave as 8-byte stores -- the three-pass kernel above reads a 16 KB row three
// times with 4-byte accesses (measured 194 us per [4, 4096, 4096] call against ~90 us for its bytes).
__global__ void __launch_bounds__(ATT_THREADS) attn_softmax_vec_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ q_mask,
                                                                       const uint8_t* __restrict__ kv_mask, int B, int H, int L, int S,
                                                                       float scale, float* P, unsigned short* __restrict__ P16) {
  __shared__ float red[ATT_THREADS / 32];
  constexpr int QPT = 4;   // quads per thread: 256 threads x 4 x 4 = 4096 keys
  const long long rows = (long long)B * H * L;
  const int kc = split16_kc(S), pitch = split16_pitch(S);
  const int S4 = S >> 2, kc4 = kc >> 2;
  const float scale2 = scale * LOG2E;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const long long bh = row / L;
    const int l = (int)(row - bh * L);
    const int b = (int)(bh / H);
    const float4* a4 = reinterpret_cast<const float4*>(logits + row * S);
    const bool mask_keys = kv_mask != nullptr && (q_mask == nullptr || q_mask[(size_t)b * L + l]);
    const uint8_t* km = kv_mask ? kv_mask + (size_t)b * S : nullptr;
    float x[QPT][4];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int q = threadIdx.x + ATT_THREADS * j;
      if (q < S4) {
        const float4 v = a4[q];
        x[j][0] = v.x * scale2; x[j][1] = v.y * scale2; x[j][2] = v.z * scale2; x[j][3] = v.w * scale2;
        if (mask_keys) {
          const uchar4 k4 = *reinterpret_cast<const uchar4*>(km + 4 * q);   // S % 4 == 0: mask rows are 4-byte aligned
          if (!k4.x) x[j][0] = -INFINITY;
          if (!k4.y) x[j][1] = -INFINITY;
          if (!k4.z) x[j][2] = -INFINITY;
          if (!k4.w) x[j][3] = -INFINITY;
        }
        m = fmaxf(m, fmaxf(fmaxf(x[j][0], x[j][1]), fmaxf(x[j][2], x[j][3])));
      } else {
        x[j][0] = x[j][1] = x[j][2] = x[j][3] = -INFINITY;
      }
    }
    m = block_reduce_max(m, red);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      if (threadIdx.x + ATT_THREADS * j < S4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          x[j][e] = ex2(x[j][e] - m);          // all keys masked: -inf - -inf = NaN, like the reference
          sum += x[j][e];
        }
      }
    }
    sum = block_reduce_sum(sum, red);
    const float inv_sum = 1.f / sum;
    int e2 = 0;
    if (inv_sum > 0.f && inv_sum <= 3.0e38f) e2 = min(max(14 - ilogbf(inv_sum), -126), 126);
    const float sc = __int_as_float((e2 + 127) << 23);
    unsigned short* o = P16 ? P16 + row * pitch : nullptr;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int q = threadIdx.x + ATT_THREADS * j;
      if (q < kc4) {
        float pr[4] = {0.f, 0.f, 0.f, 0.f};
        if (q < S4) {
#pragma unroll
          for (int e = 0; e < 4; ++e) pr[e] = x[j][e] * inv_sum;
          if (P) *reinterpret_cast<float4*>(P + row * S + 4 * q) = make_float4(pr[0], pr[1], pr[2], pr[3]);
        }
        if (o) {
          unsigned short h[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float y = pr[e] * sc;
            ss = fmaf(y, y, ss);
            asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h[e]) : "f"(y));
            float hf;
            asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h[e]));
            asm("cvt.rn.f16.f32 %0, %1;" : "=h"(lo[e]) : "f"(y - hf));
          }
          *reinterpret_cast<uint2*>(o + 4 * q) = make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 16), (uint32_t)lo[2] | ((uint32_t)lo[3] << 16));
          *reinterpret_cast<uint2*>(o + kc + 4 * q) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        }
      }
    }
    if (o) {
      ss = block_reduce_sum(ss, red);
      if (threadIdx.x == 0) {
        const float inv = __int_as_float((127 - e2) << 23);
        *reinterpret_cast<float4*>(o + 2 * kc) = make_float4(inv, sqrtf(ss) * inv, 0.f, 0.f);
      }
    }
    __syncthreads();   // `red` is reused by the next row
  }
}

// LayerNorm over the last dimension, one warp per row: y = (x - mean) / sqrt(var + eps) * w + b (biased variance, two passes
// like torch).  With a residual: out = residual + LayerNorm(in) (pre_add == 0: the geometry attention layer's x + norm2(message))
// or out = LayerNorm(in + residual) (pre_add == 1: the post-norm of vision3d's attention / feed-forward blocks).
// split16 (optional): the same rows as the LEFT split operand of the next linear ([lo | hi | tail], features.cu's format: row scale
// 2^e with the maximum in [2^14, 2^15), fp16 hi / lo halves, zero padding up to whole 64-column k-steps, tail = 1 / scale and the
// row norm), so that the consumer needs no staging launch.  The row stays in registers between the maximum and the conversion
// (C <= 32 * LN_MAXV).
constexpr int LN_MAXV = 36;   // rows of up to 1152 columns (kc(C) <= 32 * LN_MAXV)
// x = hi + lo, both fp16 (x is already row-scaled into fp16's range): hi = fp16(x), lo = fp16(x - hi)   (as features.cu)
__device__ __forceinline__ void ln_split16(float x, unsigned short& hi_bits, unsigned short& lo_bits) {
  unsigned short h;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  float hf;
  asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h));
  unsigned short l;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(l) : "f"(x - hf));
  hi_bits = h;
  lo_bits = l;
}
template <int NV>      // NV == 0: no split output; else ceil(kc(C) / 32) <= NV values of a row per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                        const float* __restrict__ bias, const float* __restrict__ residual,
                                                        long long rows, int C, float eps, int pre_add, float* __restrict__ out,
                                                        unsigned short* __restrict__ split_out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool pre = pre_add && residual != nullptr;
  const int Kp = split16_kc(C), pitch = split16_pitch(C);
  for (long long row = warp0; row < rows; row += nwarps) {
    const float* x = in + row * C;
    const float* r = residual ? residual + row * C : nullptr;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += pre ? x[c] + r[c] : x[c];
    const float mean = warp_sum(s) / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = (pre ? x[c] + r[c] : x[c]) - mean;
      v = fmaf(d, d, v);
    }
    const float rstd = rsqrtf(warp_sum(v) / (float)C + eps);
    constexpr bool SPLIT = NV > 0;
    float yv[SPLIT ? NV : 1];
    float amax = 0.f;
    bool weird = false;
    if (!SPLIT) {
      for (int c = lane; c < C; c += 32) {
        float y = ((pre ? x[c] + r[c] : x[c]) - mean) * rstd;
        if (w) y *= w[c];
        if (bias) y += bias[c];
        if (r && !pre) y += r[c];
        out[row * C + c] = y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < (SPLIT ? NV : 1); ++k) {
        const int c = lane + 32 * k;
        yv[k] = 0.f;
        if (c < C) {
          float y = ((pre ? x[c] + r[c] : x[c]) - mean) * rstd;
          if (w) y *= w[c];
          if (bias) y += bias[c];
          if (r && !pre) y += r[c];
          out[row * C + c] = y;
          yv[k] = y;
          const float m = fabsf(y);
          weird |= !(m <= 3.0e38f);
          amax = fmaxf(amax, m);
        }
      }
    }
    if (SPLIT) {
      amax = warp_max(amax);
      weird = __any_sync(0xffffffffu, weird);
      int e = 0;
      if (amax > 0.f && !weird) e = min(max(14 - ilogbf(amax), -126), 126);
      const float sc = __int_as_float((e + 127) << 23), inv = __int_as_float((127 - e) << 23);
      unsigned short* o = split_out + row * pitch;
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < (SPLIT ? NV : 1); ++k) {
        const int c = lane + 32 * k;
        if (c < Kp) {
          unsigned short hb = 0, lb = 0;
          if (c < C) {
            const float ys = yv[k] * sc;
            ss = fmaf(ys, ys, ss);
            ln_split16(ys, hb, lb);
          }
          o[c] = lb;             // pattern 0: [lo | hi | tail]
          o[Kp + c] = hb;
        }
      }
      ss = warp_sum(ss);
      if (lane == 0) *reinterpret_cast<float4*>(o + 2 * Kp) = make_float4(inv, sqrtf(ss) * inv, 0.f, 0.f);
    }
  }
}

// Fourier embedding of coordinates (vision3d FourierEmbedding with use_input: the 2D-3D fusion module's positional term):
//   out[row] = [x (n) | for l < L: sin(f_l x) (n), cos(f_l x) (n)],  f_l = 2^(k0 + l) [* pi]
//   replaces FourierEmbedding.forward   Diff-Reg-2d3d/vision3d/layers/embedding.py:75-99; `center` (optional, [n]) is subtracted
//   first (CrossModalFusionModule.create_3d_embedding: points - points.mean(dim=1), fusion_module.py:57)
__global__ void __launch_bounds__(256) fourier_embed_kernel(const float* __restrict__ x, const float* __restrict__ center, long long rows,
                                                            int n, int L, float k0, int use_pi, int use_input, float* __restrict__ out) {
  const int width = (use_input ? n : 0) + 2 * L * n;
  const long long total = rows * width;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / width;
    int c = (int)(idx - row * width);
    float val;
    if (use_input && c < n) {
      val = x[row * n + c] - (center ? center[c] : 0.f);
    } else {
      if (use_input) c -= n;
      const int l = c / (2 * n);
      const int r = c - l * 2 * n;
      const int comp = r < n ? r : r - n;
      const float e = k0 + (float)l;               // integer exponents (every shipped use): an exact power of two, as
      float f = e == rintf(e) ? ldexpf(1.f, (int)e) : exp2f(e);   // 2.0 ** arange is in fp32; exp2f alone is only good to 2 ulp
      if (use_pi) f *= 3.14159265358979323846f;
      const float th = f * (x[row * n + comp] - (center ? center[comp] : 0.f));
      val = r < n ? sinf(th) : cosf(th);
    }
    out[idx] = val;
  }
}

}  // namespace drg

using namespace drg;

extern "C" int drg_attn_softmax(const float* logits, const uint8_t* q_mask, const uint8_t* kv_mask, int B, int H, int L, int S,
                                float scale, float* P, void* P16, void* stream) {
  DRG_CHECK_ARG(logits != nullptr && (P != nullptr || P16 != nullptr), "logits and at least one output must be non-null");
  DRG_CHECK_ARG(B >= 1 && H >= 1 && L >= 1 && S >= 1, "B, H, L, S must be >= 1");
  DRG_CHECK_ARG(P16 == nullptr || (((uintptr_t)P16) & 15u) == 0, "P16 must be 16-byte aligned");
  const long long rows = (long long)B * H * L;
  const int grid = (int)(rows < (long long)NUM_SMS * 8 ? rows : (long long)NUM_SMS * 8);
  const bool vec = S % 4 == 0 && split16_kc(S) <= 4096 && (((uintptr_t)logits) & 15u) == 0 && (!P || (((uintptr_t)P) & 15u) == 0) &&
                   (!kv_mask || (((uintptr_t)kv_mask) & 3u) == 0);
  if (vec)
    attn_softmax_vec_kernel<<<grid, ATT_THREADS, 0, (cudaStream_t)stream>>>(logits, q_mask, kv_mask, B, H, L, S, scale, P,
                                                                            reinterpret_cast<unsigned short*>(P16));
  else
    attn_softmax_kernel<<<grid, ATT_THREADS, 0, (cudaStream_t)stream>>>(logits, q_mask, kv_mask, B, H, L, S, scale, P,
                                                                        reinterpret_cast<unsigned short*>(P16));
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_layernorm(const float* in, const float* weight, const float* bias, const float* residual, int pre_add, long long rows,
                             int C, float eps, float* out, void* split16_out, void* stream) {
  DRG_CHECK_ARG(in != nullptr && out != nullptr, "in / out must be non-null");
  DRG_CHECK_ARG(rows >= 1 && C >= 1, "rows and C must be >= 1");
  long long blocks = (rows * 32 + 255) / 256;
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  if (split16_out) {
    if (split16_kc(C) > 32 * LN_MAXV || (((uintptr_t)split16_out) & 15u)) {
      set_error("layernorm: the split output takes rows of up to %d columns (got %d), 16-byte aligned", 32 * LN_MAXV, C);
      return DRG_ERR_UNSUPPORTED;
    }
    unsigned short* so = reinterpret_cast<unsigned short*>(split16_out);
    const int nv = (split16_kc(C) + 31) / 32;
    if (nv <= 8) layernorm_kernel<8><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, weight, bias, residual, rows, C, eps, pre_add, out, so);
    else if (nv <= 18) layernorm_kernel<18><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, weight, bias, residual, rows, C, eps, pre_add, out, so);
    else layernorm_kernel<LN_MAXV><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, weight, bias, residual, rows, C, eps, pre_add, out, so);
  } else {
    layernorm_kernel<0><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, weight, bias, residual, rows, C, eps, pre_add, out, nullptr);
  }
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_fourier_embed(const float* x, const float* center, long long rows, int n, int length, float k0, int use_pi,
                                 int use_input, float* out, void* stream) {
  DRG_CHECK_ARG(x != nullptr && out != nullptr, "x / out must be non-null");
  DRG_CHECK_ARG(rows >= 1 && n >= 1 && length >= 1, "rows, n, length must be >= 1");
  const long long total = rows * ((use_input ? n : 0) + 2ll * length * n);
  long long blocks = (total + 255) / 256;
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  fourier_embed_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, center, rows, n, length, k0, use_pi, use_input, out);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}
