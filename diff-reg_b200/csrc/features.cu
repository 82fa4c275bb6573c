// Operand preparation for the similarity GEMM: positional embedding, 1/sqrt(C) scaling and the
// 16-bit hi/lo split that turns the kind::f16 tensor-core GEMM into an fp32-accurate product.
//
// Replaces the elementwise lines of Matching.forward between the projection and the einsum:
//   VolPE.embed_pos / embed_rotary   Diff-Reg-4dmatch/models/position_encoding.py:26-46
//   feat / feat.shape[-1] ** .5      Diff-Reg-4dmatch/models/matching.py:144-145
//
// Layout: in [rows, K] fp32 row-major.  out: [rows, K] fp32 (plain) or the SPLIT operand, 16-bit, row pitch
// 2 * Kp + 8 (Kp = K rounded up to 64):  [ seg0 (Kp) | seg1 (Kp) | tail (4 floats) ]
//   pattern 0 (left operand):  seg0 = lo, seg1 = hi        pattern 1 (right operand): seg0 = hi, seg1 = lo
//   x * 2^e = hi + lo, hi = fp16(x 2^e), lo = fp16(x 2^e - hi), e per ROW such that the row maximum lies in [2^14, 2^15);
//   tail = (2^-e, ||row||_2, 0, 0).  A'.B'^T = lo.hi + hi.lo + hi.hi, rescaled per row and column in the GEMM epilogue.
#include "common.cuh"

namespace drg {

__device__ __forceinline__ float to_tf32_rna(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return __uint_as_float(y);
}

struct PrepParams {
  const float* in;   // [rows, K]
  const float* in2;  // optional second source: rows >= rows1 come from in2[row - rows1] (two tensors, one launch)
  long long rows1;
  int pattern2;      // split pattern of the rows taken from in2
  const float* pe;   // rotary: [rows, K, 2] (cos, sin); sinusoidal: [rows, K]; or NULL
  float* embedded;   // optional [rows, K]: features after the positional embedding, before scaling
  void* out;         // split: [rows, 2 * Kp + 8] 16-bit, Kp = K rounded up to 64 (see the file header); else [rows, K] fp32
  long long rows;
  int K;
  int pe_type;       // 0 none, 1 rotary, 2 sinusoidal (additive); 3 / 4: the same two, the code computed here from xyz
  const float* xyz;       // pe_type 3 / 4: [rows, 3] point coordinates (the position code never exists in HBM)
  const float* div_term;  // [K / 6]
  float ox, oy, oz, voxel;
  int split;         // 1: 16-bit split layout (row-scaled fp16 hi + fp16 lo), 0: plain scaled copy
  int pattern;       // 0: [lo | hi] (left operand)   1: [hi | lo] (right operand)
  float scale;
  int relu;          // 1: max(x, 0) before the scaling (the ReLU between the two linears of the attention layer's MLP)
  int heads, seq;    // heads > 1: the input rows are (b, l, h) -- [B, seq, heads, K] -- and the output rows (b, h, l): the per-head
                     //   operands of the attention GEMMs leave the staging kernel already head-major
};

// x' = x * 2^e (the row's power-of-two scale) as hi + lo, both fp16: hi = fp16(x'), lo = fp16(x' - hi).  With the row maximum
// scaled into [2^14, 2^15) nothing overflows and lo stays a normal fp16 number for every entry within ~2^17 of the row
// maximum (22 significant bits, what 3xTF32 kept); smaller entries keep an absolute error of 2^-25 * 2^-e -- 2^-39 of the
// row maximum.  (The tensor core does not take fp16 x bf16 products -- kind::f16 wants one format for A and B; measured:
// illegal instruction, tools/probes/umma_fmt_probe.cu -- so a bf16 lo with fp32's exponent range is not an option.)
__device__ __forceinline__ void prep_split16(float x, unsigned short& hi_bits, unsigned short& lo_bits) {
  unsigned short h, l;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  float hf;
  asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h));
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(l) : "f"(x - hf));   // x - hf is exact in fp32
  hi_bits = h;
  lo_bits = l;
}

// positional embedding + scaling of one quad of a row (the elementwise lines between the projection and the einsum)
// PEK: which position-code branches are compiled in -- 0: none, 1: a code tensor (pe_type 1 / 2), 2: computed from xyz
// (pe_type 3 / 4; sinf / cosf with their slow paths: kept out of the other instantiations, whose unrolled loops would
// otherwise be instruction-fetch bound -- measured: 53 % of the warp samples of the split kernel were "no instruction")
template <int PEK>
__device__ __forceinline__ float4 prep_load_quad(const PrepParams& p, long long row, int k, bool store_embedded) {
  const bool second = p.in2 != nullptr && row >= p.rows1;
  float4 x = second ? *reinterpret_cast<const float4*>(p.in2 + (row - p.rows1) * p.K + k)
                    : *reinterpret_cast<const float4*>(p.in + row * p.K + k);
  if (PEK == 1 && p.pe_type == 1) {
    // x*cos + rot(x)*sin with rot(x)[2i] = -x[2i+1], rot(x)[2i+1] = x[2i]; same op order as the reference
    const float4 cs0 = *reinterpret_cast<const float4*>(p.pe + (row * p.K + k) * 2);      // cos0 sin0 cos1 sin1
    const float4 cs1 = *reinterpret_cast<const float4*>(p.pe + (row * p.K + k) * 2 + 4);  // cos2 sin2 cos3 sin3
    float4 y;
    y.x = __fadd_rn(__fmul_rn(x.x, cs0.x), __fmul_rn(-x.y, cs0.y));
    y.y = __fadd_rn(__fmul_rn(x.y, cs0.z), __fmul_rn(x.x, cs0.w));
    y.z = __fadd_rn(__fmul_rn(x.z, cs1.x), __fmul_rn(-x.w, cs1.y));
    y.w = __fadd_rn(__fmul_rn(x.w, cs1.z), __fmul_rn(x.z, cs1.w));
    x = y;
  } else if (PEK == 1 && p.pe_type == 2) {
    const float4 pe = *reinterpret_cast<const float4*>(p.pe + row * p.K + k);
    x.x += pe.x;
    x.y += pe.y;
    x.z += pe.z;
    x.w += pe.w;
  } else if (PEK == 2 && (p.pe_type == 3 || p.pe_type == 4)) {
    // VolumetricPositionEncoding.forward fused into the operand staging (position_encoding.py:49-87 + :26-46): the
    // angles, sinf / cosf and the embedding arithmetic are those of position_code_kernel + the branches above, so the
    // result is bit-identical to going through the [rows, K, 2] / [rows, K] code tensor -- which is never written.
    const int d3 = p.K / 3, d6 = p.K / 6;
    const float px = p.xyz[row * 3 + 0], py = p.xyz[row * 3 + 1], pz = p.xyz[row * 3 + 2];
    float xv[4] = {x.x, x.y, x.z, x.w};
    if (p.pe_type == 3) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {          // the two (even, odd) pairs of the quad share an angle each
        const int c = k + 2 * h;
        const int axis = c / d3;
        const int kk = (c - axis * d3) >> 1;
        const float vox = ((axis == 0 ? px : axis == 1 ? py : pz) - (axis == 0 ? p.ox : axis == 1 ? p.oy : p.oz)) / p.voxel;
        const float ang = vox * p.div_term[kk];
        const float cs = cosf(ang), sn = sinf(ang);
        const float a = xv[2 * h], b2 = xv[2 * h + 1];
        xv[2 * h] = __fadd_rn(__fmul_rn(a, cs), __fmul_rn(-b2, sn));
        xv[2 * h + 1] = __fadd_rn(__fmul_rn(b2, cs), __fmul_rn(a, sn));
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = k + e;
        const int seg = c / d6;
        const int axis = seg >> 1;
        const float vox = ((axis == 0 ? px : axis == 1 ? py : pz) - (axis == 0 ? p.ox : axis == 1 ? p.oy : p.oz)) / p.voxel;
        const float ang = vox * p.div_term[c - seg * d6];
        xv[e] += (seg & 1) ? cosf(ang) : sinf(ang);
      }
    }
    x = make_float4(xv[0], xv[1], xv[2], xv[3]);
  }
  if (store_embedded && p.embedded) *reinterpret_cast<float4*>(p.embedded + row * p.K + k) = x;
  if (p.relu) {
    x.x = fmaxf(x.x, 0.f);
    x.y = fmaxf(x.y, 0.f);
    x.z = fmaxf(x.z, 0.f);
    x.w = fmaxf(x.w, 0.f);
  }
  x.x *= p.scale;
  x.y *= p.scale;
  x.z *= p.scale;
  x.w *= p.scale;
  return x;
}

// input row (b, l, h) -> output row (b, h, l) when the rows are per-head slices of [B, seq, heads, K]
__device__ __forceinline__ long long prep_out_row(const PrepParams& p, long long row) {
  if (p.heads <= 1) return row;
  const long long bl = row / p.heads;
  const int h = (int)(row - bl * p.heads);
  const long long b = bl / p.seq;
  const long long l = bl - b * p.seq;
  return (b * p.heads + h) * p.seq + l;
}

// plain scaled copy (split == 0): one quad per thread
template <int PEK>
__global__ void __launch_bounds__(256) prep_operand_kernel(const PrepParams p) {
  const int K4 = p.K >> 2;
  const long long total = p.rows * K4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / K4;
    const int k = (int)(idx - row * K4) << 2;
    const float4 x = prep_load_quad<PEK>(p, row, k, true);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + prep_out_row(p, row) * p.K + k) = x;
  }
}

// 16-bit split operand (split == 1): ONE WARP PER ROW, because the row's power-of-two scale comes from its maximum.
// The row (<= PREP_MAXQ quads per lane, K <= 1024) stays in registers between the maximum and the conversion; wider rows
// are evaluated twice.
constexpr int PREP_MAXQ = 8;
template <int PEK>
__global__ void __launch_bounds__(256) prep_split_kernel(const PrepParams p) {
  const int Kp = (p.K + 63) & ~63;   // each segment is padded to whole 64-column k-steps (zeros)
  const int pitch = split16_pitch(p.K);
  const int K4 = p.K >> 2, Kp4 = Kp >> 2;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool in_regs = K4 <= 32 * PREP_MAXQ;
  for (long long row = warp0; row < p.rows; row += nwarps) {
    float4 xq[PREP_MAXQ];
    float amax = 0.f;
    bool weird = false;   // Inf / NaN in the row: no scaling, the conversion propagates them into the product
    auto see = [&](const float4& x) {
      const float m = fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w)));
      weird |= !(m <= 3.0e38f);
      amax = fmaxf(amax, m);
    };
    if (in_regs) {
#pragma unroll
      for (int j = 0; j < PREP_MAXQ; ++j) {
        const int q = j * 32 + lane;
        if (q < K4) {
          xq[j] = prep_load_quad<PEK>(p, row, q << 2, true);
          see(xq[j]);
        }
      }
    } else {
      for (int q = lane; q < K4; q += 32) see(prep_load_quad<PEK>(p, row, q << 2, true));
    }
    amax = warp_max(amax);
    weird = __any_sync(0xffffffffu, weird);
    int e = 0;
    if (amax > 0.f && !weird) e = min(max(14 - ilogbf(amax), -126), 126);   // amax * 2^e in [2^14, 2^15)
    const float sc = __int_as_float((e + 127) << 23), inv = __int_as_float((127 - e) << 23);
    unsigned short* o = reinterpret_cast<unsigned short*>(p.out) + prep_out_row(p, row) * pitch;
    const int pat = (p.in2 != nullptr && row >= p.rows1) ? p.pattern2 : p.pattern;
    float ss = 0.f;
    auto emit = [&](int q, float4 x) {   // quad q of the padded row; x is ignored in the padding
      const int k = q << 2;
      uint2 hv = make_uint2(0u, 0u), lv = make_uint2(0u, 0u);
      if (q < K4) {
        x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
        ss += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
        unsigned short h[4], l[4];
        prep_split16(x.x, h[0], l[0]);
        prep_split16(x.y, h[1], l[1]);
        prep_split16(x.z, h[2], l[2]);
        prep_split16(x.w, h[3], l[3]);
        hv = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        lv = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
      }
      // the GEMM issues the small cross terms (lo . hi, hi . lo) before hi . hi within every k-step
      *reinterpret_cast<uint2*>(o + k) = pat == 0 ? lv : hv;
      *reinterpret_cast<uint2*>(o + Kp + k) = pat == 0 ? hv : lv;
    };
    if (in_regs) {
#pragma unroll
      for (int j = 0; j < PREP_MAXQ; ++j) {
        const int q = j * 32 + lane;
        if (q < Kp4) emit(q, q < K4 ? xq[j] : make_float4(0.f, 0.f, 0.f, 0.f));
      }
      for (int q = PREP_MAXQ * 32 + lane; q < Kp4; q += 32) emit(q, make_float4(0.f, 0.f, 0.f, 0.f));   // (padding only)
    } else {
      for (int q = lane; q < Kp4; q += 32) emit(q, q < K4 ? prep_load_quad<PEK>(p, row, q << 2, false) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    ss = warp_sum(ss);
    // row tail: 1 / scale (what the GEMM epilogue multiplies back) and the row's Euclidean norm (true values: the bound the
    // projection's split epilogue derives ITS rows' scales from)
    if (lane == 0) *reinterpret_cast<float4*>(o + 2 * Kp) = make_float4(inv, sqrtf(ss) * inv, 0.f, 0.f);
  }
}

}  // namespace drg

using namespace drg;

static int prep_run(const float* in, const float* in2, long long rows1, const float* pe, int pe_type, long long rows, int K,
                    float scale, int split, int pattern, int pattern2, float* embedded, void* out, void* stream,
                    const float* xyz = nullptr, const float* div_term = nullptr, const float* origin3 = nullptr, float voxel = 1.f,
                    int relu = 0, int heads = 1, int seq = 1) {
  DRG_CHECK_ARG(in && out, "in/out must be non-null");
  DRG_CHECK_ARG(rows >= 1 && K >= 4, "rows >= 1 and K >= 4 required");
  DRG_CHECK_ARG(pe_type >= 0 && pe_type <= 4, "pe_type must be 0 (none), 1 (rotary), 2 (sinusoidal), 3 / 4 (the same from xyz)");
  DRG_CHECK_ARG(pe_type == 0 || pe_type >= 3 || pe != nullptr, "pe is null");
  DRG_CHECK_ARG(pe_type < 3 || (xyz && div_term && origin3 && voxel != 0.f && K % 6 == 0),
                "position code from xyz needs xyz / div_term / origin, a non-zero voxel size and K % 6 == 0");
  if (K % 4 != 0 || ((uintptr_t)in & 15u) || ((uintptr_t)out & 15u) || (pe && ((uintptr_t)pe & 15u)) ||
      (embedded && ((uintptr_t)embedded & 15u))) {
    set_error("prep_operand: K must be a multiple of 4 and all pointers 16-byte aligned (K=%d)", K);
    return DRG_ERR_UNSUPPORTED;
  }
  PrepParams p{};
  p.in = in;
  p.in2 = in2;
  p.rows1 = rows1;
  p.pattern2 = pattern2;
  p.pe = pe;
  p.embedded = embedded;
  p.out = out;
  p.rows = rows;
  p.K = K;
  p.pe_type = pe_type;
  p.xyz = xyz;
  p.div_term = div_term;
  if (origin3) {
    p.ox = origin3[0];
    p.oy = origin3[1];
    p.oz = origin3[2];
  }
  p.voxel = voxel;
  p.split = split;
  p.pattern = pattern;
  p.scale = scale;
  p.relu = relu;
  p.heads = heads;
  p.seq = seq;
  const long long total = split ? rows * 32 : rows * (K / 4);   // split: one warp per row
  long long blocks = (total + 255) / 256;
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  {
    ProfScope prof_scope(PROF_PREP_OPERAND, (cudaStream_t)stream);
    const int pek = pe_type == 0 ? 0 : pe_type <= 2 ? 1 : 2;
    cudaStream_t st = (cudaStream_t)stream;
    if (split) {
      if (pek == 0) prep_split_kernel<0><<<(int)blocks, 256, 0, st>>>(p);
      else if (pek == 1) prep_split_kernel<1><<<(int)blocks, 256, 0, st>>>(p);
      else prep_split_kernel<2><<<(int)blocks, 256, 0, st>>>(p);
    } else {
      if (pek == 0) prep_operand_kernel<0><<<(int)blocks, 256, 0, st>>>(p);
      else if (pek == 1) prep_operand_kernel<1><<<(int)blocks, 256, 0, st>>>(p);
      else prep_operand_kernel<2><<<(int)blocks, 256, 0, st>>>(p);
    }
  }
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_prep_operand(const float* in, const float* pe, int pe_type, long long rows, int K, float scale, int split,
                                int pattern, float* embedded, void* out, void* stream) {
  return prep_run(in, nullptr, rows, pe, pe_type, rows, K, scale, split, pattern, pattern, embedded, out, stream);
}

extern "C" int drg_prep_operand_ext(const float* in, const float* pe, int pe_type, long long rows, int K, float scale, int split,
                                    int pattern, int relu, int heads, int seq, float* embedded, void* out, void* stream) {
  DRG_CHECK_ARG(heads >= 1 && seq >= 1, "heads and seq must be >= 1");
  DRG_CHECK_ARG(heads == 1 || rows % ((long long)heads * seq) == 0, "rows must be a multiple of heads * seq");
  return prep_run(in, nullptr, rows, pe, pe_type, rows, K, scale, split, pattern, pattern, embedded, out, stream, nullptr, nullptr,
                  nullptr, 1.f, relu, heads, seq);
}

extern "C" int drg_prep_operand_xyz(const float* in, const float* xyz, const float* div_term, const float* origin3, float voxel_size,
                                    int pe_type, long long rows, int K, float scale, int split, int pattern, float* embedded,
                                    void* out, void* stream) {
  DRG_CHECK_ARG(pe_type == 1 || pe_type == 2, "pe_type must be 1 (rotary) or 2 (sinusoidal)");
  return prep_run(in, nullptr, rows, nullptr, pe_type + 2, rows, K, scale, split, pattern, pattern, embedded, out, stream, xyz,
                  div_term, origin3, voxel_size);
}

extern "C" int drg_prep_operand_pair(const float* in_a, long long rows_a, int pattern_a, const float* in_b, long long rows_b,
                                     int pattern_b, int K, float scale, int split, void* out, void* stream) {
  DRG_CHECK_ARG(in_b != nullptr && rows_b >= 1 && (((uintptr_t)in_b) & 15u) == 0, "second input must be non-null, non-empty, 16-byte aligned");
  return prep_run(in_a, in_b, rows_a, nullptr, 0, rows_a + rows_b, K, scale, split, pattern_a, pattern_b, nullptr, out, stream);
}

// ---- small elementwise / reduction helpers of the sampler -----------------------------------
namespace drg {
// conf_matrix_pred = sigmoid(x)                     Diff-Reg-4dmatch/models/pipeline.py:192
__global__ void __launch_bounds__(256) sigmoid_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = 1.f / (1.f + expf(-x[i]));
}
// global minimum (x.min() of the 3DMatch sampler)   Diff-Reg-3dmatch/models/pipeline.py:239,264
__global__ void __launch_bounds__(256) min_kernel(const float* __restrict__ x, size_t n, unsigned int* __restrict__ out_ordered) {
  float m = INFINITY;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fminf(m, x[i]);
  m = warp_min(m);
  if ((threadIdx.x & 31) == 0) atomicMin(out_ordered, float_to_ordered(m));
}
__global__ void min_init_kernel(unsigned int* o) { *o = 0xFFFFFFFFu; }
__global__ void min_finish_kernel(const unsigned int* o, float* out) { *out = ordered_to_float(*o); }
}  // namespace drg

namespace drg {
__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) { *c += inc; }
}  // namespace drg
// ---- volumetric position code
//   VolumetricPositionEncoding.forward  Diff-Reg-4dmatch/models/position_encoding.py:49-87: vox = (xyz - origin) / voxel_size
//   (:16-24), angle = vox[axis] * div_term[k]; rotary: [P, d, 2] = (cos, sin) with every angle duplicated over a feature
//   pair, axes x | y | z over thirds of d; sinusoidal: [P, d] = cat(sinx, cosx, siny, cosy, sinz, cosz).
//   div_term is passed in (d/6 values computed once by the host module exactly as the reference does) so that the
//   angles are bit-identical; sinf / cosf are the accurate library versions (the build does not use fast-math).
__global__ void __launch_bounds__(256) position_code_kernel(const float* __restrict__ xyz, const float* __restrict__ div_term,
                                                            long long points, int d, float ox, float oy, float oz, float voxel,
                                                            int pe_type, float* __restrict__ out) {
  const long long total = points * d;
  const int d6 = d / 6, d3 = d / 3;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long pt = e / d;
    const int c = (int)(e - pt * d);
    int axis, k;
    bool want_cos = false;
    if (pe_type == 1) {  // rotary
      axis = c / d3;
      k = (c - axis * d3) >> 1;
    } else {             // sinusoidal
      const int seg = c / d6;
      axis = seg >> 1;
      want_cos = (seg & 1) != 0;
      k = c - seg * d6;
    }
    const float origin = axis == 0 ? ox : axis == 1 ? oy : oz;
    const float vox = (xyz[pt * 3 + axis] - origin) / voxel;
    const float ang = vox * div_term[k];
    if (pe_type == 1) {
      reinterpret_cast<float2*>(out)[e] = make_float2(cosf(ang), sinf(ang));
    } else {
      out[e] = want_cos ? cosf(ang) : sinf(ang);
    }
  }
}

extern "C" int drg_position_code(const float* xyz, const float* div_term, long long points, int feature_dim, const float* origin3,
                                 float voxel_size, int pe_type, float* out, void* stream) {
  DRG_CHECK_ARG(xyz && div_term && origin3 && out, "xyz / div_term / origin / out must be non-null");
  DRG_CHECK_ARG(points >= 1 && feature_dim >= 6 && feature_dim % 6 == 0, "points >= 1 and feature_dim a positive multiple of 6");
  DRG_CHECK_ARG(pe_type == 1 || pe_type == 2, "pe_type must be 1 (rotary) or 2 (sinusoidal)");
  DRG_CHECK_ARG(voxel_size != 0.f, "voxel_size must be non-zero");
  DRG_CHECK_ARG(pe_type != 1 || (((uintptr_t)out) & 7u) == 0, "rotary output must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = points * feature_dim;
  long long blocks = (total + 255) / 256;
  if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
  position_code_kernel<<<(int)blocks, 256, 0, st>>>(xyz, div_term, points, feature_dim, origin3[0], origin3[1], origin3[2],
                                                     voxel_size, pe_type, out);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_counter_add(unsigned long long* counter, unsigned long long inc, void* stream) {
  DRG_CHECK_ARG(counter != nullptr, "counter is null");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, inc);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_sigmoid(const float* x, float* y, long long n, void* stream) {
  DRG_CHECK_ARG(x && y && n >= 1, "x/y must be non-null and n >= 1");
  long long blocks = (n + 255) / 256;
  if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
  sigmoid_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, (size_t)n);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_min_value(const float* x, long long n, float* out, unsigned int* scratch, void* stream) {
  DRG_CHECK_ARG(x && out && scratch && n >= 1, "x/out/scratch must be non-null and n >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  min_init_kernel<<<1, 1, 0, st>>>(scratch);
  DRG_LAUNCH_CHECK();
  long long blocks = (n + 255) / 256;
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  min_kernel<<<(int)blocks, 256, 0, st>>>(x, (size_t)n, scratch);
  DRG_LAUNCH_CHECK();
  min_finish_kernel<<<1, 1, 0, st>>>(scratch, out);
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}
