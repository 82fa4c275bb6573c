// Fused multi-head attention on the 5th-gen tensor cores: softmax(Q K^T * scale + mask) V in ONE kernel, the [L, S] attention
// matrix never leaves the SM.  sm_100a.
//
// Replaces, inside GeometryAttentionLayer.forward (Diff-Reg-4dmatch/models/transformer.py:79-85) and vision3d's
// MultiHeadAttention.forward (Diff-Reg-2d3d/vision3d/layers/transformer.py:127-154):
//     a = einsum("nlhd,nshd->nlsh", q, k); a.masked_fill_(...); a = softmax(a / sqrt(d)); o = einsum("nlsh,nshd->nlhd", a, v)
// which the materialised path of this library runs as three kernels (Q.K^T GEMM -> HBM, softmax -> HBM, P.V GEMM <- HBM:
// 3 x 268 MB per layer call at 4 heads x 4096 x 4096).  See include/diffreg_b200.h (drg_attention_split16).
//
// fp32 parity: the reference computes both products in fp32 (TF32 off), so both run here as three-term fp16 split products
// (gemm.cu): operands are the 16-bit split rows of features.cu -- Q' = [Q_lo | Q_hi | tail], K' = [K_hi | K_lo | tail] per head,
// V'^T = [V_hi | V_lo | tail] per head with the KEYS along the row -- and the probabilities are split into fp16 hi / lo halves
// on the fly.  S = Q_lo.K_hi + Q_hi.K_lo + Q_hi.K_hi, O += P_lo.V_hi + P_hi.V_lo + P_hi.V_hi, fp32 accumulation in TMEM.
//
// One CTA = 128 queries of one (batch, head); keys in tiles of 64.  160 threads:
//   warp 4 (one elected lane)  TMA producer and MMA issuer: Q once; per key tile the K tile -> S = Q.K^T into one of TWO TMEM
//                              accumulators (S of tile t+1 is computed while the softmax warps work on tile t), the V^T tile,
//                              and -- once the softmax warps have published P -- O += P.V (O lives in TMEM for the whole pass)
//   warps 0..3                 softmax: thread = query row = TMEM lane.  tcgen05.ld of the row's 64 logits, row / column scales of
//                              the split operands, masks, running maximum, exp2, row sum, fp16 hi / lo split of P written as the
//                              128-byte-swizzled K-major A operand of the P.V product
// Online softmax with a LAZY reference: P = 2^(s - m_ref + 6) with m_ref only moved (and O, l rescaled through tcgen05.ld / st)
// when the tile's maximum exceeds it by more than 8 (log2 units), so P <= 2^14 stays inside fp16 and the rescale of the
// accumulator is rare after the first tiles.  The 2^6 and the stale reference cancel in O / l.
#include "umma.cuh"

namespace drg {

constexpr int FA_BM = 128;        // queries per CTA (TMEM lanes)
constexpr int FA_BN = 64;         // keys per tile = one 128-byte swizzle-atom row of 16-bit P
constexpr int FA_THREADS = 160;
constexpr float FA_TAU = 8.f;     // move the reference when a tile's maximum exceeds it by more than this (log2 units)
constexpr float FA_PSHIFT = 6.f;  // P is carried as 2^6 * exp(.): hi / lo halves of the small entries stay normal fp16 numbers
constexpr int FA_Q_CHUNK = FA_BM * 128;   // bytes of one 64-column chunk of the Q tile
constexpr int FA_K_CHUNK = FA_BN * 128;
constexpr int FA_P_BYTES = FA_BM * 128;   // one half (hi or lo) of the P tile
constexpr int FA_S_COLS = 2 * FA_BN;      // two S accumulators
constexpr uint32_t FA_O_COL = FA_S_COLS;  // O accumulator: TMEM columns [128, 128 + ND)

struct FlashShape {
  int B, H, L, S, d;
  int kc;        // 16-bit columns of one segment of the Q / K operand rows (d rounded up to 64)
  int kcS;       // ... of the V^T operand rows (S rounded up to 64)
  int ND;        // d rounded up to 16: N of the P.V MMA (rows of the V^T tile; rows >= d are TMA zero fill)
  float scale2;  // softmax scale * log2(e)
  const unsigned short *Q16, *K16, *V16;
  const uint8_t *q_mask, *kv_mask;
  float* out;    // [B, L, H * d]
};

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_warps_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int TMEM_COLS>
__global__ void __launch_bounds__(FA_THREADS, 1)
    flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const FlashShape s) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t bar_q, bar_k, bar_v, bar_p, bar_pv;
  __shared__ __align__(8) uint64_t bar_s[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float colinfo[2][3][FA_BN];   // per tile parity: 1 / scale of the key rows, bias (0 / -inf) without and with the key mask
  __shared__ float vscale[256];            // 1 / scale of the V^T rows (= output channels)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, b = bh / s.H, h = bh - b * s.H;
  const int q0 = blockIdx.x * FA_BM;
  const int nq = 2 * s.kc / 64;            // 64-column chunks of a Q / K operand row (both segments)
  const int nh = s.kc / 64;                // ... of one segment
  const int T = (s.S + FA_BN - 1) / FA_BN;

  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* aligned = smem_dyn + (base - smem_u32(smem_dyn));
  uint8_t* sQ = aligned;
  uint8_t* sK = sQ + (size_t)nq * FA_Q_CHUNK;
  uint8_t* sV = sK + (size_t)nq * FA_K_CHUNK;
  uint8_t* sP = sV + (size_t)2 * s.ND * 128;   // [hi | lo]

  if (tid == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(&bar_q, 1);
    mbar_init(&bar_k, 1);
    mbar_init(&bar_v, 1);
    mbar_init(&bar_p, FA_BM);
    mbar_init(&bar_pv, 1);
    mbar_init(&bar_s[0], 1);
    mbar_init(&bar_s[1], 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<TMEM_COLS>(&tmem_base_slot);
  const size_t pitch_d = (size_t)2 * s.kc + 8, pitch_S = (size_t)2 * s.kcS + 8;
  for (int c = tid; c < 256; c += FA_THREADS)
    vscale[c] = c < s.d ? reinterpret_cast<const float*>(s.V16 + ((size_t)bh * s.d + c) * pitch_S + 2 * s.kcS)[0] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 4) {
    if (lane == 0) {
      // ===================== TMA producer + MMA issuer =====================
      const uint32_t idesc_s = make_idesc_f16(FA_BM, FA_BN, FMT_F16, FMT_F16);
      const uint32_t idesc_o = make_idesc_f16(FA_BM, s.ND, FMT_F16, FMT_F16);
      const uint32_t v_bytes = (uint32_t)s.ND * 128u;
      auto load_k = [&](int t) {
        mbar_arrive_expect_tx(&bar_k, (uint32_t)nq * FA_K_CHUNK);
        for (int c = 0; c < nq; ++c) tma_load_3d(sK + (size_t)c * FA_K_CHUNK, &tmK, c * 64, t * FA_BN, bh, &bar_k);
      };
      auto load_v = [&](int t) {
        mbar_arrive_expect_tx(&bar_v, 2u * v_bytes);
        tma_load_3d(sV, &tmV, t * FA_BN, 0, bh, &bar_v);                     // V_hi: keys of this tile along the row
        tma_load_3d(sV + v_bytes, &tmV, s.kcS + t * FA_BN, 0, bh, &bar_v);   // V_lo
      };
      auto issue_s = [&](int t) {
        const uint32_t d_tmem = tmem_base + (uint32_t)((t & 1) * FA_BN);
        uint32_t acc = 0u;
        // the small terms first (the accumulator is rounded at every step): chunk c of Q' against chunk c of K' is Q_lo.K_hi
        // for the first segment and Q_hi.K_lo for the second; then Q_hi.K_hi
        for (int c = 0; c < nq; ++c) {
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sQ + (size_t)c * FA_Q_CHUNK));
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(sK + (size_t)c * FA_K_CHUNK));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_f16(d_tmem, a_desc + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc_s, acc);
            acc = 1u;
          }
        }
        for (int c = 0; c < nh; ++c) {
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sQ + (size_t)(nh + c) * FA_Q_CHUNK));
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(sK + (size_t)c * FA_K_CHUNK));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, a_desc + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc_s, 1u);
        }
        umma_commit(&bar_s[t & 1]);
      };
      mbar_arrive_expect_tx(&bar_q, (uint32_t)nq * FA_Q_CHUNK);
      for (int c = 0; c < nq; ++c) tma_load_3d(sQ + (size_t)c * FA_Q_CHUNK, &tmQ, c * 64, q0, bh, &bar_q);
      load_k(0);
      load_v(0);
      mbar_wait(&bar_q, 0u);
      mbar_wait(&bar_k, 0u);
      tcgen05_fence_after();
      issue_s(0);
      const uint64_t ph_desc = make_smem_desc_sw128(smem_u32(sP));
      const uint64_t pl_desc = make_smem_desc_sw128(smem_u32(sP + FA_P_BYTES));
      const uint64_t vh_desc = make_smem_desc_sw128(smem_u32(sV));
      const uint64_t vl_desc = make_smem_desc_sw128(smem_u32(sV + v_bytes));
      const uint32_t o_tmem = tmem_base + FA_O_COL;
      for (int t = 0; t < T; ++t) {
        mbar_wait(&bar_s[t & 1], (uint32_t)((t >> 1) & 1));   // S(t) complete: the K tile may be replaced
        if (t + 1 < T) {
          load_k(t + 1);
          mbar_wait(&bar_k, (uint32_t)((t + 1) & 1));
          tcgen05_fence_after();
          issue_s(t + 1);            // into the other accumulator, whose tile t - 1 the softmax warps have consumed (bar_p of t - 1)
        }
        mbar_wait(&bar_v, (uint32_t)(t & 1));
        mbar_wait(&bar_p, (uint32_t)(t & 1));
        tcgen05_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t o = (uint64_t)(2 * kk);
          umma_f16(o_tmem, pl_desc + o, vh_desc + o, idesc_o, (uint32_t)((t | kk) != 0));   // P_lo . V_hi
          umma_f16(o_tmem, ph_desc + o, vl_desc + o, idesc_o, 1u);                          // P_hi . V_lo
          umma_f16(o_tmem, ph_desc + o, vh_desc + o, idesc_o, 1u);                          // P_hi . V_hi
        }
        umma_commit(&bar_pv);
        if (t + 1 < T) {
          mbar_wait(&bar_pv, (uint32_t)(t & 1));   // the V tile (and P) have been read
          load_v(t + 1);
        }
      }
    }
  } else {
    // ===================== softmax warps: thread = query row =====================
    const int row = tid, q = q0 + row;
    const bool row_ok = q < s.L;
    const float iq = row_ok ? reinterpret_cast<const float*>(s.Q16 + ((size_t)bh * s.L + q) * pitch_d + 2 * s.kc)[0] * s.scale2 : 0.f;
    // keys are masked for valid queries only (transformer.py:80-81); no query mask = every query valid (vision3d's k_masks)
    const bool use_mask = s.kv_mask != nullptr && (s.q_mask == nullptr || (row_ok && s.q_mask[(size_t)b * s.L + q] != 0));
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float NEG_INF = __int_as_float(0xff800000);
    float m_ref = NEG_INF, l = 0.f;
    uint8_t* p_row = sP + (size_t)(row >> 3) * 1024 + (size_t)(row & 7) * 128;
    // the key rows' scales and mask bytes of tile t are fetched one tile ahead (scattered 4-byte loads: their latency stays off
    // the per-tile critical path)
    float n_ik = 0.f;
    bool n_ok = false, n_kv = false;
    auto fetch_cols = [&](int t) {
      if (tid < FA_BN && t < T) {
        const int j = t * FA_BN + tid;
        n_ok = j < s.S;
        n_ik = n_ok ? reinterpret_cast<const float*>(s.K16 + ((size_t)bh * s.S + j) * pitch_d + 2 * s.kc)[0] : 0.f;
        n_kv = n_ok && (s.kv_mask == nullptr || s.kv_mask[(size_t)b * s.S + j] != 0);
      }
    };
    fetch_cols(0);
    for (int t = 0; t < T; ++t) {
      const int par = t & 1;
      if (tid < FA_BN) {
        colinfo[par][0][tid] = n_ik;
        colinfo[par][1][tid] = n_ok ? 0.f : NEG_INF;
        colinfo[par][2][tid] = n_kv ? 0.f : NEG_INF;
      }
      fetch_cols(t + 1);
      softmax_warps_sync();
      const float4* ik4 = reinterpret_cast<const float4*>(colinfo[par][0]);
      const float4* kb4 = reinterpret_cast<const float4*>(use_mask ? colinfo[par][2] : colinfo[par][1]);
      mbar_wait(&bar_s[par], (uint32_t)((t >> 1) & 1));
      tcgen05_fence_after();
      uint32_t a0[32], a1[32];
      tmem_ld_32x32b_x32(lane_addr + (uint32_t)(par * FA_BN), a0);
      tmem_ld_32x32b_x32(lane_addr + (uint32_t)(par * FA_BN + 32), a1);
      tmem_wait_ld();
      float m_tile = NEG_INF;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 i0 = ik4[j4], b0 = kb4[j4], i1 = ik4[8 + j4], b1 = kb4[8 + j4];
        const float v0 = fmaf(__uint_as_float(a0[4 * j4]), iq * i0.x, b0.x), v1 = fmaf(__uint_as_float(a0[4 * j4 + 1]), iq * i0.y, b0.y);
        const float v2 = fmaf(__uint_as_float(a0[4 * j4 + 2]), iq * i0.z, b0.z), v3 = fmaf(__uint_as_float(a0[4 * j4 + 3]), iq * i0.w, b0.w);
        const float w0 = fmaf(__uint_as_float(a1[4 * j4]), iq * i1.x, b1.x), w1 = fmaf(__uint_as_float(a1[4 * j4 + 1]), iq * i1.y, b1.y);
        const float w2 = fmaf(__uint_as_float(a1[4 * j4 + 2]), iq * i1.z, b1.z), w3 = fmaf(__uint_as_float(a1[4 * j4 + 3]), iq * i1.w, b1.w);
        a0[4 * j4] = __float_as_uint(v0); a0[4 * j4 + 1] = __float_as_uint(v1); a0[4 * j4 + 2] = __float_as_uint(v2); a0[4 * j4 + 3] = __float_as_uint(v3);
        a1[4 * j4] = __float_as_uint(w0); a1[4 * j4 + 1] = __float_as_uint(w1); a1[4 * j4 + 2] = __float_as_uint(w2); a1[4 * j4 + 3] = __float_as_uint(w3);
        m_tile = fmaxf(m_tile, fmaxf(fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)), fmaxf(fmaxf(w0, w1), fmaxf(w2, w3))));
      }
      const bool need = m_tile > m_ref + FA_TAU;       // (m_ref = -inf: any finite maximum moves it)
      const float m_new = need ? m_tile : m_ref;
      const float m_use = m_new == NEG_INF ? 0.f : m_new;   // nothing but masked keys so far: P = 2^(-inf) = 0, not NaN
      const float off = FA_PSHIFT - m_use;
      // P of this tile, split into fp16 halves: a0 <- packed hi pairs / lo pairs (16 words each), likewise a1
      float lsum = 0.f;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        unsigned short h0, l0, h1, l1;
        const float p0 = fa_ex2(__uint_as_float(a0[2 * j]) + off), p1 = fa_ex2(__uint_as_float(a0[2 * j + 1]) + off);
        lsum += p0 + p1;
        split16(p0, h0, l0);
        split16(p1, h1, l1);
        hi[j] = pack16(h0, h1);
        lo[j] = pack16(l0, l1);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        unsigned short h0, l0, h1, l1;
        const float p0 = fa_ex2(__uint_as_float(a1[2 * j]) + off), p1 = fa_ex2(__uint_as_float(a1[2 * j + 1]) + off);
        lsum += p0 + p1;
        split16(p0, h0, l0);
        split16(p1, h1, l1);
        hi[16 + j] = pack16(h0, h1);
        lo[16 + j] = pack16(l0, l1);
      }
      if (t > 0) {
        mbar_wait(&bar_pv, (uint32_t)((t - 1) & 1));   // P.V of the previous tile has read P and updated O
        tcgen05_fence_after();
        if (__any_sync(0xffffffffu, need)) {
          const float f = need ? fa_ex2(m_ref - m_new) : 1.f;   // (m_ref = -inf: O and l are zero, f = 0)
          l *= f;
          for (int c0 = 0; c0 < s.ND; c0 += 16) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(lane_addr + FA_O_COL + (uint32_t)c0, r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * f);
            tmem_st_32x32b_x16(lane_addr + FA_O_COL + (uint32_t)c0, r);
          }
          tmem_wait_st();
        }
      }
      m_ref = m_new;
      l += lsum;
      // the A operand of P.V: K-major rows of 64 fp16 = 128 bytes, 16-byte chunk c of row r at chunk position c ^ (r & 7)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int pos = (c ^ (row & 7)) << 4;
        *reinterpret_cast<uint4*>(p_row + pos) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
        *reinterpret_cast<uint4*>(p_row + FA_P_BYTES + pos) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
      }
      fence_proxy_async();
      tcgen05_fence_before();
      mbar_arrive(&bar_p);
    }
    // ---- epilogue: O / l, the V^T rows' scales undone, [B, L, H * d] ----
    mbar_wait(&bar_pv, (uint32_t)((T - 1) & 1));
    tcgen05_fence_after();
    const float inv_l = 1.f / l;     // l = 0 (a valid query without a valid key): 0 * inf = NaN, as softmax of all -inf is
    float* orow = s.out + (((size_t)b * s.L + (row_ok ? q : 0)) * s.H + h) * s.d;
    for (int c0 = 0; c0 < s.ND; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(lane_addr + FA_O_COL + (uint32_t)c0, r);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          if (c0 + j < s.d)
            *reinterpret_cast<float4*>(orow + c0 + j) =
                make_float4(__uint_as_float(r[j]) * inv_l * vscale[c0 + j], __uint_as_float(r[j + 1]) * inv_l * vscale[c0 + j + 1],
                            __uint_as_float(r[j + 2]) * inv_l * vscale[c0 + j + 2], __uint_as_float(r[j + 3]) * inv_l * vscale[c0 + j + 3]);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace drg

using namespace drg;

extern "C" int drg_attention_split16(const void* Q16, const void* K16, const void* Vt16, const uint8_t* q_mask, const uint8_t* kv_mask,
                                     int B, int H, int L, int S, int d, float scale, float* out, void* stream) {
  DRG_CHECK_ARG(Q16 && K16 && Vt16 && out, "Q16 / K16 / Vt16 / out must be non-null");
  DRG_CHECK_ARG(B >= 1 && H >= 1 && L >= 1 && S >= 1 && d >= 1, "B, H, L, S, d must be >= 1");
  if (d % 4 != 0 || d > 176 || (((uintptr_t)Q16 | (uintptr_t)K16 | (uintptr_t)Vt16 | (uintptr_t)out) & 15u)) {
    set_error("attention: head width must be a multiple of 4 and <= 176 (got %d), buffers 16-byte aligned", d);
    return DRG_ERR_UNSUPPORTED;
  }
  FlashShape s{};
  s.B = B; s.H = H; s.L = L; s.S = S; s.d = d;
  s.kc = split16_kc(d);
  s.kcS = split16_kc(S);
  s.ND = (d + 15) & ~15;
  s.scale2 = scale * LOG2E;
  s.Q16 = reinterpret_cast<const unsigned short*>(Q16);
  s.K16 = reinterpret_cast<const unsigned short*>(K16);
  s.V16 = reinterpret_cast<const unsigned short*>(Vt16);
  s.q_mask = q_mask;
  s.kv_mask = kv_mask;
  s.out = out;
  const int BH = B * H, nq = 2 * s.kc / 64;
  CUtensorMap tQ, tK, tV;
  if (!make_tmap(&tQ, Q16, BH, L, 2 * s.kc + 8, FA_BM, 64, 2)) return DRG_ERR_CUDA;
  if (!make_tmap(&tK, K16, BH, S, 2 * s.kc + 8, FA_BN, 64, 2)) return DRG_ERR_CUDA;
  if (!make_tmap(&tV, Vt16, BH, d, 2 * s.kcS + 8, s.ND, 64, 2)) return DRG_ERR_CUDA;
  const size_t smem = 1024 + (size_t)nq * (FA_Q_CHUNK + FA_K_CHUNK) + (size_t)2 * s.ND * 128 + 2 * FA_P_BYTES;
  if (smem + 3072 > 227 * 1024) {   // (+ the kernel's static shared memory)
    set_error("attention: %zu bytes of shared memory for head width %d", smem, d);
    return DRG_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)((L + FA_BM - 1) / FA_BM), (unsigned)BH);
  if (BH > 65535) {
    set_error("attention: batch * heads = %d exceeds the grid", BH);
    return DRG_ERR_UNSUPPORTED;
  }
  if (FA_O_COL + s.ND <= 256) {
    DRG_CUDA((cudaFuncSetAttribute(flash_attn_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    flash_attn_kernel<256><<<grid, FA_THREADS, smem, st>>>(tQ, tK, tV, s);
  } else {
    DRG_CUDA((cudaFuncSetAttribute(flash_attn_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    flash_attn_kernel<512><<<grid, FA_THREADS, smem, st>>>(tQ, tK, tV, s);
  }
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}
