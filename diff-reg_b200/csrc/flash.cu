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
// One CTA = 128 queries of one (batch, head); keys in tiles of 64.  320 threads:
//   warp 5   TMA producer (lane 0): Q once, then the K tiles (one tile ahead) and V^T tiles through two rings of shared-memory
//            stages (K double-buffered at every supported head width, V^T where it fits), freed by the tensor core's commits;
//            all lanes: the key tiles' column info (operand row scales, mask bytes) through a ring of four slots
//   warp 4   MMA issuer, the whole warp converged with one ELECTED lane per instruction: S = Q.K^T of tile t+1 into one of TWO
//            TMEM accumulators while the softmax warps work on tile t, then -- once they have published P -- O += P.V of tile t
//            (O lives in TMEM for the whole pass).  Straight-line issue code (the kernel is templated on the operand chunk
//            counts): a 128 x 64 x 16 MMA occupies the pipe for 32 cycles, a loop with run-time bounds needs ~90 per MMA.
//   warps 0..3, 6..9   softmax: two threads per query row (= TMEM lane), 32 of the tile's 64 logits each.  tcgen05.ld, row /
//            column scales of the split operands, masks, running maximum (the halves exchange theirs through shared memory),
//            exp2, row sum, fp16 hi / lo split of P (packed conversions) written as the 128-byte-swizzled K-major A operand
//            of the P.V product
// Head widths that are not multiples of 64 keep their Q / K tiles as 64-column chunks (128-byte swizzle) plus 16-column pieces
// (32-byte swizzle): the 132-wide heads of 4DMatch take 27 instead of 36 k-steps per S tile and 72 + 36 KB instead of 96 + 48.
// Online softmax with a LAZY reference: P = 2^(s - m_ref + 6) with m_ref only moved (and O, l rescaled through tcgen05.ld / st)
// when the tile's maximum exceeds it by more than 8 (log2 units), so P <= 2^14 stays inside fp16 and the rescale of the
// accumulator is rare after the first tiles.  The 2^6 and the stale reference cancel in O / l.
#include "umma.cuh"

namespace drg {

constexpr int FA_BM = 128;        // queries per CTA (TMEM lanes)
constexpr int FA_BN = 64;         // keys per tile = one 128-byte swizzle-atom row of 16-bit P
constexpr int FA_PARTS = 2;         // softmax threads per query row (measured: 4 -- sixteen softmax warps -- is slower, 122 vs 112 us)
constexpr int FA_CP = 64 / FA_PARTS;   // columns of a key tile per softmax thread
constexpr int FA_THREADS = 32 * (4 * FA_PARTS + 2);   // 4 FA_PARTS softmax warps + the MMA issuer (warp 4) + the TMA producer (warp 5)
constexpr float FA_TAU = 8.f;     // move the reference when a tile's maximum exceeds it by more than this (log2 units)
constexpr int FA_CH = 8;           // key tiles per accumulation chunk of O (see the drain in the softmax warps)
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
  int nfull, npiece;  // a Q / K operand segment in shared memory: nfull 64-column chunks (128-byte swizzle) + npiece 16-column
                      // pieces (32-byte swizzle) -- d rounded up to 16, not to 64: the 132-wide heads of 4DMatch take 9 k-steps
                      // per product term instead of 12 and 72 + 36 KB for the Q / K tiles instead of 96 + 48
  int KS, VS;    // shared-memory stages of the K / V^T tile rings (1 or 2)
  int ND;        // d rounded up to 16: N of the P.V MMA (rows of the V^T tile; rows >= d are TMA zero fill)
  float scale2;  // softmax scale * log2(e)
  const unsigned short *Q16, *K16, *V16;
  const uint8_t *q_mask, *kv_mask;
  float* out;    // [B, L, H * d]
  int nsplit;    // the keys are split over nsplit CTAs per (query tile, head) (grid.z); > 1: partial results go to `part` / `part_ml`
  int tiles_per_split;
  float* part;   // [nsplit, B, L, H * d] un-normalised partial outputs (relative to the split's own reference)
  float2* part_ml;  // [nsplit, B * H, L] (reference m in log2 units, row sum l)
  long long* tl; // tuning stamps of CTA (0, 0): 16 per key tile for the first 64 tiles, or NULL
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
// K-major operand tile whose rows are 32 bytes (16 fp16) wide, 32-byte swizzle: groups of 8 rows (256 bytes) 256 B apart
__device__ __forceinline__ uint64_t make_smem_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(256u >> 4) << 32;   // stride byte offset
  d |= (uint64_t)1 << 46;             // descriptor version (sm_100)
  d |= (uint64_t)6 << 61;             // SWIZZLE_32B
  return d;
}
// The MMA warp runs its loop with all 32 lanes converged and ELECTS one lane per instruction: descriptors, addresses and loop
// counters stay warp-uniform (uniform registers), so a tcgen05.mma costs a handful of issue slots.  Under `if (lane == 0)` the
// compiler cannot prove uniformity and wraps every UTCHMMA in an elect / branch loop: ~75 cycles per MMA measured
// (tools/fa_timeline.py), more than the 32 cycles a 128 x 64 x 16 MMA occupies the tensor pipe.
__device__ __forceinline__ void umma_f16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_warps_sync() { asm volatile("bar.sync 1, %0;" ::"n"(128 * FA_PARTS) : "memory"); }
// (x0, x1) = hi + lo, both fp16 pairs: one packed conversion per pair (F2FP) instead of two scalar ones on the XU pipe
__device__ __forceinline__ void split16x2(float x0, float x1, uint32_t& hi2, uint32_t& lo2) {
  uint32_t h;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));      // low half = x0
  float h0, h1;
  asm("{\n.reg .b16 a, b;\nmov.b32 {a, b}, %2;\ncvt.f32.f16 %0, a;\ncvt.f32.f16 %1, b;\n}" : "=f"(h0), "=f"(h1) : "r"(h));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(x1 - h1), "f"(x0 - h0));
  hi2 = h;
}

// Tuning stamps (tools/fa_timeline.py) are compiled in only with -DDRG_FA_TIMELINE: even predicated off, their stores sit in the
// softmax warps' instruction stream
#ifdef DRG_FA_TIMELINE
#define FA_STAMP(k)                                                                                  \
  do {                                                                                               \
    if (s.tl && blockIdx.x == 0 && blockIdx.y == 0 && t < 64 && stid == 0) s.tl[t * 16 + (k)] = clock64(); \
  } while (0)
#define FA_STAMP_MMA(tt, k)                                                                                  \
  do {                                                                                                       \
    if (s.tl && blockIdx.x == 0 && blockIdx.y == 0 && (tt) >= 0 && (tt) < 64 && lane == 0) s.tl[(tt) * 16 + (k)] = clock64(); \
    __syncwarp();                                                                                            \
  } while (0)
#else
#define FA_STAMP(k) do { } while (0)
#define FA_STAMP_MMA(tt, k) do { } while (0)
#endif

template <int NFULL, int NPIECE>
__global__ void __launch_bounds__(FA_THREADS, 1)
    flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmQp,
                      const __grid_constant__ CUtensorMap tmKp, const FlashShape s) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t bar_q, bar_p, bar_pv;
  __shared__ __align__(8) uint64_t bar_s[2], bar_kfull[2], bar_kfree[2], bar_vfull[2], bar_vfree[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(8) uint64_t bar_ci[4];
  __shared__ __align__(16) float colinfo[4][3][FA_BN];   // ring of four key tiles: 1 / scale of the key rows, bias (0 / -inf) without and with the key mask
  __shared__ float rowmax[2][FA_PARTS][FA_BM];   // per tile parity: the column parts' row maxima
  __shared__ float rowsum[FA_PARTS][FA_BM];      // the parts' row sums (epilogue)
  __shared__ float vscale[256];            // 1 / scale of the V^T rows (= output channels)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, b = bh / s.H, h = bh - b * s.H;
  const int q0 = blockIdx.x * FA_BM;
  // NFULL 64-column chunks + NPIECE 16-column pieces per operand segment: compile-time, so that the MMA warp's issue loops are
  // straight-line code (a loop with run-time bounds costs ~90 cycles per MMA in that single warp, unrolled code ~40)
  constexpr int nfull = NFULL, npiece = NPIECE;
  constexpr int TMEM_COLS = (FA_S_COLS + 2 * 16 * (4 * NFULL + NPIECE)) <= 256 ? 256 : 512;   // S x 2 + O chunk + O total (ND = 16 (4 NFULL + NPIECE) columns each)
  const uint32_t q_seg = (uint32_t)(nfull * FA_Q_CHUNK + npiece * (FA_Q_CHUNK / 4));   // bytes of one segment (lo or hi) of the Q tile
  const uint32_t k_seg = (uint32_t)(nfull * FA_K_CHUNK + npiece * (FA_K_CHUNK / 4));
  // this CTA's key tiles: [tile0, tile0 + T) of the (S + 63) / 64 tiles (all of them unless the keys are split over grid.z)
  const int tile0 = (int)blockIdx.z * s.tiles_per_split;
  const int T = min(s.tiles_per_split, (s.S + FA_BN - 1) / FA_BN - tile0);

  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* aligned = smem_dyn + (base - smem_u32(smem_dyn));
  uint8_t* sQ = aligned;
  uint8_t* sK = sQ + (size_t)2 * q_seg;
  const uint32_t k_bytes = 2u * k_seg;                  // one K tile stage
  const uint32_t v_bytes = (uint32_t)s.ND * 128u;       // one half (hi or lo) of a V^T tile stage
  uint8_t* sV = sK + (size_t)s.KS * k_bytes;
  uint8_t* sP = sV + (size_t)s.VS * 2 * v_bytes;   // [hi | lo]

  if (tid == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    if (npiece) {
      prefetch_tmap(&tmQp);
      prefetch_tmap(&tmKp);
    }
    mbar_init(&bar_q, 1);
    mbar_init(&bar_p, FA_PARTS * FA_BM);
    mbar_init(&bar_pv, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_ci[i], 32);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_kfull[i], 1);
      mbar_init(&bar_kfree[i], 1);
      mbar_init(&bar_vfull[i], 1);
      mbar_init(&bar_vfree[i], 1);
    }
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

  if (warp == 5) {
    {
      // ===================== TMA producer (lane 0) + the key tiles' column info (all lanes) =====================
      // Column info of a key tile = the K operand rows' 1 / scale and the mask bytes: scattered 4- / 1-byte global loads.  In the
      // softmax warps such a load stalled its warp for ~500 cycles per tile at ISSUE, however far ahead it was requested
      // (tools/fa_timeline.py; 117.7 -> 103.3 us without them), so this warp -- which runs ahead of everybody anyway -- fetches
      // them one tile ahead into registers and publishes them through a ring of four shared-memory slots (bar_ci).  A slot's
      // previous user, tile t - 4, is done: K(t)'s stage is only free once S(t - KS) has completed, which the MMA warp issued
      // after the softmax warps had published P(t - KS - 1).
      const float NEG_INF = __int_as_float(0xff800000);
      float c_ik[2] = {0.f, 0.f};
      bool c_ok[2] = {false, false}, c_kv[2] = {false, false};
      auto fetch_cols = [&](int t) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int j = (tile0 + t) * FA_BN + lane + 32 * i;
          c_ok[i] = t < T && j < s.S;
          c_ik[i] = c_ok[i] ? reinterpret_cast<const float*>(s.K16 + ((size_t)bh * s.S + j) * pitch_d + 2 * s.kc)[0] : 0.f;
          c_kv[i] = c_ok[i] && (s.kv_mask == nullptr || s.kv_mask[(size_t)b * s.S + j] != 0);
        }
      };
      auto publish_cols = [&](int t) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          colinfo[t & 3][0][lane + 32 * i] = c_ik[i];
          colinfo[t & 3][1][lane + 32 * i] = c_ok[i] ? 0.f : NEG_INF;
          colinfo[t & 3][2][lane + 32 * i] = c_kv[i] ? 0.f : NEG_INF;
        }
        mbar_arrive(&bar_ci[t & 3]);     // (release; 32 arrivals complete the phase)
      };
      fetch_cols(0);
      auto load_k = [&](int t) {
        const int st = t % s.KS;
        mbar_wait(&bar_kfree[st], (uint32_t)(((t / s.KS) & 1) ^ 1));   // S of tile t - KS has read the stage (first pass: free)
        mbar_arrive_expect_tx(&bar_kfull[st], k_bytes);
        uint8_t* dst = sK + (size_t)st * k_bytes;
        for (int seg = 0; seg < 2; ++seg) {
          uint8_t* sd = dst + (size_t)seg * k_seg;
          for (int c = 0; c < nfull; ++c) tma_load_3d(sd + (size_t)c * FA_K_CHUNK, &tmK, seg * s.kc + c * 64, (tile0 + t) * FA_BN, bh, &bar_kfull[st]);
          for (int c = 0; c < npiece; ++c)
            tma_load_3d(sd + (size_t)nfull * FA_K_CHUNK + (size_t)c * (FA_K_CHUNK / 4), &tmKp, seg * s.kc + nfull * 64 + c * 16, (tile0 + t) * FA_BN, bh,
                        &bar_kfull[st]);
        }
      };
      auto load_v = [&](int t) {
        const int st = t % s.VS;
        mbar_wait(&bar_vfree[st], (uint32_t)(((t / s.VS) & 1) ^ 1));   // P.V of tile t - VS has read the stage
        mbar_arrive_expect_tx(&bar_vfull[st], 2u * v_bytes);
        uint8_t* dst = sV + (size_t)st * 2 * v_bytes;
        tma_load_3d(dst, &tmV, (tile0 + t) * FA_BN, 0, bh, &bar_vfull[st]);                     // V_hi: keys of this tile along the row
        tma_load_3d(dst + v_bytes, &tmV, s.kcS + (tile0 + t) * FA_BN, 0, bh, &bar_vfull[st]);   // V_lo
      };
      if (lane == 0) {
        mbar_arrive_expect_tx(&bar_q, 2u * q_seg);
        for (int seg = 0; seg < 2; ++seg) {
          uint8_t* sd = sQ + (size_t)seg * q_seg;
          for (int c = 0; c < nfull; ++c) tma_load_3d(sd + (size_t)c * FA_Q_CHUNK, &tmQ, seg * s.kc + c * 64, q0, bh, &bar_q);
          for (int c = 0; c < npiece; ++c)
            tma_load_3d(sd + (size_t)nfull * FA_Q_CHUNK + (size_t)c * (FA_Q_CHUNK / 4), &tmQp, seg * s.kc + nfull * 64 + c * 16, q0, bh, &bar_q);
        }
        load_k(0);
      }
      __syncwarp();
      publish_cols(0);
      fetch_cols(1);
      for (int t = 0; t < T; ++t) {     // K runs one tile ahead of V: S(t+1) is issued before P.V(t), so its stage frees first
        if (t + 1 < T) {
          if (lane == 0) load_k(t + 1);
          __syncwarp();
          publish_cols(t + 1);
          fetch_cols(t + 2);
        }
        if (lane == 0) load_v(t);
        __syncwarp();
      }
    }
  } else if (warp == 4) {
    {
      // ===================== MMA issuer: the whole warp, converged; one elected lane per instruction =====================
      // (a spin loop leaves the lanes formally diverged: reconverge, or everything after it is compiled for a diverged warp)
      auto wwait = [&](uint64_t* bar, uint32_t parity) {
        mbar_wait(bar, parity);
        __syncwarp();
      };
      const uint32_t idesc_s = make_idesc_f16(FA_BM, FA_BN, FMT_F16, FMT_F16);
      const uint32_t idesc_o = make_idesc_f16(FA_BM, s.ND, FMT_F16, FMT_F16);
      const uint64_t q128 = make_smem_desc_sw128(smem_u32(sQ)), q32 = make_smem_desc_sw32(smem_u32(sQ));
      const uint64_t k128 = make_smem_desc_sw128(smem_u32(sK)), k32 = make_smem_desc_sw32(smem_u32(sK));
      auto issue_s = [&](int t) {
        const int st = t % s.KS;
        wwait(&bar_kfull[st], (uint32_t)((t / s.KS) & 1));
        tcgen05_fence_after();
        FA_STAMP_MMA(t - 1, 13);
        const uint32_t d_tmem = tmem_base + (uint32_t)((t & 1) * FA_BN);
        uint32_t acc = 0u;
        // Descriptors are base + constant offsets in 16-byte units (the start-address field is the low word: plain 64-bit
        // adds, all warp-uniform).  The small terms first (the accumulator is rounded at every step): Q' = [lo | hi],
        // K' = [hi | lo], so segment 0 x segment 0 is Q_lo.K_hi and segment 1 x segment 1 is Q_hi.K_lo; then Q_hi.K_hi
        const uint64_t kofs = (uint64_t)(((uint32_t)st * k_bytes) >> 4);
        auto term = [&](uint32_t qs, uint32_t ks) {
          const uint64_t qo = (uint64_t)((qs * q_seg) >> 4), ko = kofs + (uint64_t)((ks * k_seg) >> 4);
#pragma unroll
          for (int c = 0; c < nfull; ++c) {
            const uint64_t a_desc = q128 + qo + (uint64_t)(c * (FA_Q_CHUNK >> 4)), b_desc = k128 + ko + (uint64_t)(c * (FA_K_CHUNK >> 4));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_f16_elect(d_tmem, a_desc + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc_s, acc);
              acc = 1u;
            }
          }
#pragma unroll
          for (int c = 0; c < npiece; ++c) {     // 16-column pieces: one K = 16 step each, rows of 32 bytes
            umma_f16_elect(d_tmem, q32 + qo + (uint64_t)(nfull * (FA_Q_CHUNK >> 4) + c * (FA_Q_CHUNK >> 6)),
                           k32 + ko + (uint64_t)(nfull * (FA_K_CHUNK >> 4) + c * (FA_K_CHUNK >> 6)), idesc_s, acc);
            acc = 1u;
          }
        };
        term(0, 0);
        term(1, 1);
        term(1, 0);
        FA_STAMP_MMA(t - 1, 14);
        umma_commit_elect(&bar_kfree[st]);      // the K stage may be refilled ...
        umma_commit_elect(&bar_s[t & 1]);       // ... and the logits are ready
      };
      wwait(&bar_q, 0u);
      issue_s(0);
      const uint64_t ph_desc = make_smem_desc_sw128(smem_u32(sP));
      const uint64_t pl_desc = make_smem_desc_sw128(smem_u32(sP + FA_P_BYTES));
      const uint32_t o_tmem = tmem_base + FA_O_COL;
      for (int t = 0; t < T; ++t) {
        // S(t+1) into the other accumulator, whose tile t - 1 the softmax warps have consumed (bar_p of t - 1, waited below)
        FA_STAMP_MMA(t, 0);
        if (t + 1 < T) issue_s(t + 1);
        FA_STAMP_MMA(t, 1);
        const int vs = t % s.VS;
        wwait(&bar_vfull[vs], (uint32_t)((t / s.VS) & 1));
        wwait(&bar_p, (uint32_t)(t & 1));
        tcgen05_fence_after();
        FA_STAMP_MMA(t, 2);
        const uint64_t vh_desc = make_smem_desc_sw128(smem_u32(sV + (size_t)vs * 2 * v_bytes));
        const uint64_t vl_desc = make_smem_desc_sw128(smem_u32(sV + (size_t)vs * 2 * v_bytes + v_bytes));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t o = (uint64_t)(2 * kk);
          umma_f16_elect(o_tmem, pl_desc + o, vh_desc + o, idesc_o, (uint32_t)(((t % FA_CH) | kk) != 0));   // P_lo . V_hi (a new chunk of FA_CH tiles starts from zero)
          umma_f16_elect(o_tmem, ph_desc + o, vl_desc + o, idesc_o, 1u);                          // P_hi . V_lo
          umma_f16_elect(o_tmem, ph_desc + o, vh_desc + o, idesc_o, 1u);                          // P_hi . V_hi
        }
        umma_commit_elect(&bar_vfree[vs]);
        umma_commit_elect(&bar_pv);             // P may be overwritten, O holds tile t
        FA_STAMP_MMA(t, 3);
      }
    }
  } else {
    // ===================== softmax warps: FA_PARTS threads per query row (= TMEM lane), 64 / FA_PARTS of the tile's 64 keys each =====================
    // The chain wait-S / tcgen05.ld / scale / max / exp2 / split / store / arrive of a tile is latency bound (measured with two
    // threads per row: 2500 cycles per tile at 35 % issue utilisation), so it is spread over sixteen warps.  A warp may only read
    // TMEM lanes 32 (w % 4) ...: the four warps of a lane quadrant take the four 16-column parts of the tile.
    const int quad = warp & 3, part = warp < 4 ? 0 : 1 + (warp - 6) / 4;   // warps 0..3, 6..9 (, 10..13, 14..17)
    [[maybe_unused]] const int stid = (part * 4 + quad) * 32 + lane;        // 0..511
    const int row = quad * 32 + lane, q = q0 + row;
    const bool row_ok = q < s.L;
    const float iq = row_ok ? reinterpret_cast<const float*>(s.Q16 + ((size_t)bh * s.L + q) * pitch_d + 2 * s.kc)[0] * s.scale2 : 1.f;
    // keys are masked for valid queries only (transformer.py:80-81); no query mask = every query valid (vision3d's k_masks)
    const bool use_mask = s.kv_mask != nullptr && (s.q_mask == nullptr || (row_ok && s.q_mask[(size_t)b * s.L + q] != 0));
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float NEG_INF = __int_as_float(0xff800000);
    float m_ref = NEG_INF, l = 0.f;                        // l: this thread's part of the row sum (the parts share m_ref)
    uint8_t* p_row = sP + (size_t)(row >> 3) * 1024 + (size_t)(row & 7) * 128;
    // this thread's 16-column chunks of the O accumulator (rescale, epilogue)
    const int n16 = s.ND / 16, c16_lo = (part * n16) / FA_PARTS, c16_hi = ((part + 1) * n16) / FA_PARTS;
    for (int t = 0; t < T; ++t) {
      const int par = t & 1;
      FA_STAMP(15);
      const float4* ik4 = reinterpret_cast<const float4*>(colinfo[t & 3][0] + part * FA_CP);
      const float4* kb4 = reinterpret_cast<const float4*>((use_mask ? colinfo[t & 3][2] : colinfo[t & 3][1]) + part * FA_CP);
      mbar_wait(&bar_ci[t & 3], (uint32_t)((t >> 2) & 1));     // the tile's column info (producer warp)
      FA_STAMP(4);
      mbar_wait(&bar_s[par], (uint32_t)((t >> 1) & 1));
      tcgen05_fence_after();
      FA_STAMP(5);
      uint32_t a0[FA_CP];
#pragma unroll
      for (int c = 0; c < FA_CP / 16; ++c)
        tmem_ld_32x32b_x16(lane_addr + (uint32_t)(par * FA_BN + part * FA_CP + 16 * c), *reinterpret_cast<uint32_t(*)[16]>(a0 + 16 * c));
      tmem_wait_ld();
      FA_STAMP(6);
      // x_j = acc_j / scale(key j) (+ -inf where masked); the logit is iq x_j with iq > 0 the row's factor: the maximum is taken
      // over x and the factor rides in the exponent's FFMA
      float m_part = NEG_INF;
#pragma unroll
      for (int j4 = 0; j4 < FA_CP / 4; ++j4) {
        const float4 i0 = ik4[j4], b0 = kb4[j4];
        const float v0 = fmaf(__uint_as_float(a0[4 * j4]), i0.x, b0.x), v1 = fmaf(__uint_as_float(a0[4 * j4 + 1]), i0.y, b0.y);
        const float v2 = fmaf(__uint_as_float(a0[4 * j4 + 2]), i0.z, b0.z), v3 = fmaf(__uint_as_float(a0[4 * j4 + 3]), i0.w, b0.w);
        a0[4 * j4] = __float_as_uint(v0); a0[4 * j4 + 1] = __float_as_uint(v1); a0[4 * j4 + 2] = __float_as_uint(v2); a0[4 * j4 + 3] = __float_as_uint(v3);
        m_part = fmaxf(m_part, fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)));
      }
      m_part *= iq;
      rowmax[par][part][row] = m_part;
      FA_STAMP(7);
      softmax_warps_sync();     // the other parts' maxima
      FA_STAMP(8);
      float m_tile = rowmax[par][0][row];
#pragma unroll
      for (int pp = 1; pp < FA_PARTS; ++pp) m_tile = fmaxf(m_tile, rowmax[par][pp][row]);
      const bool need = m_tile > m_ref + FA_TAU;       // (m_ref = -inf: any finite maximum moves it); identical in all parts
      const float m_new = need ? m_tile : m_ref;
      const float m_use = m_new == NEG_INF ? 0.f : m_new;   // nothing but masked keys so far: P = 2^(-inf) = 0, not NaN
      const float off = FA_PSHIFT - m_use;
      // P of this part of the tile, split into fp16 halves, two entries per conversion
      float lsum = 0.f;
      uint32_t hi[FA_CP / 2], lo[FA_CP / 2];
#pragma unroll
      for (int j = 0; j < FA_CP / 2; ++j) {
        const float p0 = fa_ex2(fmaf(__uint_as_float(a0[2 * j]), iq, off)), p1 = fa_ex2(fmaf(__uint_as_float(a0[2 * j + 1]), iq, off));
        lsum += p0 + p1;
        split16x2(p0, p1, hi[j], lo[j]);
      }
      FA_STAMP(9);
      if (t > 0) {
        mbar_wait(&bar_pv, (uint32_t)((t - 1) & 1));   // P.V of the previous tile has read P and updated O
        tcgen05_fence_after();
        FA_STAMP(10);
        // The tensor core's fp32 accumulator TRUNCATES at every step: over thousands of keys (hundreds of MMAs into one accumulator)
        // that bias reaches ~2^-17 relative.  O is therefore accumulated in chunks of FA_CH key tiles (96 MMAs) and the chunks are
        // summed here with round-to-nearest adds into a second TMEM region (O total, columns FA_O_COL + ND ...).
        const bool drain = (t % FA_CH) == 0;          // O chunk is complete: P.V of this tile starts the next one from zero
        const bool has_total = t > FA_CH;             // (the first drain, at t == FA_CH, initialises O total)
        const bool any_need = __any_sync(0xffffffffu, need);
        if (any_need || drain) {
          const float f = need ? fa_ex2(m_ref - m_new) : 1.f;   // (m_ref = -inf: O and l are zero, f = 0)
          l *= f;
          for (int c = c16_lo; c < c16_hi; ++c) {
            uint32_t r[16], tt[16];
            const uint32_t col = FA_O_COL + (uint32_t)(c * 16);
            tmem_ld_32x32b_x16(lane_addr + col, r);
            if (has_total) tmem_ld_32x32b_x16(lane_addr + col + (uint32_t)s.ND, tt);
            tmem_wait_ld();
            if (drain) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                tt[j] = __float_as_uint(has_total ? fmaf(__uint_as_float(tt[j]), f, __uint_as_float(r[j]) * f) : __uint_as_float(r[j]) * f);
              tmem_st_32x32b_x16(lane_addr + col + (uint32_t)s.ND, tt);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * f);
              tmem_st_32x32b_x16(lane_addr + col, r);
              if (has_total) {
#pragma unroll
                for (int j = 0; j < 16; ++j) tt[j] = __float_as_uint(__uint_as_float(tt[j]) * f);
                tmem_st_32x32b_x16(lane_addr + col + (uint32_t)s.ND, tt);
              }
            }
          }
          tmem_wait_st();
        }
      }
      m_ref = m_new;
      l += lsum;
      // the A operand of P.V: K-major rows of 64 fp16 = 128 bytes, 16-byte chunk c of row r at chunk position c ^ (r & 7)
#pragma unroll
      for (int c = 0; c < FA_CP / 8; ++c) {
        const int pos = (((FA_CP / 8) * part + c) ^ (row & 7)) << 4;
        *reinterpret_cast<uint4*>(p_row + pos) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
        *reinterpret_cast<uint4*>(p_row + FA_P_BYTES + pos) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
      }
      FA_STAMP(11);
      fence_proxy_async();
      tcgen05_fence_before();
      mbar_arrive(&bar_p);
      FA_STAMP(12);
    }
    // ---- epilogue: O / l, the V^T rows' scales undone, [B, L, H * d] ----
    rowsum[part][row] = l;
    softmax_warps_sync();
    float l_row = rowsum[0][row];
#pragma unroll
    for (int pp = 1; pp < FA_PARTS; ++pp) l_row += rowsum[pp][row];
    mbar_wait(&bar_pv, (uint32_t)((T - 1) & 1));
    tcgen05_fence_after();
    // one CTA per (query tile, head): O / l.  Keys split over grid.z: the un-normalised O with (m_ref, l) for attn_combine_kernel
    const bool split = s.nsplit > 1;
    const float inv_l = split ? 1.f : 1.f / l_row;   // l = 0 (a valid query without a valid key): 0 * inf = NaN, as softmax of all -inf is
    float* orow = (split ? s.part + (size_t)blockIdx.z * s.B * s.L * s.H * s.d : s.out) + (((size_t)b * s.L + (row_ok ? q : 0)) * s.H + h) * s.d;
    if (split && part == 0 && row_ok) s.part_ml[((size_t)blockIdx.z * s.B * s.H + bh) * s.L + q] = make_float2(m_ref, l_row);
    for (int c = c16_lo; c < c16_hi; ++c) {
      const int c0 = c * 16;
      uint32_t r[16];
      tmem_ld_32x32b_x16(lane_addr + FA_O_COL + (uint32_t)c0, r);
      if (T > FA_CH) {                 // the chunks drained so far
        uint32_t tt[16];
        tmem_ld_32x32b_x16(lane_addr + FA_O_COL + (uint32_t)(s.ND + c0), tt);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(tt[j]));
      }
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

// Keys split over nsplit CTAs: out = sum_z 2^(m_z - m*) O_z / sum_z 2^(m_z - m*) l_z with m* = max_z m_z (each split's O_z and l_z are
// relative to its own reference m_z).  One thread per four output channels.
__global__ void __launch_bounds__(256) attn_combine_kernel(const float* __restrict__ part, const float2* __restrict__ part_ml, int nsplit,
                                                          int B, int H, int L, int d, float* __restrict__ out) {
  const int d4 = d >> 2;
  const size_t total = (size_t)B * L * H * d4;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % d4);
  const size_t r = idx / d4;             // (b * L + l) * H + h
  const int h = (int)(r % H);
  const size_t bl = r / H;
  const int l = (int)(bl % L), b = (int)(bl / L);
  const size_t slab = (size_t)B * L * H * d, mlslab = (size_t)B * H * L;
  const size_t mlo = ((size_t)b * H + h) * L + l;
  float m_star = __int_as_float(0xff800000);
  for (int z = 0; z < nsplit; ++z) m_star = fmaxf(m_star, part_ml[z * mlslab + mlo].x);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float lsum = 0.f;
  for (int z = 0; z < nsplit; ++z) {
    const float2 ml = part_ml[z * mlslab + mlo];
    const float w = ml.x == m_star ? 1.f : exp2f(ml.x - m_star);     // (all splits masked: m* = -inf, w = 1, l = 0 -> NaN below)
    const float4 o = *reinterpret_cast<const float4*>(part + z * slab + r * d + 4 * c4);
    acc.x = fmaf(w, o.x, acc.x); acc.y = fmaf(w, o.y, acc.y); acc.z = fmaf(w, o.z, acc.z); acc.w = fmaf(w, o.w, acc.w);
    lsum = fmaf(w, ml.y, lsum);
  }
  const float inv = 1.f / lsum;
  *reinterpret_cast<float4*>(out + r * d + 4 * c4) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}

// The V operand of the fused attention: V [B, S, H * d] fp32 -> per head the RIGHT split operand of V^T, [B * H, d, 2 kc(S) + 8]
// = [hi | lo | tail] with the KEYS along the row (what four launches -- permute, contiguous, pad, drg_prep_operand -- produced).
// Two kernels: the channels' maxima over the keys (row scale 2^e: maximum into [2^14, 2^15), the rule of features.cu) with
// atomicMax on the bit patterns of |x|, then 64-key x 32-channel tiles transposed through shared memory, split and written
// key-contiguous.  The tail holds 1 / scale only (nothing reads a norm of V^T rows).
__global__ void __launch_bounds__(256) vt_colmax_kernel(const float* __restrict__ V, int H, int S, int d, unsigned int* __restrict__ cmax) {
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int C = H * d;
  const float* vb = V + (size_t)b * S * C + (size_t)h * d;
  const int j0 = blockIdx.x * 128, j1 = min(j0 + 128, S);
  // thread = (channel quad, key phase): d / 4 quads side by side, 16-byte loads
  const int nquad = d >> 2, per = 256 / nquad > 0 ? 256 / nquad : 1;
  const int qd = threadIdx.x % nquad, ph = threadIdx.x / nquad;
  __shared__ unsigned int smax[1024];
  for (int i = threadIdx.x; i < d; i += 256) smax[i] = 0u;
  __syncthreads();
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = j0 + ph; j < (ph < per ? j1 : 0); j += per) {
    const float4 x = *reinterpret_cast<const float4*>(vb + (size_t)j * C + 4 * qd);
    m.x = fmaxf(m.x, fabsf(x.x)); m.y = fmaxf(m.y, fabsf(x.y)); m.z = fmaxf(m.z, fabsf(x.z)); m.w = fmaxf(m.w, fabsf(x.w));
    if (!(fabsf(x.x) <= 3.0e38f)) m.x = __int_as_float(0x7f800000);   // Inf / NaN: no scaling for that channel
    if (!(fabsf(x.y) <= 3.0e38f)) m.y = __int_as_float(0x7f800000);
    if (!(fabsf(x.z) <= 3.0e38f)) m.z = __int_as_float(0x7f800000);
    if (!(fabsf(x.w) <= 3.0e38f)) m.w = __int_as_float(0x7f800000);
  }
  // non-negative floats order like their bit patterns: the CTA's maxima in shared memory, then one global atomic per channel
  atomicMax(&smax[4 * qd], __float_as_uint(m.x));
  atomicMax(&smax[4 * qd + 1], __float_as_uint(m.y));
  atomicMax(&smax[4 * qd + 2], __float_as_uint(m.z));
  atomicMax(&smax[4 * qd + 3], __float_as_uint(m.w));
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += 256) atomicMax(cmax + (size_t)bh * d + i, smax[i]);
}

__global__ void __launch_bounds__(256) vt_split_kernel(const float* __restrict__ V, const unsigned int* __restrict__ cmax, int H, int S, int d,
                                                      unsigned short* __restrict__ out) {
  __shared__ float tile[64][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int bh = blockIdx.z, b = bh / H, h = bh - b * H, c0 = blockIdx.y * 32, j0 = blockIdx.x * 64;
  const int C = H * d, kcS = split16_kc(S);
  const size_t pitch = (size_t)2 * kcS + 8;
  const float* vb = V + (size_t)b * S * C + (size_t)h * d;
  const int c = c0 + tx;
  float sc = 1.f;
  if (c < d) {
    const float amax = __uint_as_float(cmax[(size_t)bh * d + c]);
    int e = 0;
    if (amax > 0.f && amax <= 3.0e38f) e = min(max(14 - ilogbf(amax), -126), 126);
    sc = __int_as_float((e + 127) << 23);
    if (j0 == 0 && ty == 0)
      *reinterpret_cast<float4*>(out + ((size_t)bh * d + c) * pitch + 2 * kcS) = make_float4(__int_as_float((127 - e) << 23), 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int key = j0 + ty + 8 * i;
    tile[ty + 8 * i][tx] = (key < S && c < d) ? vb[(size_t)key * C + c] * sc : 0.f;   // (keys S .. kc(S): the segment's zero padding)
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = ty + 8 * i;                  // 32 channels x 64 keys: lane = key pair
    if (c0 + ch < d) {
      unsigned short h0, l0, h1, l1;
      split16(tile[2 * tx][ch], h0, l0);
      split16(tile[2 * tx + 1][ch], h1, l1);
      unsigned short* o = out + ((size_t)bh * d + c0 + ch) * pitch + j0 + 2 * tx;
      *reinterpret_cast<uint32_t*>(o) = pack16(h0, h1);          // [hi | lo]: the right operand's pattern
      *reinterpret_cast<uint32_t*>(o + kcS) = pack16(l0, l1);
    }
  }
}

}  // namespace drg

using namespace drg;

// How many CTAs share the keys of one (query tile, head): with `units` such pairs of T key tiles each on NUM_SMS SMs (one CTA per
// SM), n splits cost ceil(units n / NUM_SMS) rounds of ceil(T / n) tiles + ~3 tiles' worth of prologue / epilogue per CTA
// (+ the combine pass).  4 heads x 4096 queries: 128 units -> 1; the 2D-3D flavour's 2048 / 4800 tokens (64 / 152 units) -> 2 .. 4.
constexpr int FA_MAX_SPLIT = 8;
static int fa_choose_nsplit(long long units, int T) {
  int best_n = 1;
  long long best = -1;
  for (int n : {1, 2, 3, 4, 6, 8}) {
    const int per = (T + n - 1) / n;
    if (n > 1 && per < 4) break;
    const long long rounds = (units * n + NUM_SMS - 1) / NUM_SMS;
    const long long cost = rounds * (per + 3) + (n > 1 ? 2 : 0);
    if (best < 0 || cost < best) {
      best = cost;
      best_n = n;
    }
  }
  return best_n;
}

// the split count drg_attention_split16 uses for nsplit = 0 (chosen) or a requested nsplit (clipped to the number of key tiles)
static int fa_effective_nsplit(int B, int H, int L, int S, int nsplit) {
  const int qtiles = (L + FA_BM - 1) / FA_BM, T = (S + FA_BN - 1) / FA_BN;
  if (nsplit == 0) nsplit = fa_choose_nsplit((long long)qtiles * B * H, T);
  if (nsplit > T) nsplit = T;
  const int per = (T + nsplit - 1) / nsplit;
  return (T + per - 1) / per;        // (no empty split)
}

extern "C" size_t drg_attention_workspace_bytes(int B, int H, int L, int S, int d, int nsplit) {
  if (B < 1 || H < 1 || L < 1 || S < 1 || d < 1 || nsplit < 0 || nsplit > FA_MAX_SPLIT) return 0;
  const int n = fa_effective_nsplit(B, H, L, S, nsplit);
  if (n <= 1) return 0;
  return align_up((size_t)n * ((size_t)B * L * H * d * sizeof(float) + (size_t)B * H * L * sizeof(float2)), 256);
}

extern "C" int drg_attention_split16(const void* Q16, const void* K16, const void* Vt16, const uint8_t* q_mask, const uint8_t* kv_mask,
                                     int B, int H, int L, int S, int d, float scale, float* out, int nsplit, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  DRG_CHECK_ARG(Q16 && K16 && Vt16 && out, "Q16 / K16 / Vt16 / out must be non-null");
  DRG_CHECK_ARG(nsplit >= 0 && nsplit <= FA_MAX_SPLIT, "nsplit must be 0 (choose) .. 8");
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
  s.tl = g_tuning_stamps;   // tuning hook (drg_tuning_set_stamp_buffer; tools/fa_timeline.py): NULL unless a tool set it
  const int BH = B * H;
  const int kc16 = (d + 15) & ~15;
  s.nfull = kc16 / 64;
  s.npiece = (kc16 % 64) / 16;
  CUtensorMap tQ, tK, tV, tQp, tKp;
  if (!make_tmap(&tQ, Q16, BH, L, 2 * s.kc + 8, FA_BM, 64, 2)) return DRG_ERR_CUDA;
  if (!make_tmap(&tK, K16, BH, S, 2 * s.kc + 8, FA_BN, 64, 2)) return DRG_ERR_CUDA;
  if (!make_tmap(&tV, Vt16, BH, d, 2 * s.kcS + 8, s.ND, 64, 2)) return DRG_ERR_CUDA;
  if (s.npiece) {
    if (!make_tmap(&tQp, Q16, BH, L, 2 * s.kc + 8, FA_BM, 16, 2, true, true)) return DRG_ERR_CUDA;
    if (!make_tmap(&tKp, K16, BH, S, 2 * s.kc + 8, FA_BN, 16, 2, true, true)) return DRG_ERR_CUDA;
  } else {
    tQp = tQ;
    tKp = tK;
  }
  const size_t budget = 227 * 1024 - 9216;     // (the kernel's static shared memory)
  const size_t k_stage = (size_t)2 * (s.nfull * FA_K_CHUNK + s.npiece * (FA_K_CHUNK / 4)), v_stage = (size_t)2 * s.ND * 128;
  size_t smem = 1024 + (size_t)2 * (s.nfull * FA_Q_CHUNK + s.npiece * (FA_Q_CHUNK / 4)) + k_stage + v_stage + 2 * FA_P_BYTES;
  s.KS = s.VS = 1;
  if (smem + k_stage <= budget) {
    s.KS = 2;
    smem += k_stage;
  }
  if (smem + v_stage <= budget) {
    s.VS = 2;
    smem += v_stage;
  }
  if (smem > budget) {
    set_error("attention: %zu bytes of shared memory for head width %d", smem, d);
    return DRG_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int qtiles = (L + FA_BM - 1) / FA_BM, T = (S + FA_BN - 1) / FA_BN;
  const int requested = nsplit;
  nsplit = fa_effective_nsplit(B, H, L, S, requested);
  s.tiles_per_split = (T + nsplit - 1) / nsplit;
  s.nsplit = nsplit;
  if (nsplit > 1) {
    DRG_CHECK_ARG(workspace != nullptr && workspace_bytes >= drg_attention_workspace_bytes(B, H, L, S, d, requested) &&
                      (((uintptr_t)workspace) & 15u) == 0,
                  "split keys need the workspace of drg_attention_workspace_bytes (16-byte aligned)");
    s.part = reinterpret_cast<float*>(workspace);
    s.part_ml = reinterpret_cast<float2*>(s.part + (size_t)nsplit * B * L * H * d);
  }
  const dim3 grid((unsigned)qtiles, (unsigned)BH, (unsigned)nsplit);
  if (BH > 65535) {
    set_error("attention: batch * heads = %d exceeds the grid", BH);
    return DRG_ERR_UNSUPPORTED;
  }
  bool launched = false;
#define FA_CASE(NF, NP)                                                                                                   \
  if (s.nfull == NF && s.npiece == NP) {                                                                                  \
    DRG_CUDA((cudaFuncSetAttribute(flash_attn_kernel<NF, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))); \
    flash_attn_kernel<NF, NP><<<grid, FA_THREADS, smem, st>>>(tQ, tK, tV, tQp, tKp, s);                                  \
    launched = true;                                                                                                      \
  }
  FA_CASE(0, 1) FA_CASE(0, 2) FA_CASE(0, 3) FA_CASE(1, 0) FA_CASE(1, 1) FA_CASE(1, 2) FA_CASE(1, 3)
  FA_CASE(2, 0) FA_CASE(2, 1) FA_CASE(2, 2) FA_CASE(2, 3)
#undef FA_CASE
  if (!launched) {
    set_error("attention: no kernel for head width %d", d);
    return DRG_ERR_UNSUPPORTED;
  }
  if (nsplit > 1) {
    DRG_LAUNCH_CHECK();
    const size_t total = (size_t)B * L * H * (d / 4);
    attn_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(s.part, s.part_ml, nsplit, B, H, L, d, out);
  }
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

extern "C" int drg_prep_vt_split16(const float* V, int B, int H, int S, int d, void* Vt16, void* workspace, void* stream) {
  DRG_CHECK_ARG(V && Vt16 && workspace, "V / Vt16 / workspace must be non-null");
  DRG_CHECK_ARG(B >= 1 && H >= 1 && S >= 1 && d >= 4 && d % 4 == 0 && B * H <= 65535, "B, H, S >= 1, d a multiple of 4, B * H <= 65535");
  DRG_CHECK_ARG(((((uintptr_t)Vt16) | ((uintptr_t)V)) & 15u) == 0 && d <= 1024, "V / Vt16 must be 16-byte aligned, d <= 1024");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* cmax = reinterpret_cast<unsigned int*>(workspace);     // [B * H, d]
  DRG_CUDA(cudaMemsetAsync(cmax, 0, (size_t)B * H * d * sizeof(unsigned int), st));
  vt_colmax_kernel<<<dim3((unsigned)((S + 127) / 128), (unsigned)(B * H)), 256, 0, st>>>(V, H, S, d, cmax);
  DRG_LAUNCH_CHECK();
  const int kcS = split16_kc(S);
  vt_split_kernel<<<dim3((unsigned)(kcS / 64), (unsigned)((d + 31) / 32), (unsigned)(B * H)), 256, 0, st>>>(
      V, cmax, H, S, d, reinterpret_cast<unsigned short*>(Vt16));
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}
