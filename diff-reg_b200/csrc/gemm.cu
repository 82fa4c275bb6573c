// Feature-similarity contraction  C[b] = alpha * A[b] . B[b]^T  on the 5th-gen tensor cores.  sm_100a.
//
// Replaces torch.einsum("bsc,btc->bst", src_feats, tgt_feats) (Diff-Reg-4dmatch/models/matching.py:149,161;
// Diff-Reg-2d3d/experiments/<exp>/matching.py:110,122) and the nn.Linear projections in front of it
// (matching.py:127-128).  See include/diffreg_b200.h.
//
// Both operands are K-major fp32 ([rows, K], K contiguous).  The kernel is a persistent,
// warp-specialised tcgen05 pipeline:
//   warp 0      TMA producer: cp.async.bulk.tensor (128B swizzle) of a 128 x 32 A block and a
//               BN x 32 B block per k-step into a ring of shared-memory stages (mbarrier full/empty)
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8),
//               four per k-step, accumulating in TMEM; tcgen05.commit releases the smem stage and,
//               after the last k-step, publishes the accumulator to the epilogue
//   warps 2..9  epilogue: tcgen05.ld (32 lanes x 32 columns per warp), scale, swizzled st.shared,
//               TMA tensor store of 32 x 32 boxes (clipped at the matrix edge by the TMA unit)
// Two TMEM accumulator stages (2 x BN columns) let the MMA of tile t+1 overlap the epilogue of t.
//
// fp32 parity: kind::tf32 keeps 10 mantissa bits of each operand.  The fp32-accurate mode of this kernel (SPLIT) takes
// 16-bit SPLIT operands instead (features.cu): every row is scaled by a power of two 2^e that puts its maximum into
// [2^14, 2^15), then x 2^e = hi + lo with hi = fp16(x 2^e) (11 significant bits, as many as tf32 keeps) and
// lo = fp16(x 2^e - hi); the kernel accumulates
//     A_lo.B_hi + A_hi.B_lo + A_hi.B_hi          (kind::f16, fp32 accumulation in TMEM; dropped: A_lo.B_lo ~ 2^-22)
// and the epilogue multiplies the row and column scales back (exact: powers of two).  fp16 x bf16 products (a bf16 lo would
// need no scaling) are not an option: kind::f16 wants ONE format for A and B -- mixed descriptors raise an illegal-instruction
// fault (measured, tools/probes/umma_fmt_probe.cu).  Against round 1's 3xTF32 (the same three terms on fp32 operands with
// kind::tf32) every operand byte and every MMA covers twice as many columns: the operand traffic from L2 -- what bounds this
// kernel -- and the tensor-pipe time are halved, at the same accuracy (measured against fp64 in tests/test_gemm_gpu.py).
#include "umma.cuh"

namespace drg {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 32;  // fp32 elements per k-step: 128 bytes = one swizzle-atom row
constexpr int GEMM_EPI_WARPS = 8;   // two warps per TMEM lane quadrant (they split the tile's 32-column chunks)
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * GEMM_BK * 4;  // 16 KB
constexpr int GEMM_OUT_BOX_BYTES = 32 * 32 * 4;            // one epilogue box: 32 rows x 32 columns
constexpr int GEMM_MAX_STAGES = 8;

struct GemmShape {
  int N, M, K, batch;
  float alpha;
  int nstage;
  int split_tma;     // 1: the split epilogue stages hi / lo boxes in shared memory and writes them with TMA stores (tmS)
  int direct_store;  // 1: C rows not 16-byte aligned (M % 4 != 0) -> scalar st.global from the staging tile; 0: st.global.v4
  int kc;            // SPLIT kernels: 16-bit columns of ONE operand segment (the operand rows are [seg0 | seg1 | tail], pitch
                     //   2 * kc + 8; kc = K rounded up to 64, the padding holds zeros), a multiple of 64 = one 128-byte k-step
  const unsigned short* A16;  // SPLIT kernels: the operands themselves, for the row tails (1 / scale, norm) the epilogue reads
  const unsigned short* B16;
  int wide_tiles;    // tail balancing: tiles [0, wide_tiles) are BN wide; the remaining (tiles - wide_tiles) tiles -- the
                     //   last, partial wave -- are processed as twice as many HALF-width tiles so that one round of the
                     //   persistent grid finishes them in half a tile time (wide_tiles == tiles: no split)
  float* C;          // may be NULL when only the split output is wanted
  const float* bias; // optional [M]: C[., j] += bias[j] (the bias of an nn.Linear whose weight is the right operand)
  // optional fused operand preparation of the NEXT GEMM (projection -> similarity): split_out [N, 2 * split_kc + 8] (16-bit)
  // receives scale*alpha*acc as a split operand: [lo | hi | tail] for rows < split_rows0 (left operand), [hi | lo | tail] for the
  // others (right operand); columns [M, split_kc) of each segment are the caller's zero padding.  The row scale of an output
  // row comes from a bound instead of its maximum (a row is spread over several tiles): |y_ij| <= ||x_i|| max_j ||W_j||.
  unsigned short* split_out;
  int split_kc;
  int split_rows0;
  float split_scale;
};

// SPLIT3 (the 16-bit split mode): the operands are A' = [A_lo | A_hi | tail], B' = [B_hi | B_lo | tail] (fp16), each segment
// kc 16-bit columns.  A stage holds the four tiles A_lo, A_hi, B_hi, B_lo of one 64-column k-step (128-byte rows, the same
// tile geometry and swizzle as the fp32 mode) and the MMA warp issues lo.hi, hi.lo, hi.hi (kind::f16, K = 16, four each
// per k-step) from them.
// MC (launched as clusters of two CTAs): the two CTAs of a cluster work on the two 128-row tiles (2p, 2p + 1) of the SAME
// column block, so they need the same B tile: each loads half of its rows and multicasts them into both CTAs' shared memory
// (cp.async.bulk.tensor ... .multicast::cluster).  A third less operand traffic from L2 per tile -- the kernel runs at the
// L2 -> SM rate (ncu: 244 MB over the crossbar in 38 us = 6.3 TB/s against ~8-9.6 TB/s for a pure L2 stream,
// profiles/r1_stream_l2_microbench.log, while also writing the output).  A stage is refilled by BOTH CTAs, so it is free only
// when both have consumed it: the MMA warp's tcgen05.commit arrives on the stage's "empty" barrier of both CTAs (count 2).
// MODE 2 (PAIR, clusters of two CTAs as well): the pair runs ONE tcgen05.mma.cta_group::2 of 256 x BN per step -- each CTA holds
// its 128 rows of A and only HALF of the B tile's rows (the tensor core reads the other half out of the peer's shared memory), so
// a CTA takes in 64 KB instead of 96 KB per k-step and three stages fit instead of two.  The leader CTA's MMA warp issues for
// the pair; both CTAs' TMA loads signal the leader's "full" barrier; its commits arrive on both CTAs' "empty" / "accumulator
// full" barriers, and both CTAs' epilogue warps arrive on the leader's "accumulator empty" barrier.
template <int BN, bool SPLIT3, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmS,
                     const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBq, const GemmShape s) {
  constexpr bool MC = MODE == 1, PAIR = MODE == 2, CL = MODE != 0;
  static_assert(!PAIR || (SPLIT3 && BN == 256), "pair mode: split operands, 256-wide tiles");
  constexpr int B_TILE_BYTES = (PAIR ? BN / 2 : BN) * GEMM_BK * 4;   // PAIR: this CTA's half of the tile's rows
  constexpr int A_STAGE_BYTES = (SPLIT3 ? 2 : 1) * GEMM_A_STAGE_BYTES;
  constexpr int B_STAGE_BYTES = (SPLIT3 ? 2 : 1) * B_TILE_BYTES;
  constexpr int TMEM_COLS = 2 * BN;  // 128, 256 or 512: a power of two >= 32
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN");
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[GEMM_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[GEMM_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nstage = s.nstage;
  // 1024-byte aligned carve-up (swizzle atoms)
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* aligned = smem_dyn + (base - smem_u32(smem_dyn));
  uint8_t* sA = aligned;
  uint8_t* sB = sA + (size_t)nstage * A_STAGE_BYTES;
  uint8_t* sOut = sB + (size_t)nstage * B_STAGE_BYTES;  // [4 warps][2][4 KB]

  // MC: the units below are PAIRS of vertically adjacent tiles (the host guarantees an even number of row blocks); this CTA
  // takes row block 2 * mb + rank of a pair
  const uint32_t crank = CL ? cluster_ctarank() : 0u;
  const int first_unit = CL ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = CL ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tiles_m = CL ? ((s.N + GEMM_BM - 1) / GEMM_BM) / 2 : (s.N + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = (s.M + BN - 1) / BN;
  const int tiles = s.batch * tiles_m * tiles_n;
  constexpr int KSTEP = SPLIT3 ? 64 : GEMM_BK;   // operand columns per k-step: 128 bytes either way
  const int kblocks = SPLIT3 ? s.kc / KSTEP : (s.K + GEMM_BK - 1) / GEMM_BK;
  // work units of the persistent loop: the wide tiles, then two half-width units per remaining tile
  const int units = s.wide_tiles + 2 * (tiles - s.wide_tiles);
  // unit -> (batch, row block, first column, narrow?)
  auto decode = [&](int unit, int& b, int& mb, int& n0, bool& narrow) {
    narrow = unit >= s.wide_tiles;
    const int tile = narrow ? s.wide_tiles + ((unit - s.wide_tiles) >> 1) : unit;
    b = tile / (tiles_m * tiles_n);
    const int rem = tile - b * tiles_m * tiles_n;
    mb = rem / tiles_n;
    const int nb = rem - mb * tiles_n;
    n0 = nb * BN + ((narrow && ((unit - s.wide_tiles) & 1)) ? BN / 2 : 0);
    if (CL) mb = 2 * mb + (int)crank;
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (CL || s.wide_tiles < tiles) prefetch_tmap(&tmBh);
    if (PAIR && s.wide_tiles < tiles) prefetch_tmap(&tmBq);
    if (s.split_tma) prefetch_tmap(&tmS);
    for (int i = 0; i < nstage; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], MC ? 2 : 1);   // MC: released by the MMA warps of both CTAs (PAIR: one commit of the leader, multicast)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], PAIR ? 2 * GEMM_EPI_WARPS : GEMM_EPI_WARPS);  // one arrival per epilogue warp (PAIR: of both CTAs, on the leader's)
    }
    fence_mbar_init();
  }
  if (PAIR) {
    __syncthreads();
    cluster_sync_all();           // both CTAs are resident and their barriers initialised before the pair allocates TMEM together
    if (warp == 1) tmem_alloc_2sm<TMEM_COLS>(&tmem_base_slot);
  } else if (warp == 1) {
    tmem_alloc<TMEM_COLS>(&tmem_base_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL) cluster_sync_all();   // the peer's barriers are initialised before anything of ours can reach them
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = first_unit; unit < units; unit += unit_stride) {
        int b, mb, n0;
        bool narrow;
        decode(unit, b, mb, n0, narrow);
        const CUtensorMap* tb = narrow ? &tmBh : &tmB;   // half-height boxes for the half-width units
        const uint32_t bytes = (uint32_t)(A_STAGE_BYTES + (narrow ? B_STAGE_BYTES / 2 : B_STAGE_BYTES));
        for (int k = 0; k < kblocks; ++k) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* a_dst = sA + (size_t)stage * A_STAGE_BYTES;
          uint8_t* b_dst = sB + (size_t)stage * B_STAGE_BYTES;
          if (PAIR) {
            // both CTAs' tiles are accounted on the LEADER's barrier (the leader expects the bytes of both)
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * bytes);
            const uint32_t lbar = map_to_cta(&full_bar[stage], 0u);
            const int rows_b = narrow ? BN / 4 : BN / 2;          // this CTA's half of the (narrow: half-width) B tile
            const CUtensorMap* tbp = narrow ? &tmBq : &tmBh;
            const int r0 = n0 + (int)crank * rows_b;
            tma_load_3d_2sm(a_dst, &tmA, k * KSTEP, mb * GEMM_BM, b, lbar);                                   // A_lo
            tma_load_3d_2sm(a_dst + GEMM_A_STAGE_BYTES, &tmA, s.kc + k * KSTEP, mb * GEMM_BM, b, lbar);       // A_hi
            tma_load_3d_2sm(b_dst, tbp, k * KSTEP, r0, b, lbar);                                             // B_hi (half)
            tma_load_3d_2sm(b_dst + B_TILE_BYTES, tbp, s.kc + k * KSTEP, r0, b, lbar);                       // B_lo (half)
            if (++stage == nstage) {
              stage = 0;
              phase ^= 1u;
            }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          tma_load_3d(a_dst, &tmA, k * KSTEP, mb * GEMM_BM, b, &full_bar[stage]);   // SPLIT3: A_lo
          if (SPLIT3) tma_load_3d(a_dst + GEMM_A_STAGE_BYTES, &tmA, s.kc + k * KSTEP, mb * GEMM_BM, b, &full_bar[stage]);  // A_hi
          if (MC && !narrow) {
            // this CTA's half of the B tile's rows, into both CTAs (tmBh boxes are BN / 2 rows high)
            const int r0 = n0 + (int)crank * (BN / 2);
            uint8_t* dst = b_dst + (size_t)crank * (B_TILE_BYTES / 2);
            tma_load_3d_mc(dst, &tmBh, k * KSTEP, r0, b, &full_bar[stage], (uint16_t)3);                               // B_hi
            if (SPLIT3) tma_load_3d_mc(dst + B_TILE_BYTES, &tmBh, s.kc + k * KSTEP, r0, b, &full_bar[stage], (uint16_t)3);  // B_lo
          } else {
            tma_load_3d(b_dst, tb, k * KSTEP, n0, b, &full_bar[stage]);               // SPLIT3: B_hi
            if (SPLIT3) tma_load_3d(b_dst + B_TILE_BYTES, tb, s.kc + k * KSTEP, n0, b, &full_bar[stage]);  // B_lo
          }
          if (++stage == nstage) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (PAIR: the leader CTA's, for both) =====================
    // the whole warp runs the loop, converged; one ELECTED lane per instruction (umma.cuh)
    if (!(PAIR && crank != 0)) {
      constexpr uint32_t idesc_wide = make_idesc_tf32(GEMM_BM, BN);
      constexpr uint32_t idesc_half = make_idesc_tf32(GEMM_BM, BN / 2);
      // 16-bit split mode: fp16 x fp16 for all three terms
      constexpr uint32_t id16[2] = {make_idesc_f16(PAIR ? 2 * GEMM_BM : GEMM_BM, BN, FMT_F16, FMT_F16),
                                    make_idesc_f16(PAIR ? 2 * GEMM_BM : GEMM_BM, BN / 2, FMT_F16, FMT_F16)};
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int unit = first_unit; unit < units; unit += unit_stride) {
        const int nw = unit >= s.wide_tiles ? 1 : 0;
        const uint32_t idesc = nw ? idesc_half : idesc_wide;
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1u);
        __syncwarp();
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int k = 0; k < kblocks; ++k) {
          mbar_wait(&full_bar[stage], phase);
          __syncwarp();
          tcgen05_fence_after();
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA + (size_t)stage * A_STAGE_BYTES));
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(sB + (size_t)stage * B_STAGE_BYTES));
          if (SPLIT3) {
            const uint64_t ah_desc = make_smem_desc_sw128(smem_u32(sA + (size_t)stage * A_STAGE_BYTES + GEMM_A_STAGE_BYTES));
            const uint64_t bl_desc = make_smem_desc_sw128(smem_u32(sB + (size_t)stage * B_STAGE_BYTES + B_TILE_BYTES));
            // (the small correction terms first: the accumulator is rounded at every step, so they are added while it is small)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {     // 4 x 16 columns; +32 bytes along K inside the swizzle atom = +2 in (addr >> 4)
              const uint64_t o = (uint64_t)(2 * kk);
              if (PAIR) {
                umma_f16_2sm_e(d_tmem, a_desc + o, b_desc + o, id16[nw], (uint32_t)((k | kk) != 0));  // lo . hi
                umma_f16_2sm_e(d_tmem, ah_desc + o, bl_desc + o, id16[nw], 1u);                        // hi . lo
                umma_f16_2sm_e(d_tmem, ah_desc + o, b_desc + o, id16[nw], 1u);                         // hi . hi
              } else {
                umma_f16_e(d_tmem, a_desc + o, b_desc + o, id16[nw], (uint32_t)((k | kk) != 0));  // lo . hi
                umma_f16_e(d_tmem, ah_desc + o, bl_desc + o, id16[nw], 1u);                        // hi . lo
                umma_f16_e(d_tmem, ah_desc + o, b_desc + o, id16[nw], 1u);                         // hi . hi
              }
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < GEMM_BK / 8; ++kk) {
              // advance 8 tf32 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
              umma_tf32_e(d_tmem, a_desc + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc, (uint32_t)((k | kk) != 0));
            }
          }
          if (PAIR) umma_commit_2sm_mc_e(&empty_bar[stage], (uint16_t)3);   // both CTAs' producers may refill their halves
          else if (MC) umma_commit_mc_e(&empty_bar[stage], (uint16_t)3);   // ... in both CTAs: either may refill (its half of) the stage
          else umma_commit_e(&empty_bar[stage]);  // smem stage reusable once these MMAs have read it
          if (++stage == nstage) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (PAIR) umma_commit_2sm_mc_e(&tmem_full_bar[as], (uint16_t)3);  // accumulator complete, in both CTAs
        else umma_commit_e(&tmem_full_bar[as]);  // accumulator complete
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Eight warps: a warp may read the TMEM lane quadrant (warp % 4) only, so two warps share a quadrant and take alternate
    // 32-column chunks of the tile (measured: no faster than four warps on the 4096^2 GEMM -- the epilogue does not set the pace
    // there -- but 10 % on the one-tile-per-CTA projection GEMM).
    const int ew = warp - 2;
    const int q = warp & 3;   // TMEM lane quadrant this warp may read: lanes 32q .. 32q+31
    const int half = ew >> 2; // which of the quadrant's two warps
    uint8_t* stg = sOut + (size_t)ew * GEMM_OUT_BOX_BYTES;   // one 4 KB staging box per warp
    int as = 0;
    uint32_t aphase = 0;
    // SPLIT3: the operands' row tails.  op_pitch in 16-bit units; tail floats: [0] = 1 / row scale, [1] = row norm
    const int op_pitch = 2 * s.kc + 8;
    auto tail_of = [&](const unsigned short* op, size_t row) { return reinterpret_cast<const float*>(op + row * op_pitch + 2 * s.kc); };
    float wmax = 0.f;   // split epilogue: the largest row norm of the right operand (the weight)
    if (SPLIT3 && s.split_out) {
      for (int j = lane; j < s.M; j += 32) wmax = fmaxf(wmax, tail_of(s.B16, (size_t)j)[1]);
      wmax = warp_max(wmax);
    }
    for (int unit = first_unit; unit < units; unit += unit_stride) {
      int b, mb, n0;
      bool narrow;
      decode(unit, b, mb, n0, narrow);
      const int row0 = mb * GEMM_BM + q * 32;
      const int nchunks = narrow ? BN / 64 : BN / 32;
      // this lane's accumulator row: 1 / scale of the left operand's row (and, for the split epilogue, its norm)
      float rs_a = 1.f, nx_a = 0.f;
      if (SPLIT3 && row0 + lane < s.N) {
        const float* t = tail_of(s.A16, (size_t)b * s.N + row0 + lane);
        rs_a = t[0];
        nx_a = t[1];
      }
      // 1 / scale of the right operand's rows = this tile's output columns (lane = column of every 32-column chunk), fetched
      // while the MMA warp still works on the tile: a dependent global load per box would sit on the epilogue's critical path
      float cscale[BN / 32];
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        const int col = n0 + c * 32 + lane;
        cscale[c] = (SPLIT3 && c < nchunks && col < s.M) ? tail_of(s.B16, (size_t)b * s.M + col)[0] : 1.f;
      }
      mbar_wait(&tmem_full_bar[as], aphase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = half; c < nchunks; c += 2) {
        const int col0 = n0 + c * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), r);
        float rs_b = cscale[0];   // 1 / scale of output column col0 + lane (register select: the chunk loop is not unrolled)
#pragma unroll
        for (int t = 1; t < BN / 32; ++t) rs_b = (c == t) ? cscale[t] : rs_b;
        tmem_wait_ld();
        if (row0 >= s.N || col0 >= s.M) continue;  // whole box outside the matrix (warp-uniform)
        if (SPLIT3) {
          // undo the operands' power-of-two row scales (exact): acc_ij / (scale_i scale_j)
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * __shfl_sync(0xffffffffu, rs_b, j));
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * rs_a);
        }
        // split epilogue: this row's output scale 2^e from the bound |alpha scale| ||x_i|| max_j ||W_j|| (one octave of margin)
        float sc_out = 1.f;
        if (SPLIT3 && s.split_out) {
          const float bound = fabsf(s.alpha * s.split_scale) * nx_a * wmax;
          int e = 0;
          if (bound > 0.f && bound <= 3.0e38f) e = min(max(13 - ilogbf(bound), -126), 126);
          sc_out = __int_as_float((e + 127) << 23);
          if (col0 == 0 && row0 + lane < s.N)
            *reinterpret_cast<float4*>(s.split_out + (size_t)(row0 + lane) * (2 * s.split_kc + 8) + 2 * s.split_kc) =
                make_float4(__int_as_float((127 - e) << 23), 0.f, 0.f, 0.f);
        }
        if (s.split_out && s.split_tma) {
          // fused drg_prep_operand(split=1) of the projected features: the fp16 hi and lo boxes (32 x 32 x 2 bytes, plain
          // row-major) are staged in this warp's two buffers and leave as two TMA tensor stores into the split operand
          const float sc = s.alpha * s.split_scale * sc_out;
          uint8_t* hi_box = stg;
          uint8_t* lo_box = stg + GEMM_OUT_BOX_BYTES / 2;   // (16-bit boxes are 2 KB each)
          if (lane == 0) bulk_wait_group_read<0>();  // the previous chunk's stores have finished reading both buffers
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {              // lane = row: 32 columns = 4 chunks of 8 values (16 bytes of 16-bit)
            unsigned short h[8], l[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split16(__uint_as_float(r[8 * j + e]) * sc, h[e], l[e]);
            *reinterpret_cast<uint4*>(hi_box + lane * 64 + j * 16) = make_uint4(pack16(h[0], h[1]), pack16(h[2], h[3]), pack16(h[4], h[5]), pack16(h[6], h[7]));
            *reinterpret_cast<uint4*>(lo_box + lane * 64 + j * 16) = make_uint4(pack16(l[0], l[1]), pack16(l[2], l[3]), pack16(l[4], l[5]), pack16(l[6], l[7]));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const bool left = row0 < s.split_rows0;  // split_rows0 % 32 == 0 in this mode: a box never straddles the boundary
            tma_store_3d(&tmS, left ? lo_box : hi_box, col0, row0, 0);
            tma_store_3d(&tmS, left ? hi_box : lo_box, s.split_kc + col0, row0, 0);
            bulk_commit_group();
          }
          continue;
        }
        if (s.split_out) {
          // the same without TMA (left / right boundary not box-aligned): lane = row, 32 consecutive columns
          const int row = row0 + lane;
          if (row < s.N) {
            const bool left = row < s.split_rows0;
            unsigned short* o = s.split_out + (size_t)row * (2 * s.split_kc + 8) + col0;
            const float sc = s.alpha * s.split_scale * sc_out;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (col0 + 4 * j < s.M) {  // M % 4 == 0 in this mode
                unsigned short h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split16(__uint_as_float(r[4 * j + e]) * sc, h[e], l[e]);
                const uint2 hv = make_uint2(pack16(h[0], h[1]), pack16(h[2], h[3])), lv = make_uint2(pack16(l[0], l[1]), pack16(l[2], l[3]));
                *reinterpret_cast<uint2*>(o + 4 * j) = left ? lv : hv;
                *reinterpret_cast<uint2*>(o + s.split_kc + 4 * j) = left ? hv : lv;
              }
            }
          }
          if (s.C == nullptr) continue;
        }
        // The box goes through this warp's swizzled staging buffer (row = lane, 16-byte chunk j at chunk j ^ (lane & 7)) and
        // leaves with plain vector stores: four rows x 128 contiguous bytes per instruction.  (Round 1 / early round 2 used TMA
        // tensor stores of 32 x 32 boxes here; measured, they drain at ~7.5 B/clk/SM -- 17 cycles per 128-byte box row -- which
        // made the OUTPUT path, not the operands or the MMAs, the bound of the 4096^2 GEMM: 1.8 TB/s for the 64 MB result,
        // where the Sinkhorn's final pass writes the same matrix with st.global.v4 at 5.6 TB/s.)
        uint8_t* box = stg;
        __syncwarp();   // the previous box has been read out of the buffer by every lane
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          o.x = __uint_as_float(r[4 * j + 0]) * s.alpha;
          o.y = __uint_as_float(r[4 * j + 1]) * s.alpha;
          o.z = __uint_as_float(r[4 * j + 2]) * s.alpha;
          o.w = __uint_as_float(r[4 * j + 3]) * s.alpha;
          *reinterpret_cast<float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
        }
        __syncwarp();
        float* Cb = s.C + (size_t)b * s.N * s.M;
        if (!s.direct_store) {
          // 16-byte aligned rows (M % 4 == 0): lanes 8 i .. 8 i + 7 write the eight 16-byte chunks of one row
          const int chunk = lane & 7;
          const int col = col0 + 4 * chunk;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3);
            const int row = row0 + rr;
            float4 v = *reinterpret_cast<const float4*>(box + rr * 128 + ((chunk ^ (rr & 7)) << 4));
            if (row < s.N && col < s.M) {
              if (s.bias) {
                const float4 bq = *reinterpret_cast<const float4*>(s.bias + col);
                v.x += bq.x; v.y += bq.y; v.z += bq.z; v.w += bq.w;
              }
              *reinterpret_cast<float4*>(Cb + (size_t)row * s.M + col) = v;
            }
          }
        } else {
          // coalesced scalar stores: lane = column, loop over the 32 rows of the box
          const int col = col0 + lane;
          for (int rr = 0; rr < 32; ++rr) {
            const int row = row0 + rr;
            if (row < s.N && col < s.M) {
              const float val = *reinterpret_cast<const float*>(box + rr * 128 + (((lane >> 2) ^ (rr & 7)) << 4) + ((lane & 3) << 2));
              Cb[(size_t)row * s.M + col] = s.bias ? val + s.bias[col] : val;
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(map_to_cta(&tmem_empty_bar[as], 0u));   // the pair's MMA issuer lives in the leader CTA
        else mbar_arrive(&tmem_empty_bar[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
    if (s.split_tma && lane == 0) bulk_wait_group<0>();  // smem must outlive the last TMA stores (split epilogue)
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CL) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  tcgen05_fence_after();
  if (warp == 1) {
    if (PAIR) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---- host side --------------------------------------------------------------------------
template <int BN, bool SPLIT3, int MODE>
static int launch_cluster(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const CUtensorMap& tS,
                          const CUtensorMap& tBh, const CUtensorMap& tBq, GemmShape s, size_t smem, cudaStream_t st) {
  const int tiles_m = (s.N + GEMM_BM - 1) / GEMM_BM, tiles_n = (s.M + BN - 1) / BN;
  DRG_CUDA((cudaFuncSetAttribute(gemm_tf32_kernel<BN, SPLIT3, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  const long long pairs = (long long)s.batch * (tiles_m / 2) * tiles_n;
  const int nclusters = (int)(pairs < NUM_SMS / 2 ? pairs : NUM_SMS / 2);
  const int rem = (int)(pairs % nclusters);
  s.wide_tiles = (int)pairs;   // (counted in pairs of tiles in the cluster modes)
  if (pairs > nclusters && rem > 0 && 2 * rem <= nclusters) s.wide_tiles = (int)pairs - rem;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * nclusters);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
  {
    ProfScope prof_scope(PROF_GEMM, st);
    e = cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<BN, SPLIT3, MODE>, tA, tB, tC, tS, tBh, tBq, s);
  }
  if (e != cudaSuccess) {   // a device / partition that cannot place two-CTA clusters: the caller falls back to single CTAs
    (void)cudaGetLastError();
    return -1;
  }
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

template <int BN, bool SPLIT3>
static int launch_gemm(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const CUtensorMap& tS,
                       const CUtensorMap& tBh, const CUtensorMap& tBq, GemmShape s, cudaStream_t st) {
  const int tiles_m = (s.N + GEMM_BM - 1) / GEMM_BM, tiles_n = (s.M + BN - 1) / BN;
  // Two-CTA clusters (kernel template MODE 1 / 2): split operands, wide tiles, an even number of row blocks, a B operand worth
  // sharing (a weight of a few hundred rows is not) and no split epilogue.  256-wide tiles: the pair runs cta_group::2 MMAs
  // (half of the B tile per CTA); 128-wide: each CTA loads half of the B tile and multicasts it.
  const bool cluster_ok = SPLIT3 && BN >= 128 && tiles_m >= 2 && (tiles_m % 2) == 0 && s.split_out == nullptr && s.M >= 2 * BN;
  const bool pair = cluster_ok && BN == 256;
  const int stage_bytes = (SPLIT3 ? 2 : 1) * (GEMM_A_STAGE_BYTES + (pair ? BN / 2 : BN) * GEMM_BK * 4);
  const size_t out_bytes = (size_t)GEMM_EPI_WARPS * GEMM_OUT_BOX_BYTES;
  const size_t budget = 227 * 1024 - 1024 /*alignment slack*/ - out_bytes - 256 /*static*/;
  int nstage = (int)(budget / stage_bytes);
  if (nstage > GEMM_MAX_STAGES) nstage = GEMM_MAX_STAGES;
  const int kblocks = SPLIT3 ? s.kc / 64 : (s.K + GEMM_BK - 1) / GEMM_BK;
  if (nstage > 2 * kblocks) nstage = 2 * kblocks;  // no point in more stages than two tiles of k-steps
  if (nstage < 2) nstage = 2;
  s.nstage = nstage;
  const size_t smem = 1024 + (size_t)nstage * stage_bytes + out_bytes;
  if constexpr (SPLIT3 && BN == 256) {
    if (pair) {
      const int rc = launch_cluster<BN, SPLIT3, 2>(tA, tB, tC, tS, tBh, tBq, s, smem, st);
      if (rc != -1) return rc;
      // cluster launch refused: single-CTA kernel with its own (two-stage) geometry
      GemmShape s1 = s;
      const int sb1 = 2 * (GEMM_A_STAGE_BYTES + BN * GEMM_BK * 4);
      int ns1 = (int)(budget / sb1);
      if (ns1 > 2 * kblocks) ns1 = 2 * kblocks;
      if (ns1 < 2) ns1 = 2;
      s1.nstage = ns1;
      const size_t smem1 = 1024 + (size_t)ns1 * sb1 + out_bytes;
      DRG_CUDA((cudaFuncSetAttribute(gemm_tf32_kernel<BN, SPLIT3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1)));
      const long long tiles1 = (long long)s.batch * tiles_m * tiles_n;
      const int grid1 = (int)(tiles1 < NUM_SMS ? tiles1 : NUM_SMS);
      const int rem1 = (int)(tiles1 % grid1);
      s1.wide_tiles = (int)tiles1;
      if (tiles1 > grid1 && rem1 > 0 && 2 * rem1 <= grid1) s1.wide_tiles = (int)tiles1 - rem1;
      {
        ProfScope prof_scope(PROF_GEMM, st);
        gemm_tf32_kernel<BN, SPLIT3, 0><<<grid1, GEMM_THREADS, smem1, st>>>(tA, tB, tC, tS, tBh, tBq, s1);
      }
      DRG_LAUNCH_CHECK();
      return DRG_OK;
    }
  }
  if constexpr (SPLIT3 && BN == 128) {
    if (cluster_ok) {
      const int rc = launch_cluster<BN, SPLIT3, 1>(tA, tB, tC, tS, tBh, tBq, s, smem, st);
      if (rc != -1) return rc;   // (same stage geometry as the single-CTA kernel: fall through)
    }
  }
  DRG_CUDA((cudaFuncSetAttribute(gemm_tf32_kernel<BN, SPLIT3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  const long long tiles = (long long)s.batch * tiles_m * tiles_n;
  const int grid = (int)(tiles < NUM_SMS ? tiles : NUM_SMS);
  // Tail balancing: with `rem` tiles left for the last, partial wave of the persistent grid, the wave costs a full tile
  // time although only rem of the grid's CTAs work (4096^2 with 128 x 256 tiles: 512 tiles = 3.46 waves of 148).  When
  // twice as many half-width tiles still fit into one wave, they finish it in half the time instead.
  const int rem = (int)(tiles % grid);
  s.wide_tiles = (int)tiles;
  if (BN >= 128 && tiles > grid && rem > 0 && 2 * rem <= grid) s.wide_tiles = (int)tiles - rem;
  {
    ProfScope prof_scope(PROF_GEMM, st);
    gemm_tf32_kernel<BN, SPLIT3, 0><<<grid, GEMM_THREADS, smem, st>>>(tA, tB, tC, tS, tBh, tBq, s);
  }
  DRG_LAUNCH_CHECK();
  return DRG_OK;
}

}  // namespace drg

using namespace drg;

// A, B: fp32 [batch, rows, K] (split16 == false) or 16-bit split operands [batch, rows, 2 * kc] with kc = K rounded up to 64
// (split16 == true; drg_prep_operand(split = 1) / the split epilogue write them).  split_out (optional, batch == 1): the 16-bit
// split operand [N, 2 * split_kc] of the next GEMM, split_kc = M rounded up to 64 (its padding columns must be zero already).
static int gemm_run(const void* A, const void* B, float* C, int batch, int N, int M, int K, float alpha, unsigned short* split_out,
                    int split_rows0, float split_scale, void* stream, bool split16 = false, const float* bias = nullptr) {
  DRG_CHECK_ARG(A && B && (C || split_out), "A/B and an output must be non-null");
  DRG_CHECK_ARG(batch >= 1 && N >= 1 && M >= 1 && K >= 1, "batch, N, M, K must be >= 1");
  if ((!split16 && K % 4 != 0) || ((uintptr_t)A & 15u) || ((uintptr_t)B & 15u)) {
    set_error("gemm: K must be a multiple of 4 and A, B 16-byte aligned (TMA row pitch); got K=%d", K);
    return DRG_ERR_UNSUPPORTED;
  }
  if (split_out && (M % 4 != 0 || batch != 1 || ((uintptr_t)split_out & 15u) || !split16)) {
    set_error("gemm: the split epilogue needs split operands, batch == 1, M %% 4 == 0 and a 16-byte aligned output");
    return DRG_ERR_UNSUPPORTED;
  }
  const int kc = (K + 63) & ~63;                 // 16-bit columns per operand segment (split mode)
  cudaStream_t st = (cudaStream_t)stream;
  // Tile width by a two-term cost model (cycles): the MMA time of the busiest SM, and the operand traffic through L2
  // (every tile re-reads its A and B k-blocks; measured on B200 the fabric delivers ~6.5 KB/clk to the SMs, which is
  // what bounds both the 4096^2 similarity GEMM and the skinny projection GEMM -- profiles/r1_gemm_ncu.txt).
  int BN = 64;
  {
    // k-steps of 128-byte operand rows: 32 fp32 columns, or 64 16-bit columns of each of the two segments
    const double kblocks = split16 ? (double)(kc / 64) : (double)((K + GEMM_BK - 1) / GEMM_BK);
    const double terms = split16 ? 3.0 : 1.0, tiles_per_step = split16 ? 2.0 : 1.0;
    double best = 1e300;
    for (int bn : {64, 128, 256}) {
      const double tiles = (double)batch * ((N + 127) / 128) * ((M + bn - 1) / bn);
      const double rounds = (double)((long long)((tiles + NUM_SMS - 1) / NUM_SMS));
      const double t_mma = rounds * kblocks * terms * 4.0 * (bn / 2.0);            // one 128 x bn MMA (K = 8 tf32 / 16 f16) = bn/2 clk
      const double t_l2 = tiles * kblocks * tiles_per_step * (double)((128 + bn) * 128) / 6500.0;
      const double t = (t_mma > t_l2 ? t_mma : t_l2) + rounds * 1500.0;            // + per-tile epilogue exposure
      if (t < best) {
        best = t;
        BN = bn;
      }
    }
    // the split epilogue without TMA stores (left / right boundary not box-aligned) writes row-strided 8-byte pieces from
    // four warps only: it wants many small tiles in flight rather than few wide ones
    const bool split_tma_early = split_out != nullptr && C == nullptr && (split_rows0 % 32) == 0;
    if (split_out && !split_tma_early) BN = 64;
  }
  GemmShape s{};
  s.N = N;
  s.M = M;
  s.K = K;
  s.batch = batch;
  s.alpha = alpha;
  s.C = C;
  s.bias = bias;
  s.split_out = split_out;
  s.split_kc = (M + 63) & ~63;
  s.split_rows0 = split_rows0;
  s.split_scale = split_scale;
  s.direct_store = (C == nullptr || M % 4 != 0 || ((uintptr_t)C & 15u)) ? 1 : 0;
  s.kc = split16 ? kc : 0;
  s.A16 = split16 ? reinterpret_cast<const unsigned short*>(A) : nullptr;
  s.B16 = split16 ? reinterpret_cast<const unsigned short*>(B) : nullptr;
  // split epilogue through TMA stores when only the split operand is wanted and the left / right boundary is box-aligned
  s.split_tma = (split_out != nullptr && C == nullptr && (split_rows0 % 32) == 0) ? 1 : 0;
  CUtensorMap tA, tB, tC, tS, tBh, tBq;
  const int eb = split16 ? 2 : 4, kcols = split16 ? 2 * kc + 8 : K, kbox = split16 ? 64 : GEMM_BK;   // (+8: the row tail)
  if (!make_tmap(&tA, A, batch, N, kcols, GEMM_BM, kbox, eb)) return DRG_ERR_CUDA;
  if (!make_tmap(&tB, B, batch, M, kcols, BN, kbox, eb)) return DRG_ERR_CUDA;
  if (BN >= 128) {
    if (!make_tmap(&tBh, B, batch, M, kcols, BN / 2, kbox, eb)) return DRG_ERR_CUDA;   // half-width units of the tail wave / half tiles of a pair
    if (!make_tmap(&tBq, B, batch, M, kcols, BN / 4, kbox, eb)) return DRG_ERR_CUDA;   // a pair's halves of half-width units
  } else {
    tBh = tB;
    tBq = tB;
  }
  if (!s.direct_store) {
    if (!make_tmap(&tC, C, batch, N, M, 32, 32)) return DRG_ERR_CUDA;
  } else {
    tC = tA;  // unused
  }
  if (s.split_tma) {
    if (!make_tmap(&tS, split_out, 1, N, 2 * s.split_kc + 8, 32, 32, 2, false)) return DRG_ERR_CUDA;
  } else {
    tS = tA;  // unused
  }
  if (split16) {
    switch (BN) {
      case 256: return launch_gemm<256, true>(tA, tB, tC, tS, tBh, tBq, s, st);
      case 128: return launch_gemm<128, true>(tA, tB, tC, tS, tBh, tBq, s, st);
      default: return launch_gemm<64, true>(tA, tB, tC, tS, tBh, tBq, s, st);
    }
  }
  switch (BN) {
    case 256: return launch_gemm<256, false>(tA, tB, tC, tS, tBh, tBq, s, st);
    case 128: return launch_gemm<128, false>(tA, tB, tC, tS, tBh, tBq, s, st);
    default: return launch_gemm<64, false>(tA, tB, tC, tS, tBh, tBq, s, st);
  }
}

extern "C" int drg_gemm_nt_tf32(const float* A, const float* B, float* C, int batch, int N, int M, int K, float alpha,
                                void* stream) {
  DRG_CHECK_ARG(C != nullptr, "C is null");
  return gemm_run(A, B, C, batch, N, M, K, alpha, nullptr, 0, 1.f, stream);
}

// fp32-accurate product from 16-bit split operands (drg_prep_operand(split = 1), patterns 0 / 1): A16 [batch, N, 2 kc],
// B16 [batch, M, 2 kc], kc = K rounded up to 64.
extern "C" int drg_gemm_nt_split16(const void* A16, const void* B16, float* C, int batch, int N, int M, int K, float alpha,
                                   void* stream) {
  DRG_CHECK_ARG(C != nullptr, "C is null");
  return gemm_run(A16, B16, C, batch, N, M, K, alpha, nullptr, 0, 1.f, stream, true);
}

// the same with a bias row added to every output row: C = alpha * A . B^T + bias (an nn.Linear with bias, B = its weight)
extern "C" int drg_gemm_nt_split16_bias(const void* A16, const void* B16, const float* bias, float* C, int batch, int N, int M, int K,
                                        float alpha, void* stream) {
  DRG_CHECK_ARG(C != nullptr, "C is null");
  DRG_CHECK_ARG(bias == nullptr || (((uintptr_t)bias) & 15u) == 0, "bias must be 16-byte aligned");
  return gemm_run(A16, B16, C, batch, N, M, K, alpha, nullptr, 0, 1.f, stream, true, bias);
}

extern "C" int drg_project_split16(const void* A16, const void* W16, int rows, int rows_left, int C_out, int K, float scale,
                                   float* plain_out, void* split_out16, void* stream) {
  DRG_CHECK_ARG(split_out16 != nullptr, "split_out is null");
  DRG_CHECK_ARG(rows_left >= 0 && rows_left <= rows, "rows_left out of range");
  return gemm_run(A16, W16, plain_out, 1, rows, C_out, K, 1.f, reinterpret_cast<unsigned short*>(split_out16), rows_left, scale, stream,
                  true);
}
