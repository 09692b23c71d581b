// (b)+(c) correlation GEMM on the 5th-generation tensor cores with a fused row (max, idx, runner-up)
// epilogue.  Replaces models/IPSRFunction.py:59 (conv_enc(ref) -> S[1,N,H,W] materialised) and
// util/MaxCoord.py:21-22 (zeros_like(S) + torch.max(S, 1)); S never leaves the SM.
//
// Arithmetic: S = R * Xn^T on power-of-two pre-scaled fp16 operands (ipsr_prep.cu), fp32 accumulation in TMEM,
// as a precision CASCADE:
//   pass 1 (PASSES = 1)  hi(R) * hi(X) over EVERY row: one tcgen05.mma pass at the full fp16 rate.  Its error is
//                        rigorously bounded per row by rerr[q] + ||R~[q]|| * max_p xerr[p] (+ accumulation),
//                        so a row whose top-2 gap exceeds twice that bound already has its fp32 arg-max;
//   pass 2 (PASSES = 3)  hi*lo + lo*hi + hi*hi (~fp32 accurate) over the compacted ambiguous rows only
//                        (ipsr_compact_rows; a few % of the rows), against the whole bank;
//   pass 3               rows still inside the error band of pass 2 are recomputed in exact fp32
//                        (ipsr_correlate_argmax_fp32).
//
// One CTA = ROWT x 128 query rows (one TMEM lane each) x a range of bank columns, walked in blocks of 128
// columns.  Warp roles (96 + 128*ROWT threads):
//   warp 0, last producers (alternate stages): bulk async copies (UBLKCP) of pre-swizzled 16 KiB tile images -> smem ring
//   warp 1       TMEM allocation + single-thread tcgen05.mma issue, commits free the ring slots
//   warps 2..    epilogue, 4 warps per row tile: tcgen05.ld of the finished accumulator (double-buffered in
//                TMEM) and the running (best, idx, second) per row in registers -- thread == row, no shuffles.
// A_RESIDENT: the R row tiles stay in shared memory for the whole CTA and only bank tiles stream; otherwise both
// operands stream per 64-channel block.  Pass 1 runs BN = 256 (one row tile, two adjacent bank tiles per stage, one
// 128 x 256 x 16 instruction per K step) whenever the shape allows it, else ROWT = 2 with BN = 128 (every streamed bank
// tile feeds 2 x 128 rows); the three-pass launches run ROWT = 1, BN = 128.
#include <stdlib.h>

#include "ipsr_common.cuh"

namespace ipsr {

constexpr int kBlockN = 128;

struct TcParams {
  const uint8_t* r_tiles;   // [B][KB][a_parts][RB] tile images; PASSES 1 reads part 0 (hi), PASSES 3 parts 0, 1 (hi, lo)
  int a_parts;
  const uint8_t* x_tiles;   // [B][KB][2][RB] tile images (hi, lo)
  const int* row_limit;     // optional [B]: only the first row_limit[b] rows of r_tiles are populated
  int B, KB, RB, N;
  int col_begin;            // first bank column (multiple of 128)
  int blocks_total;         // number of 128-column blocks in [col_begin, col_end)
  int psplit;
  int stages;
  int nprod;                // producer warps in use (1 or 2): issuing one bulk copy costs its thread ~650 cycles, two
                            // warps issue alternate stages
  float* part_best;
  int* part_idx;
  float* part_second;
  int* part_idx2;           // PASSES 3 only (optional): column of the runner-up and the third-best score, so that rows
  float* part_third;        // with exactly two candidates inside the error band need two exact dot products, not N
  float* s_dump;
  TcFinalize fin;           // fin.ind != NULL: decide per row right here (psplit == 1, PASSES == 3)
  int n_valid;              // bank columns >= n_valid are padding (patch maps whose position count is no multiple of 128):
                            // they never win
};

// CL = 2: clusters of two CTAs (adjacent row-tile groups of the same image and column range) share every streamed
// bank tile: each CTA fetches half of a stage and multicasts it into both shared memories, which halves the
// L2 -> SM traffic (the limiter of the small and of the compacted launches).  A ring slot is refilled only after the
// MMAs of BOTH CTAs have retired it (multicast commits onto both empty barriers).
// BN: bank columns per accumulator block = N of the tcgen05.mma instruction (128, or 256 with ROWT = 1 and one pass: two
// adjacent bank tiles per stage, one 128 x 256 x 16 instruction instead of two 128 x 128 x 16 ones).
// A_TMEM: the (resident) row tile lives in TENSOR memory instead of shared memory -- the epilogue warps copy their rows
// there once (tcgen05.st), the MMAs take A from TMEM -- so that shared memory feeds the B operand only (a 128 x 128 x 16
// instruction with both operands in shared memory reads 8 KB per 64 tensor cycles: all of the shared-memory bandwidth) and
// the whole shared memory is ring.  Needs 2 BN + (C/64) * 32 * parts <= 512 TMEM columns.
// ES: epilogue warps per TMEM lane quadrant (1 or 2).  With 2, the two warps of a quadrant take the lower and the upper half
// of every accumulator block's columns and the upper one hands its running (best, runner-up, third) to the lower one at
// the end of the CTA: the compare chain per accumulator block halves.
template <int ROWT, int PASSES, bool A_RESIDENT, int CL, int BN = 128, bool A_TMEM = false, int ES = 1>
__global__ void __launch_bounds__(160 + 128 * ROWT * ES, 1) corr_tc_kernel(const TcParams prm) {
  static_assert(ES == 1 || ES == 2 || (ES == 4 && ROWT == 1), "one, two or (one row tile) four epilogue warps per lane quadrant");
  static_assert(!A_TMEM || (A_RESIDENT && ROWT == 1), "the TMEM row tile replaces a resident one, one row tile per CTA");
  static_assert(BN == 128 || (BN == 256 && ROWT == 1 && PASSES == 1), "256-column blocks: one row tile, one pass");
  constexpr int kBlockN = BN;                                  // shadows the namespace constant inside this kernel
  constexpr int kBT = BN / 128;                                // bank tiles per block
  static_assert(PASSES == 1 || PASSES == 3, "one hi*hi pass or the three-pass split");
  static_assert(CL == 1 || CL == 2, "single CTAs or CTA pairs");
  static_assert(ROWT == 1 || (A_RESIDENT && ROWT == 2), "two row tiles need the resident A operand");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // manual 1024-byte alignment (SWIZZLE_128B atoms repeat every 1024 bytes)
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  constexpr int AH = (PASSES == 3) ? 2 : 1;                    // operand parts per 64-channel block (hi[, lo])
  constexpr uint32_t kStageBytes = (A_RESIDENT ? 0u : (uint32_t)AH * kTileBytes) + (uint32_t)AH * kTileBytes * kBT;
  constexpr int kEpiWarps = 4 * ROWT * ES;
  constexpr uint32_t kTmemCols = A_TMEM ? 512u : (uint32_t)ROWT * 2u * kBlockN;
  constexpr uint32_t kATmemCol0 = 2u * kBlockN;                // the row tile sits behind the two accumulators
  pdl_trigger();
  pdl_wait();                                                  // (row_limit below is the previous kernel's output)
  const int KB = prm.KB, RB = prm.RB;
  const uint32_t a_bytes = (A_RESIDENT && !A_TMEM) ? (uint32_t)ROWT * KB * AH * kTileBytes : 0u;
  const uint32_t a_base = base;
  const uint32_t ring_base = base + a_bytes;
  const uint32_t bar_base = ring_base + (uint32_t)prm.stages * kStageBytes;
  // barriers: full[stages], empty[stages], tmem_full[2], tmem_empty[2], a_full[KB] (one per 64-channel block of the resident
  // row tile: the first MMAs start when ITS block has landed, not the whole tile); then the TMEM address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (prm.stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * prm.stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * prm.stages + 2 + s); };
  auto a_full_bar = [&](int kb) { return bar_base + 8u * (2 * prm.stages + 4 + kb); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * prm.stages + 4 + (A_RESIDENT ? KB : 1));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work decomposition: blockIdx.x -> (.., split[, rank in the pair]).  With a compacted operand (row_limit) the
  // populated tiles are the FIRST of every image: row-tile group slowest, so every CTA that has work is scheduled
  // before the ones that only find out that they have none
  const int RBG = RB / ROWT;                 // row-tile groups per image
  const int RBGc = RBG / CL;                 // ... per cluster
  const uint32_t crank = (CL == 2) ? cluster_ctarank() : 0u;
  int t = (CL == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int split = t % prm.psplit; t /= prm.psplit;
  // full operand: image slowest, so that the CTAs resident together share one image's bank tiles in L2
  const int b = prm.row_limit ? t % prm.B : t / RBGc;
  const int rbgc = prm.row_limit ? t / prm.B : t % RBGc;
  const int rbg = rbgc * CL + (int)crank;
  // (a pair leaves only together: the decision looks at the pair's first tile; a CTA whose own tile lies beyond the
  // limit multiplies stale rows, which nobody reads)
  if (prm.row_limit && rbgc * CL * ROWT * kTileRows >= prm.row_limit[b]) return;
  const int per = (prm.blocks_total + prm.psplit - 1) / prm.psplit;
  const int blk0 = split * per;
  const int blk1 = min(prm.blocks_total, blk0 + per);
  const int nblk = max(0, blk1 - blk0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < prm.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CL);             // one commit per CTA of the cluster
    }
    if (A_RESIDENT)                              // per block: one expect_tx arrival per copy | (TMEM row tile) the four loading warps
      for (int kb = 0; kb < KB; ++kb) mbar_init(a_full_bar(kb), A_TMEM ? 4 : (uint32_t)(ROWT * AH));
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiWarps);     // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (CL == 2) cluster_sync_all();             // the peer's barriers exist before anything is sent to them

  constexpr int kProducer2 = 2 + 4 * ROWT * ES;                  // the other producer warps (up to 3) sit behind the epilogue warps
  if (warp == 0 || warp >= kProducer2) {
    // ------------------------------------------------------------------ producers
    // Measured (scripts/micro/bulk_bw.cu): a thread gets ONE bulk copy out per ~650 cycles whatever its size -- about the
    // time the tensor core needs for a whole stage -- while copies of different warps proceed side by side.  So the
    // stages alternate between two issuing warps.
    const int pid = warp == 0 ? 0 : warp - kProducer2 + 1;
    const int nprod = prm.nprod;
    if (lane == 0 && nblk > 0 && pid < nprod) {
      if (A_RESIDENT && !A_TMEM) {
        int idx = 0;
        for (int kb = 0; kb < KB; ++kb)                      // block 0 first: the MMAs of the first stage wait for it only
          for (int rt = 0; rt < ROWT; ++rt)
            for (int hl = 0; hl < AH; ++hl, ++idx)
              if (idx % nprod == pid) {
                mbar_expect_tx(a_full_bar(kb), kTileBytes);
                bulk_g2s(a_base + (uint32_t)((rt * KB + kb) * AH + hl) * kTileBytes,
                         prm.r_tiles + tile_offset_bytes_n(b, kb, hl, rbg * ROWT + rt, KB, RB, prm.a_parts), kTileBytes, a_full_bar(kb));
              }
      }
      int it = 0;
      for (int blk = blk0; blk < blk1; ++blk) {
        const int cb = (prm.col_begin >> 7) + blk * kBT;          // first 128-row bank tile of the block
        for (int kb = 0; kb < KB; ++kb, ++it) {
          if (it % nprod != pid) continue;
          const int s = it % prm.stages;
          const uint32_t ph = (uint32_t)(it / prm.stages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), kStageBytes);
          uint32_t dst = ring_base + (uint32_t)s * kStageBytes;
          if (!A_RESIDENT) {
            for (int hl = 0; hl < AH; ++hl, dst += kTileBytes)
              bulk_g2s(dst, prm.r_tiles + tile_offset_bytes_n(b, kb, hl, rbg, KB, RB, prm.a_parts), kTileBytes, full_bar(s));
          }
          if (kBT == 2) {                        // two adjacent bank tiles (contiguous in the tile image); a pair fetches one each
            if (CL == 1) bulk_g2s(dst, prm.x_tiles + tile_offset_bytes(b, kb, 0, cb, KB, RB), 2 * kTileBytes, full_bar(s));
            else bulk_g2s_multicast(dst + crank * kTileBytes, prm.x_tiles + tile_offset_bytes(b, kb, 0, cb + (int)crank, KB, RB),
                                    kTileBytes, full_bar(s), (uint16_t)0x3);
          } else if (CL == 1) {
            for (int hl = 0; hl < AH; ++hl, dst += kTileBytes)
              bulk_g2s(dst, prm.x_tiles + tile_offset_bytes(b, kb, hl, cb, KB, RB), kTileBytes, full_bar(s));
          } else if (AH == 2) {                  // this CTA fetches the hi (rank 0) or the lo (rank 1) tile for both
            bulk_g2s_multicast(dst + crank * kTileBytes, prm.x_tiles + tile_offset_bytes(b, kb, (int)crank, cb, KB, RB), kTileBytes,
                               full_bar(s), (uint16_t)0x3);
          } else {                               // ... the upper or the lower 64 rows of the hi tile
            bulk_g2s_multicast(dst + crank * (kTileBytes / 2),
                               prm.x_tiles + tile_offset_bytes(b, kb, 0, cb, KB, RB) + crank * (kTileBytes / 2), kTileBytes / 2,
                               full_bar(s), (uint16_t)0x3);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && nblk > 0) {
      const uint32_t idesc = umma_idesc_f16(128, kBlockN);
      if (A_TMEM) {
        mbar_wait(a_full_bar(0), 0);
        tc_fence_after();
      }
      int it = 0;
      for (int j = 0; j < nblk; ++j) {
        const int as = j & 1;
        const uint32_t aph = (uint32_t)(j >> 1) & 1u;
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % prm.stages;
          const uint32_t ph = (uint32_t)(it / prm.stages) & 1u;
          mbar_wait(full_bar(s), ph);
          if (A_RESIDENT && !A_TMEM && j == 0) mbar_wait(a_full_bar(kb), 0);   // this block of the row tile has landed
          tc_fence_after();
          const uint32_t stage = ring_base + (uint32_t)s * kStageBytes;
          const uint32_t b_hi = stage + (A_RESIDENT ? 0u : (uint32_t)AH * kTileBytes);
          const uint32_t b_lo = b_hi + kTileBytes;
#pragma unroll
          for (int rt = 0; rt < ROWT; ++rt) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((rt * 2 + as) * kBlockN);
            const uint32_t a_hi = A_RESIDENT ? a_base + (uint32_t)((rt * KB + kb) * AH) * kTileBytes : stage;
            const uint32_t a_lo = a_hi + kTileBytes;
            if (A_TMEM) {
              const uint32_t t_hi = tmem_base + kATmemCol0 + (uint32_t)kb * 32u, t_lo = t_hi + (uint32_t)KB * 32u;
              const uint64_t dbh = umma_desc_k_sw128(b_hi), dbl = umma_desc_k_sw128(b_lo);
              if (PASSES == 3) {
                const uint32_t ta[3] = {t_hi, t_lo, t_hi};
                const uint64_t db[3] = {dbl, dbh, dbh};
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                  for (int k = 0; k < kTileK / 16; ++k)
                    umma_f16_ts(d_tmem, ta[pass] + (uint32_t)(8 * k), db[pass] + (uint64_t)(2 * k), idesc,
                                (kb > 0 || pass > 0 || k > 0) ? 1u : 0u);
                }
              } else {
#pragma unroll
                for (int k = 0; k < kTileK / 16; ++k)
                  umma_f16_ts(d_tmem, t_hi + (uint32_t)(8 * k), dbh + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
              }
            } else if (PASSES == 3) {
              const uint64_t da[3] = {umma_desc_k_sw128(a_hi), umma_desc_k_sw128(a_lo), umma_desc_k_sw128(a_hi)};
              const uint64_t db[3] = {umma_desc_k_sw128(b_lo), umma_desc_k_sw128(b_hi), umma_desc_k_sw128(b_hi)};
              // small cross terms first, hi*hi last
#pragma unroll
              for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                for (int k = 0; k < kTileK / 16; ++k) {
                  // +32 bytes per 16-element K step inside the 128-byte swizzle row (descriptor units of 16 B)
                  umma_f16(d_tmem, da[pass] + (uint64_t)(2 * k), db[pass] + (uint64_t)(2 * k), idesc,
                           (kb > 0 || pass > 0 || k > 0) ? 1u : 0u);
                }
              }
            } else {
              const uint64_t da = umma_desc_k_sw128(a_hi), db = umma_desc_k_sw128(b_hi);
#pragma unroll
              for (int k = 0; k < kTileK / 16; ++k)
                umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (CL == 2) umma_commit_multicast(empty_bar(s), (uint16_t)0x3);   // both CTAs must retire the slot
          else umma_commit(empty_bar(s));     // ring slot reusable once these MMAs retire
        }
        umma_commit(tfull_bar(as));           // accumulators complete -> epilogue
      }
    }
  } else if (warp < kProducer2) {
    // ------------------------------------------------------------------ epilogue (thread == row)
    const int rt = (warp - 2) / (4 * ES);
    const int half = ((warp - 2) >> 2) % ES;             // which part of every block's columns this warp compares
    const int quad = warp & 3;                           // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    const int q = (rbg * ROWT + rt) * kTileRows + row;   // (compact) row index
    if (A_TMEM && nblk > 0 && half == 0) {
      // this thread's row of the tile image (8 x 16-byte chunks per 64-channel block) -> its TMEM lane
      const int nparts = (PASSES == 3) ? 2 : 1;
      for (int hl = 0; hl < nparts; ++hl)
        for (int kb = 0; kb < KB; ++kb) {
          const uint8_t* src = prm.r_tiles + tile_offset_bytes_n(b, kb, hl, rbg, KB, RB, prm.a_parts);
          uint32_t v[32];
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint4 qv = __ldg(reinterpret_cast<const uint4*>(src + tile_chunk_offset(row, ch)));
            v[4 * ch + 0] = qv.x; v[4 * ch + 1] = qv.y; v[4 * ch + 2] = qv.z; v[4 * ch + 3] = qv.w;
          }
          tmem_st32(tmem_base + ((uint32_t)(quad * 32) << 16) + kATmemCol0 + (uint32_t)(hl * KB + kb) * 32u, v);
        }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full_bar(0));
    }
    float best = -INFINITY, second = -INFINITY, third = -INFINITY;
    int bidx = prm.col_begin + blk0 * kBlockN, sidx = bidx;
    const bool top3 = prm.part_idx2 != nullptr;
    float* dump = prm.s_dump ? prm.s_dump + ((size_t)b * prm.N + q) * prm.N : nullptr;
    for (int j = 0; j < nblk; ++j) {
      const int as = j & 1;
      const uint32_t aph = (uint32_t)(j >> 1) & 1u;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const int colb = prm.col_begin + (blk0 + j) * kBlockN;
      // the TMEM load of chunk ch+1 is in flight while chunk ch is compared; the accumulator is handed back to the MMA
      // warp as soon as its last chunk sits in registers
      uint32_t rbuf[2][32];
      constexpr int kChunks = kBlockN / 32 / ES;           // 32-column chunks of this warp
      const uint32_t tsrc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((rt * 2 + as) * kBlockN + half * kChunks * 32);
      tmem_ld32(tsrc, rbuf[0]);
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch) {
        uint32_t (&r)[32] = rbuf[ch & 1];
        tmem_ld_wait();
        if (ch + 1 < kChunks) {
          tmem_ld32(tsrc + (uint32_t)((ch + 1) * 32), rbuf[(ch + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
        const int c0 = colb + (half * kChunks + ch) * 32;
        if (c0 + 32 > prm.n_valid) {                       // (warp-uniform) padding columns: -inf
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c0 + e >= prm.n_valid) r[e] = 0xFF800000u;
        }
        if (top3) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float v = __uint_as_float(r[e]);
            const bool g1 = v > best;                     // strict: the lowest column wins ties
            const bool g2 = v > second;
            third = g2 ? second : fmaxf(third, v);
            sidx = g1 ? bidx : (g2 ? (c0 + e) : sidx);
            second = g1 ? best : (g2 ? v : second);
            bidx = g1 ? (c0 + e) : bidx;
            best = g1 ? v : best;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float v = __uint_as_float(r[e]);
            const bool gt = v > best;                     // strict: the lowest column wins ties
            second = gt ? best : fmaxf(second, v);
            bidx = gt ? (c0 + e) : bidx;
            best = gt ? v : best;
          }
        }
        if (dump) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dump + c0 + e) =
                make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]),
                            __uint_as_float(r[e + 3]));
        }
      }

    }
    if (ES >= 2) {
      // the other warps of the quadrant hand their states over (the ring is idle by now: every MMA that read it has completed)
      float* mrg0 = reinterpret_cast<float*>(smem + (ring_base - base)) + (size_t)(rt * kTileRows + row) * 5;
      constexpr size_t kMrgPart = (size_t)ROWT * kTileRows * 5;
      if (half > 0) {
        float* mrg = mrg0 + (size_t)(half - 1) * kMrgPart;
        mrg[0] = best; mrg[1] = __int_as_float(bidx); mrg[2] = second; mrg[3] = __int_as_float(sidx); mrg[4] = third;
      }
      asm volatile("bar.sync 2, %0;" ::"n"(128 * ROWT * ES) : "memory");
      if (half == 0 && nblk > 0) {
        // (value, column) pairs in descending value, ascending column order: equal values keep the lower column first
        auto insert = [&](float v, int i) {
          if (v > best || (v == best && i < bidx)) {
            third = second; second = best; sidx = bidx; best = v; bidx = i;
          } else if (v > second || (v == second && i < sidx)) {
            third = second; second = v; sidx = i;
          } else {
            third = fmaxf(third, v);
          }
        };
#pragma unroll
        for (int h = 1; h < ES; ++h) {
          const float* mrg = mrg0 + (size_t)(h - 1) * kMrgPart;
          insert(mrg[0], __float_as_int(mrg[1]));
          insert(mrg[2], top3 ? __float_as_int(mrg[3]) : 0x7FFFFFFF);
          third = fmaxf(third, mrg[4]);
        }
      }
    }
    if (ES >= 2 && half > 0) {
      // (the lower-half warp reports for the row)
    } else if (prm.fin.ind != nullptr) {
      // the finalize decision inline (same arithmetic as finalize_kernel)
      const TcFinalize& f = prm.fin;
      const size_t bq = (size_t)b * prm.N + q;
      const float rn = f.rnorm[bq], rs = f.rscale[bq];
      const float tol = f.rerr ? fmaf(2.0f, fmaf(rn, fmaf(1.001f, f.xerr_max[b], f.tol_rel), f.rerr[bq]), f.tol_abs)   // single pass
                               : fmaf(f.tol_rel, rn, f.tol_abs);                                                      // three-pass split
      const float gap = __fmul_rn(best - second, rs);
      const bool bad = (f.nonfinite && f.nonfinite[b] != 0);
      f.ind[bq] = bidx;                                      // provisional for untrusted rows
      bool amb = bad || !(gap >= tol);                       // !(gap >= tol) also catches NaN and the all -inf row
      if (amb && top3 && !bad && (__fmul_rn(best - third, rs) >= tol)) {
        amb = false;                                         // exactly two candidates: two exact dot products settle the row
        f.cand2[bq] = sidx;
        f.pair_list[(size_t)b * prm.N + atomicAdd(f.npair + b, 1)] = q;
      }
      if (amb) f.list[(size_t)b * prm.N + atomicAdd(f.nlist + b, 1)] = q;
      if (f.packed) f.packed[bq] = kPackedIdentity;
    } else {
      const size_t o = ((size_t)split * prm.B + b) * prm.N + q;
      prm.part_best[o] = best;
      prm.part_idx[o] = bidx;
      prm.part_second[o] = second;
      if (top3) {
        prm.part_idx2[o] = sidx;
        prm.part_third[o] = third;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();             // nobody leaves while the peer may still write into this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// The single pass, PERSISTENT: one CTA pair per SM pair walks its share of the (image, row-tile pair, column split) items.
// Inside its MMA loop the one-item kernel above already runs at the tensor pipe's sustained rate (one 128 x 256 x 16 MMA
// per ~181 cycles); what it loses is per CTA -- launch, barrier / TMEM set-up, the row tile that must land before the first
// MMA, the epilogue of the last block, 14 times per SM at 64 x 64 x 256 (batch 64).  Here
//   * barriers, TMEM and the cluster handshake are set up once;
//   * the row tile of the NEXT item replaces the current one block by block: its 64-channel block kb is fetched as soon as
//     the last MMAs that read block kb of the current tile have retired (the producer learns that from the ring slot it is
//     about to refill) and lands while the current item's remaining stages run -- one A buffer, the ring keeps its depth;
//   * the epilogue of an item's last block overlaps the first MMAs of the next item (the accumulators keep alternating).
// ROWT = 1, BN = 256, CTA pairs (multicast bank tiles), two epilogue warps per TMEM lane quadrant; writes the per-split
// partial results (ipsr_finalize_argmax merges them).  352 threads: warp 0 / warp 10 producers, warp 1 MMA, warps 2..9 epilogue.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(352, 1) corr_tc_p1_persistent_kernel(const TcParams prm, int nitems) {
  constexpr int kBN = 256, ES = 2;
  constexpr uint32_t kStageBytes = 2u * kTileBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  pdl_trigger();
  pdl_wait();
  const int KB = prm.KB, RB = prm.RB, stages = prm.stages;
  const uint32_t a_base = base;
  const uint32_t ring_base = base + (uint32_t)KB * kTileBytes;
  const uint32_t bar_base = ring_base + (uint32_t)stages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * stages + 2 + s); };
  auto a_full_bar = [&](int kb) { return bar_base + 8u * (2 * stages + 4 + kb); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 4 + KB);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));
  float* mrg_all = reinterpret_cast<float*>(smem + (tmem_slot + 16u - base));        // [128][3] hand-over of the upper-half warps

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int cl = (int)(blockIdx.x >> 1), ncl = (int)(gridDim.x >> 1);
  const int RBGc = RB / 2;
  const int per = (prm.blocks_total + prm.psplit - 1) / prm.psplit;
  struct Item { int b, rbg, split, blk0, nblk; };
  auto decode = [&](int item) {
    Item w;
    int t = item;
    w.split = t % prm.psplit; t /= prm.psplit;
    w.b = t / RBGc;
    w.rbg = (t % RBGc) * 2 + (int)crank;
    w.blk0 = w.split * per;
    w.nblk = max(0, min(prm.blocks_total, w.blk0 + per) - w.blk0);
    return w;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 2);              // one commit per CTA of the pair
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4 * ES);
    }
    for (int kb = 0; kb < KB; ++kb) mbar_init(a_full_bar(kb), 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  cluster_sync_all();

  if (warp == 0 || warp == 10) {
    // ------------------------------------------------------------------ producers (alternate stages)
    const int pid = warp == 0 ? 0 : 1;
    if (lane == 0 && cl < nitems) {
      auto load_a = [&](const Item& w, int kb) {
        mbar_expect_tx(a_full_bar(kb), kTileBytes);
        bulk_g2s(a_base + (uint32_t)kb * kTileBytes, prm.r_tiles + tile_offset_bytes_n(w.b, kb, 0, w.rbg, KB, RB, prm.a_parts), kTileBytes,
                 a_full_bar(kb));
      };
      {
        const Item w0 = decode(cl);
        for (int kb = pid; kb < KB; kb += 2) load_a(w0, kb);
      }
      long long it = 0;
      // reload_at[kb]: the ring index whose slot was last used by (previous item, last block, kb): once that slot is free
      // again the MMAs that read block kb of the previous row tile have retired
      long long reload_at[16];
      Item reload_item[16];
      for (int kb = 0; kb < KB; ++kb) reload_at[kb] = -1;
      for (int item = cl; item < nitems; item += ncl) {
        const Item w = decode(item);
        const bool has_next = item + ncl < nitems;
        const Item nxt = has_next ? decode(item + ncl) : w;
        for (int j = 0; j < w.nblk; ++j) {
          const int cb = (prm.col_begin >> 7) + (w.blk0 + j) * 2;
          for (int kb = 0; kb < KB; ++kb, ++it) {
            const bool mine = (it % 2) == pid;           // (both producers keep the whole schedule; each acts on its own indices)
            if (mine) {
              const int s = (int)(it % stages);
              const uint32_t ph = (uint32_t)(it / stages) & 1u;
              mbar_wait(empty_bar(s), ph ^ 1u);
              for (int k2 = 0; k2 < KB; ++k2)
                if (reload_at[k2] == it) load_a(reload_item[k2], k2);
              mbar_expect_tx(full_bar(s), kStageBytes);
              bulk_g2s_multicast(ring_base + (uint32_t)s * kStageBytes + crank * kTileBytes,
                                 prm.x_tiles + tile_offset_bytes(w.b, kb, 0, cb + (int)crank, KB, RB), kTileBytes, full_bar(s), (uint16_t)0x3);
            }
            if (j == w.nblk - 1 && has_next) {     // this index holds (item, last block, kb): its slot comes free at it + stages
              reload_at[kb] = it + stages;
              reload_item[kb] = nxt;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && cl < nitems) {
      const uint32_t idesc = umma_idesc_f16(128, kBN);
      long long it = 0;
      int jt = 0, n = 0;
      for (int item = cl; item < nitems; item += ncl, ++n) {
        const Item w = decode(item);
        for (int j = 0; j < w.nblk; ++j, ++jt) {
          const int as = jt & 1;
          mbar_wait(tempty_bar(as), ((uint32_t)(jt >> 1) & 1u) ^ 1u);
          tc_fence_after();
          for (int kb = 0; kb < KB; ++kb, ++it) {
            const int s = (int)(it % stages);
            mbar_wait(full_bar(s), (uint32_t)(it / stages) & 1u);
            if (j == 0) mbar_wait(a_full_bar(kb), (uint32_t)n & 1u);     // block kb of THIS item's row tile has landed
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * kBN);
            const uint64_t da = umma_desc_k_sw128(a_base + (uint32_t)kb * kTileBytes);
            const uint64_t db = umma_desc_k_sw128(ring_base + (uint32_t)s * kStageBytes);
#pragma unroll
            for (int k = 0; k < kTileK / 16; ++k)
              umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_multicast(empty_bar(s), (uint16_t)0x3);
          }
          umma_commit(tfull_bar(as));
        }
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ epilogue (thread == row, two warps per lane quadrant)
    const int half = ((warp - 2) >> 2) % ES;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    float* mrg = mrg_all + (size_t)row * 3;
    int jt = 0;
    for (int item = cl; item < nitems; item += ncl) {
      const Item w = decode(item);
      const int q = w.rbg * kTileRows + row;
      float best = -INFINITY, second = -INFINITY;
      int bidx = prm.col_begin + w.blk0 * kBN;
      for (int j = 0; j < w.nblk; ++j, ++jt) {
        const int as = jt & 1;
        mbar_wait(tfull_bar(as), (uint32_t)(jt >> 1) & 1u);
        tc_fence_after();
        const int colb = prm.col_begin + (w.blk0 + j) * kBN;
        uint32_t rbuf[2][32];
        constexpr int kChunks = kBN / 32 / ES;
        const uint32_t tsrc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * kBN + half * kChunks * 32);
        tmem_ld32(tsrc, rbuf[0]);
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch) {
          uint32_t (&r)[32] = rbuf[ch & 1];
          tmem_ld_wait();
          if (ch + 1 < kChunks) {
            tmem_ld32(tsrc + (uint32_t)((ch + 1) * 32), rbuf[(ch + 1) & 1]);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(as));
          }
          const int c0 = colb + (half * kChunks + ch) * 32;
          if (c0 + 32 > prm.n_valid) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c0 + e >= prm.n_valid) r[e] = 0xFF800000u;
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float v = __uint_as_float(r[e]);
            const bool gt = v > best;                     // strict: the lowest column wins ties
            second = gt ? best : fmaxf(second, v);
            bidx = gt ? (c0 + e) : bidx;
            best = gt ? v : best;
          }
        }
      }
      // hand-over of the upper-half warp, then the row's partial result of this item
      if (half == 1) {
        mrg[0] = best; mrg[1] = __int_as_float(bidx); mrg[2] = second;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (half == 0) {
        if (w.nblk > 0) {
          const float b2 = mrg[0], s2 = mrg[2];
          const int i2 = __float_as_int(mrg[1]);
          if (b2 > best || (b2 == best && i2 < bidx)) {
            second = fmaxf(best, s2); best = b2; bidx = i2;
          } else {
            second = fmaxf(second, b2);
          }
        }
        const size_t o = ((size_t)w.split * prm.B + w.b) * prm.N + q;
        prm.part_best[o] = best;
        prm.part_idx[o] = bidx;
        prm.part_second[o] = second;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");       // the hand-over buffer is free for the next item
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Merge the per-split triples (scores in scaled units; true score = scaled * rscale[q] > 0 scaling) and decide
// which rows can be trusted.
//   rows: list_in == NULL: row index = q; else row index r < nlist_in[b] is position q = list_in[b][r]
//   after the single pass (rerr != NULL):  tol[q] = 2 (rerr[q] + rnorm[q] (1.001 xerr_max[b] + tol_rel)) + tol_abs
//   after the three-pass split (rerr == NULL):  tol[q] = tol_rel rnorm[q] + tol_abs
// Trusted rows get ind[b,q]; the others are appended to list_out[b] (order arbitrary).  Pass 1 with c_tiles != NULL
// also COMPACTS: the hi and lo tile rows of every appended position are copied to row `pos` of the compact tile
// images the three-pass split reads (one warp per row, 16-byte chunks, coalesced 128-byte tile rows).
__global__ void finalize_kernel(const float* __restrict__ part_best, const int* __restrict__ part_idx,
                                const float* __restrict__ part_second, int psplit, const float* __restrict__ rnorm,
                                const float* __restrict__ rscale, const float* __restrict__ rerr,
                                const float* __restrict__ xerr_max, const int* __restrict__ nonfinite,
                                const int* __restrict__ list_in, const int* __restrict__ nlist_in, int B, int N,
                                float tol_rel, float tol_abs, int* __restrict__ ind, int* __restrict__ list_out,
                                int* __restrict__ nlist_out, long long* __restrict__ packed,
                                const uint8_t* __restrict__ r_tiles, uint8_t* __restrict__ c_tiles, int KB,
                                const int* __restrict__ part_idx2, const float* __restrict__ part_third,
                                int* __restrict__ cand2, int* __restrict__ pair_list, int* __restrict__ npair, int n_valid) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = r < n_valid;                                // rows >= n_valid are padding of the tile image
  int q = r;
  if (list_in) {
    live = live && r < min(nlist_in[b], N);
    if (live) q = list_in[(size_t)b * N + r];
  }
  bool amb = false;
  int pos = 0;
  if (live) {
    float best = -INFINITY, second = -INFINITY, third = -INFINITY;
    int idx = 0, idx2 = 0;
    const bool top3 = part_idx2 != nullptr;
    auto insert = [&](float v, int i) {
      if (v > best) {
        third = second;
        second = best;
        idx2 = idx;
        best = v;
        idx = i;
      } else if (v > second) {
        third = second;
        second = v;
        idx2 = i;
      } else {
        third = fmaxf(third, v);
      }
    };
    for (int s = 0; s < psplit; ++s) {                     // splits are in ascending column order
      const size_t o = ((size_t)s * B + b) * N + r;
      const float vb = part_best[o], vs = part_second[o];
      const float vt = top3 ? part_third[o] : -INFINITY;
      const int ib = part_idx[o], is = top3 ? part_idx2[o] : 0;
      // insert (vb, ib), (vs, is), vt into the running top three; strict comparisons keep the lowest column on ties
      // (vt <= vs <= the running second once vs is in, so it only competes for third place)
      insert(vb, ib);
      insert(vs, is);
      third = fmaxf(third, vt);
    }
    const size_t bq = (size_t)b * N + q;
    const float rn = rnorm[bq];
    float tol;
    if (rerr == nullptr) tol = fmaf(tol_rel, rn, tol_abs);                 // after the three-pass split
    else tol = fmaf(2.0f, fmaf(rn, fmaf(1.001f, xerr_max[b], tol_rel), rerr[bq]), tol_abs);   // after the single pass
    const float rs = rscale[bq];
    const float gap = __fmul_rn(best - second, rs);
    const bool bad = (nonfinite && nonfinite[b] != 0);
    ind[bq] = idx;                                         // provisional for untrusted rows
    // !(gap >= tol) also catches NaN and the all -inf row
    amb = bad || !(gap >= tol);
    if (amb && top3 && !bad && (__fmul_rn(best - third, rs) >= tol)) {
      // exactly two candidates inside the error band: two exact dot products settle the row (ipsr_resolve_rows)
      amb = false;
      cand2[bq] = idx2;
      pair_list[(size_t)b * N + atomicAdd(npair + b, 1)] = q;
    }
    if (amb) {
      pos = atomicAdd(nlist_out + b, 1);
      list_out[(size_t)b * N + pos] = q;
    }
    if (packed) packed[bq] = kPackedIdentity;
  }
  if (c_tiles == nullptr) return;
  // compaction (pass 1 only; N % 32 == 0 on the tensor path, so whole warps arrive here)
  const int lane = threadIdx.x & 31;
  const int RB = N / kTileRows;
  unsigned todo = __ballot_sync(0xffffffffu, amb);
  while (todo) {
    const int src_lane = __ffs(todo) - 1;
    todo &= todo - 1;
    const int qs = __shfl_sync(0xffffffffu, q, src_lane);
    const int ps = __shfl_sync(0xffffffffu, pos, src_lane);
    const int srb = qs / kTileRows, sr = qs % kTileRows, drb = ps / kTileRows, dr = ps % kTileRows;
    for (int i = lane; i < KB * 16; i += 32) {             // (kb, part, chunk)
      const int kb = i >> 4, part = (i >> 3) & 1, chunk = i & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(r_tiles + tile_offset_bytes(b, kb, part, srb, KB, RB) + tile_chunk_offset(sr, chunk));
      *reinterpret_cast<uint4*>(c_tiles + tile_offset_bytes(b, kb, part, drb, KB, RB) + tile_chunk_offset(dr, chunk)) = v;
    }
  }
}

static int tc_stage_count(size_t a_bytes, size_t stage, size_t* smem_out, int cap = 8) {
  const size_t fixed = a_bytes + 1024 /*alignment slack*/ + 512 /*barriers*/;
  const size_t budget = 227 * 1024;
  if (fixed + 2 * stage > budget) {
    *smem_out = 0;
    return 0;
  }
  int stages = (int)((budget - fixed) / stage);
  if (stages > cap) stages = cap;
  *smem_out = fixed + (size_t)stages * stage;
  return stages;
}

template <int ROWT, int PASSES, bool A_RES, int CL, int BN, bool A_TMEM, int ES>
static int launch_tc_es(TcParams prm, int C, long long ctas, cudaStream_t st);

template <int ROWT, int PASSES, bool A_RES, int CL, int BN = 128, bool A_TMEM = false>
static int launch_tc(TcParams prm, int C, long long ctas, cudaStream_t st) {
  static const int epi = [] {                              // IPSR_TC_EPI=<1|2|4> (A/B runs): epilogue warps per lane quadrant
    const char* e = getenv("IPSR_TC_EPI");
    const int v = e ? atoi(e) : 2;
    return (v == 1 || v == 4) ? v : 2;
  }();
  if constexpr (ROWT == 1) {
    if (epi == 4) return launch_tc_es<ROWT, PASSES, A_RES, CL, BN, A_TMEM, 4>(prm, C, ctas, st);
  }
  return epi >= 2 ? launch_tc_es<ROWT, PASSES, A_RES, CL, BN, A_TMEM, 2>(prm, C, ctas, st)
                  : launch_tc_es<ROWT, PASSES, A_RES, CL, BN, A_TMEM, 1>(prm, C, ctas, st);
}

template <int ROWT, int PASSES, bool A_RES, int CL, int BN, bool A_TMEM, int ES>
static int launch_tc_es(TcParams prm, int C, long long ctas, cudaStream_t st) {
  constexpr int AH = (PASSES == 3) ? 2 : 1;
  const size_t a_bytes = (A_RES && !A_TMEM) ? (size_t)ROWT * (C / kTileK) * AH * kTileBytes : 0;
  const size_t stage = (A_RES ? 0 : (size_t)AH * kTileBytes) + (size_t)AH * kTileBytes * (BN / 128);
  size_t smem = 0;
  prm.stages = tc_stage_count(a_bytes, stage, &smem, A_TMEM ? 13 : 8);
  static const int producers = [] {                        // IPSR_TC_PRODUCERS=<1..4> (A/B runs): issuing warps
    const char* e = getenv("IPSR_TC_PRODUCERS");
    const int v = e ? atoi(e) : 2;
    return v < 1 ? 1 : (v > 4 ? 4 : v);
  }();
  prm.nprod = producers;
  IPSR_REQUIRE(prm.stages >= 2, IPSR_ERR_UNSUPPORTED, "ipsr_correlate_argmax_tc: C=%d leaves %d pipeline stages", C, prm.stages);
  auto kern = corr_tc_kernel<ROWT, PASSES, A_RES, CL, BN, A_TMEM, ES>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "corr_tc smem attribute (%zu B): %s", smem, cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(64 + 128 * ROWT * ES + 32 * (prm.nprod - 1));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, kern, prm);
  IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "ipsr_correlate_argmax_tc: launch failed: %s", cudaGetErrorString(e));
  return check_launch("ipsr_correlate_argmax_tc");
}

template <int ROWT, int PASSES, bool A_RES>
static int launch_tc_cl(const TcParams& prm, int C, long long ctas, bool pairs, cudaStream_t st) {
  return pairs ? launch_tc<ROWT, PASSES, A_RES, 2>(prm, C, ctas, st) : launch_tc<ROWT, PASSES, A_RES, 1>(prm, C, ctas, st);
}

}  // namespace ipsr

// Does the single pass over [col_begin, col_end) with this column split run the 128 x 256 x 16 variant?  (Also asked by
// the fused forward, which sizes the split for one row tile per CTA in that case.)
bool ipsr::tc_pass1_wide(int B, int C, int N, int col_begin, int col_end, int psplit) {
  static const int wide_stages = [] {            // minimum ring depth; 0 = never
    const char* e = getenv("IPSR_TC_BN256");
    return e ? atoi(e) : 3;
  }();
  if (wide_stages <= 0 || psplit < 1 || (col_end - col_begin) % 256 != 0) return false;
  const int blocks256 = (col_end - col_begin) / 256;
  const size_t a_one = (size_t)(C / kTileK) * kTileBytes;
  return blocks256 % psplit == 0 && a_one + (size_t)wide_stages * (2 * kTileBytes) + 1280 <= 227 * 1024 &&
         (long long)B * (N / kTileRows) * psplit >= 100;
}

extern "C" int ipsr_tensor_path_supported(int C, int N) {
  return (C > 0 && N > 0 && C % ipsr::kTileK == 0 && N % ipsr::kTileRows == 0 && N <= 65536) ? 1 : 0;
}

extern "C" int ipsr_correlate_argmax_tc(const void* r_tiles, const void* x_tiles, int B, int C, int N,
                                        int col_begin, int col_end, int psplit, int passes, int r_parts,
                                        const int32_t* row_limit, float* part_best, int32_t* part_idx, float* part_second,
                                        int32_t* part_idx2, float* part_third, float* s_dump, void* stream) {
  return ipsr_correlate_argmax_tc_valid(r_tiles, x_tiles, B, C, N, col_begin, col_end, psplit, passes, r_parts, row_limit, part_best,
                                        part_idx, part_second, part_idx2, part_third, s_dump, N, stream);
}

extern "C" int ipsr_correlate_argmax_tc_valid(const void* r_tiles, const void* x_tiles, int B, int C, int N,
                                              int col_begin, int col_end, int psplit, int passes, int r_parts,
                                              const int32_t* row_limit, float* part_best, int32_t* part_idx, float* part_second,
                                              int32_t* part_idx2, float* part_third, float* s_dump, int n_valid, void* stream) {
  return ipsr::correlate_argmax_tc_ex(r_tiles, x_tiles, B, C, N, col_begin, col_end, psplit, passes, r_parts, row_limit, part_best,
                                      part_idx, part_second, part_idx2, part_third, s_dump, n_valid, nullptr, stream);
}

int ipsr::correlate_argmax_tc_ex(const void* r_tiles, const void* x_tiles, int B, int C, int N, int col_begin, int col_end, int psplit,
                                 int passes, int r_parts, const int32_t* row_limit, float* part_best, int32_t* part_idx,
                                 float* part_second, int32_t* part_idx2, float* part_third, float* s_dump, int n_valid,
                                 const TcFinalize* fin, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(!fin || (psplit == 1 && n_valid == N && !row_limit && fin->ind && fin->rnorm && fin->rscale && fin->list &&
                        fin->nlist && (!part_third || (fin->cand2 && fin->pair_list && fin->npair)) &&
                        (passes == 3 || (fin->rerr && fin->xerr_max))),
               IPSR_ERR_INVALID_ARG, "ipsr_correlate_argmax_tc: the fused finalize needs psplit = 1 and the whole operand");
  IPSR_REQUIRE(n_valid > 0 && n_valid <= N, IPSR_ERR_INVALID_ARG, "ipsr_correlate_argmax_tc: n_valid=%d outside (0, %d]", n_valid, N);
  IPSR_REQUIRE(r_tiles && x_tiles && part_best && part_idx && part_second, IPSR_ERR_INVALID_ARG,
               "ipsr_correlate_argmax_tc: null pointer");
  IPSR_REQUIRE(ipsr_tensor_path_supported(C, N), IPSR_ERR_UNSUPPORTED,
               "ipsr_correlate_argmax_tc: shape C=%d N=%d not supported (need C %% 64 == 0, N %% 128 == 0)", C, N);
  IPSR_REQUIRE(B > 0 && col_begin >= 0 && col_end <= N && col_begin < col_end && col_begin % 128 == 0 &&
                   col_end % 128 == 0 && psplit >= 1,
               IPSR_ERR_INVALID_ARG, "ipsr_correlate_argmax_tc: bad column range [%d,%d) psplit=%d", col_begin, col_end, psplit);
  IPSR_REQUIRE(passes == 1 || passes == 3, IPSR_ERR_INVALID_ARG, "ipsr_correlate_argmax_tc: passes must be 1 or 3");
  IPSR_REQUIRE(r_parts == 2 || (r_parts == 1 && passes == 1), IPSR_ERR_INVALID_ARG,
               "ipsr_correlate_argmax_tc: r_tiles must hold 2 parts (or 1 for a single pass)");
  TcParams prm;
  prm.r_tiles = reinterpret_cast<const uint8_t*>(r_tiles);
  prm.x_tiles = reinterpret_cast<const uint8_t*>(x_tiles);
  prm.row_limit = row_limit;
  prm.a_parts = r_parts;
  prm.B = B; prm.KB = C / kTileK; prm.RB = N / kTileRows; prm.N = N;
  prm.col_begin = col_begin;
  prm.blocks_total = (col_end - col_begin) / kBlockN;
  // psplit > blocks_total is allowed: the surplus splits own no column block and report -inf
  prm.psplit = psplit;
  prm.stages = 0;
  prm.part_best = part_best; prm.part_idx = part_idx; prm.part_second = part_second; prm.s_dump = s_dump;
  prm.part_idx2 = part_third ? part_idx2 : nullptr;       // runner-up column + third-best score (pair resolve)
  prm.part_third = part_third;
  prm.n_valid = n_valid;
  if (fin) prm.fin = *fin;
  else memset(&prm.fin, 0, sizeof(prm.fin));
  cudaStream_t st = as_stream(stream);
  const int AH = passes == 3 ? 2 : 1;
  const size_t a_one = (size_t)(C / kTileK) * AH * kTileBytes;           // resident bytes per 128-row tile
  const bool a_res = a_one + 3 * (size_t)AH * kTileBytes + 2048 <= 227 * 1024;    // resident rows + >= 3 ring stages
  if (passes == 3) {
    const long long ctas = (long long)B * prm.RB * psplit;
    IPSR_REQUIRE(ctas <= 0x7FFFFFFFll, IPSR_ERR_UNSUPPORTED, "ipsr_correlate_argmax_tc: grid too large");
    // IPSR_TC_ATMEM=1 / IPSR_TC_ATMEM1=1 (A/B runs): the row tile in tensor memory.  Measured SLOWER on the B200: 37.0 us against
    // 35.0 us for the three-pass launch at 32x32x256 (batch 16), 832 us against 609 us for the single pass at 64x64x256
    // (batch 64; 128-column instructions instead of 256-column ones) -- the A operand's shared-memory reads are not what
    // holds the tensor pipe back.
    static const int env_atmem = [] { const char* e = getenv("IPSR_TC_ATMEM"); return e ? atoi(e) : 0; }();
    if (env_atmem && s_dump == nullptr && 2 * 128 + (C / kTileK) * 32 * 2 <= 512)
      return (prm.RB % 2 == 0) ? launch_tc<1, 3, true, 2, 128, true>(prm, C, ctas, st) : launch_tc<1, 3, true, 1, 128, true>(prm, C, ctas, st);
    static const int env_pairs = [] { const char* e = getenv("IPSR_TC_PAIRS"); return e ? atoi(e) : 1; }();     // A/B runs
    static const int env_ares = [] { const char* e = getenv("IPSR_TC_ARES"); return e ? atoi(e) : 1; }();
    const bool pairs = (prm.RB % 2 == 0) && s_dump == nullptr && env_pairs != 0;
    return (a_res && env_ares) ? launch_tc_cl<1, 3, true>(prm, C, ctas, pairs, st) : launch_tc_cl<1, 3, false>(prm, C, ctas, pairs, st);
  }
  // two row tiles per CTA when the grid still fills most of the machine and both fit next to >= 4 ring stages
  const bool two = a_res && (prm.RB % 2 == 0) && (2 * a_one + 4 * (size_t)kTileBytes + 2048 <= 227 * 1024) &&
                   ((long long)B * (prm.RB / 2) * psplit >= 100) && s_dump == nullptr;
  const long long ctas = (long long)B * (two ? prm.RB / 2 : prm.RB) * psplit;
  IPSR_REQUIRE(ctas <= 0x7FFFFFFFll, IPSR_ERR_UNSUPPORTED, "ipsr_correlate_argmax_tc: grid too large");
  const bool pairs = ((two ? prm.RB / 2 : prm.RB) % 2 == 0) && s_dump == nullptr;
  // One row tile per CTA and 128 x 256 x 16 instructions (two adjacent bank tiles per stage) instead of two row tiles
  // and 128 x 128 x 16 instructions: measured 487 us against 550 us at B = 64, 64 x 64 x 256 (0.81 against 0.72 of the
  // cuBLAS rate) -- the wider instruction amortises its issue cost, and the CTA pairs still halve the L2 -> SM traffic.
  // At C = 512 (128 KiB resident row tile, only 3 stages of 32 KiB): 440 us against 739 us at B = 32, 64 x 64 x 512 (0.90
  // against 0.54).  Taken when the resident row tile leaves >= 3 stages of 32 KiB (C <= 512), the splits stay whole
  // blocks and the grid fills the machine.  IPSR_TC_BN256=<n> in the environment sets the minimum ring depth (0 turns it off; A/B runs).
  static const int env_atmem1 = [] { const char* e = getenv("IPSR_TC_ATMEM1"); return e ? atoi(e) : 0; }();       // A/B runs
  if (env_atmem1 && s_dump == nullptr && 2 * 128 + (C / kTileK) * 32 <= 512) {
    // single pass with the row tile in tensor memory: 128 x 128 x 16 instructions fed from shared memory with B only
    const long long ctas1 = (long long)B * prm.RB * psplit;
    IPSR_REQUIRE(ctas1 <= 0x7FFFFFFFll, IPSR_ERR_UNSUPPORTED, "ipsr_correlate_argmax_tc: grid too large");
    return (prm.RB % 2 == 0) ? launch_tc<1, 1, true, 2, 128, true>(prm, C, ctas1, st) : launch_tc<1, 1, true, 1, 128, true>(prm, C, ctas1, st);
  }
  if (s_dump == nullptr && tc_pass1_wide(B, C, N, col_begin, col_end, psplit)) {
    prm.blocks_total = (col_end - col_begin) / 256;
    {
      // persistent variant (IPSR_TC_PERSIST=0 turns it off): CTA pairs walk the items; needs whole pairs, several items
      // per pair and items long enough for the rolling reload of the row tile
      // (read per call: 0 = never, 1 = when every pair gets >= 4 items, 2 = whenever legal -- tests)
      const char* pe = getenv("IPSR_TC_PERSIST");
      const int env_persist = pe ? atoi(pe) : 1;
      const int KB = C / kTileK;
      const long long items = (long long)B * (prm.RB / 2) * psplit;
      const int per = prm.blocks_total / psplit;
      int stages = 0;
      size_t smem = 0;
      for (int sg = 8; sg >= 3; --sg) {
        const size_t need = 1024 + (size_t)KB * kTileBytes + (size_t)sg * 2 * kTileBytes + (size_t)(2 * sg + 4 + KB) * 8 + 16 + 1536;
        if (need <= 227 * 1024) { stages = sg; smem = need; break; }
      }
      if (env_persist && row_limit == nullptr && prm.RB % 2 == 0 && KB <= 16 && stages >= 3 && (items >= 4 * 74 || env_persist == 2) &&
          (long long)per * KB >= stages + KB && items <= 0x7FFFFFFF) {
        prm.stages = stages;
        prm.nprod = 2;
        cudaError_t e = cudaFuncSetAttribute(corr_tc_p1_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "corr_tc persistent smem attribute (%zu B): %s", smem, cudaGetErrorString(e));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148);
        cfg.blockDim = dim3(352);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        // as many CTA pairs as can be resident at once (GPCs with an odd SM count leave SMs without a partner)
        int nclusters = 0;
        e = cudaOccupancyMaxActiveClusters(&nclusters, corr_tc_p1_persistent_kernel, &cfg);
        IPSR_REQUIRE(e == cudaSuccess && nclusters > 0, IPSR_ERR_CUDA, "corr_tc persistent occupancy query: %s", cudaGetErrorString(e));
        if (getenv("IPSR_TC_PERSIST_VERBOSE")) fprintf(stderr, "ipsr: persistent single pass: %d CTA pairs resident, %lld items\n", nclusters, items);
        if ((long long)nclusters > items) nclusters = (int)items;
        cfg.gridDim = dim3(2 * nclusters);
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        e = cudaLaunchKernelEx(&cfg, corr_tc_p1_persistent_kernel, prm, (int)items);
        IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "ipsr_correlate_argmax_tc: launch failed: %s", cudaGetErrorString(e));
        return check_launch("ipsr_correlate_argmax_tc");
      }
    }
    const long long ctas1 = (long long)B * prm.RB * psplit;
    IPSR_REQUIRE(ctas1 <= 0x7FFFFFFFll, IPSR_ERR_UNSUPPORTED, "ipsr_correlate_argmax_tc: grid too large");
    return (prm.RB % 2 == 0) ? launch_tc<1, 1, true, 2, 256>(prm, C, ctas1, st) : launch_tc<1, 1, true, 1, 256>(prm, C, ctas1, st);
  }
  if (two) return launch_tc_cl<2, 1, true>(prm, C, ctas, pairs, st);
  return a_res ? launch_tc_cl<1, 1, true>(prm, C, ctas, pairs, st) : launch_tc_cl<1, 1, false>(prm, C, ctas, pairs, st);
}

extern "C" int ipsr_finalize_argmax(const float* part_best, const int32_t* part_idx, const float* part_second,
                                    int psplit, const float* rnorm, const float* rscale, const float* rerr,
                                    const float* xerr_max, const int32_t* nonfinite,
                                    const int32_t* list_in, const int32_t* nlist_in,
                                    int B, int N, float tol_rel, float tol_abs,
                                    int32_t* ind, int32_t* list_out, int32_t* nlist_out, int64_t* packed,
                                    const void* r_tiles, void* c_tiles, int C,
                                    const int32_t* part_idx2, const float* part_third,
                                    int32_t* cand2, int32_t* pair_list, int32_t* npair, void* stream) {
  return ipsr_finalize_argmax_valid(part_best, part_idx, part_second, psplit, rnorm, rscale, rerr, xerr_max, nonfinite, list_in,
                                    nlist_in, B, N, tol_rel, tol_abs, ind, list_out, nlist_out, packed, r_tiles, c_tiles, C, part_idx2,
                                    part_third, cand2, pair_list, npair, N, stream);
}

extern "C" int ipsr_finalize_argmax_valid(const float* part_best, const int32_t* part_idx, const float* part_second,
                                          int psplit, const float* rnorm, const float* rscale, const float* rerr,
                                          const float* xerr_max, const int32_t* nonfinite,
                                          const int32_t* list_in, const int32_t* nlist_in,
                                          int B, int N, float tol_rel, float tol_abs,
                                          int32_t* ind, int32_t* list_out, int32_t* nlist_out, int64_t* packed,
                                          const void* r_tiles, void* c_tiles, int C,
                                          const int32_t* part_idx2, const float* part_third,
                                          int32_t* cand2, int32_t* pair_list, int32_t* npair, int n_valid, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(n_valid > 0 && n_valid <= N, IPSR_ERR_INVALID_ARG, "ipsr_finalize_argmax: n_valid=%d outside (0, %d]", n_valid, N);
  IPSR_REQUIRE(part_best && part_idx && part_second && rnorm && rscale && ind && list_out && nlist_out,
               IPSR_ERR_INVALID_ARG, "ipsr_finalize_argmax: null pointer");
  IPSR_REQUIRE(!list_in || nlist_in, IPSR_ERR_INVALID_ARG, "ipsr_finalize_argmax: list_in needs nlist_in");
  IPSR_REQUIRE(!rerr || xerr_max, IPSR_ERR_INVALID_ARG, "ipsr_finalize_argmax: the single-pass bound needs rerr and xerr_max");
  IPSR_REQUIRE(B > 0 && N > 0 && psplit >= 1 && B <= 65535, IPSR_ERR_INVALID_ARG, "ipsr_finalize_argmax: bad dims");
  if (c_tiles)
    IPSR_REQUIRE(r_tiles && !list_in && ipsr_tensor_path_supported(C, N), IPSR_ERR_INVALID_ARG,
                 "ipsr_finalize_argmax: compaction needs pass 1, r_tiles and a tensor-path shape");
  if (part_idx2)
    IPSR_REQUIRE(part_third && cand2 && pair_list && npair, IPSR_ERR_INVALID_ARG,
                 "ipsr_finalize_argmax: the pair list needs part_third, cand2, pair_list and npair");
  {
    cudaError_t le__ = launch_pdl(finalize_kernel, dim3((N + 255) / 256, B), dim3(256), 0, as_stream(stream), part_best, part_idx, part_second,
                                  psplit, rnorm, rscale, rerr, xerr_max, nonfinite, list_in, nlist_in, B, N, tol_rel, tol_abs, ind, list_out,
                                  nlist_out, reinterpret_cast<long long*>(packed), reinterpret_cast<const uint8_t*>(r_tiles),
                                  reinterpret_cast<uint8_t*>(c_tiles), c_tiles ? C / kTileK : 0, part_idx2, part_third, cand2, pair_list,
                                  npair, n_valid);
    IPSR_REQUIRE(le__ == cudaSuccess, IPSR_ERR_CUDA, "ipsr_finalize_argmax: launch failed: %s", cudaGetErrorString(le__));
  }
  return check_launch("ipsr_finalize_argmax");
}
