// (b)+(c) correlation GEMM on the 5th-generation tensor cores with a fused row (max, idx, runner-up)
// epilogue.  Replaces models/IPSRFunction.py:59 (conv_enc(ref) -> S[1,N,H,W] materialised) and
// util/MaxCoord.py:21-22 (zeros_like(S) + torch.max(S, 1)); S never leaves the SM.
//
// Arithmetic: S = R * Xn^T with both operands split into bf16 hi + lo parts; three tcgen05.mma passes
// (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) give |error| <~ 1e-5 * ||R[q]||, and rows whose
// top-2 gap is inside that band are recomputed in exact fp32 (ipsr_correlate_argmax_fp32).
//
// One CTA = 128 query rows (one TMEM lane each) x a range of bank columns, walked in blocks of
// BLOCK_N columns.  Warp roles (192 threads):
//   warp 0      producer: bulk async copies (UBLKCP) of pre-swizzled 16 KiB tile images -> smem ring
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue, commits free the ring slots
//   warps 2..5  epilogue: tcgen05.ld of the finished accumulator (double-buffered in TMEM) and the
//               running (best, idx, second) per row in registers -- thread == row, no shuffles.
// A_RESIDENT: the R row tile (C <= 256: 128 KiB hi+lo) stays in shared memory for the whole CTA and
// only bank tiles stream; otherwise both operands stream per 64-channel block.
#include "ipsr_common.cuh"

namespace ipsr {

constexpr int kTcThreads = 192;

struct TcParams {
  const uint8_t* r_tiles;
  const uint8_t* x_tiles;
  int B, KB, RB, N;
  int col_begin;        // first bank column (multiple of 128)
  int blocks_total;     // number of BLOCK_N column blocks in [col_begin, col_end)
  int psplit;
  int stages;
  float* part_best;
  int* part_idx;
  float* part_second;
  float* s_dump;
};

template <int BLOCK_N, bool A_RESIDENT>
__global__ void __launch_bounds__(kTcThreads, 1) corr_tc_kernel(const TcParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // manual 1024-byte alignment (SWIZZLE_128B atoms repeat every 1024 bytes)
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  constexpr int kBTiles = BLOCK_N / 128;                       // 16 KiB tiles per bank operand half
  constexpr uint32_t kStageBytes = (A_RESIDENT ? 0u : 2u * kTileBytes) + 2u * kBTiles * kTileBytes;
  const int KB = prm.KB, RB = prm.RB;
  const uint32_t a_bytes = A_RESIDENT ? (uint32_t)KB * 2u * kTileBytes : 0u;
  const uint32_t a_base = base;
  const uint32_t ring_base = base + a_bytes;
  const uint32_t bar_base = ring_base + (uint32_t)prm.stages * kStageBytes;
  // barriers: full[stages], empty[stages], a_full, tmem_full[2], tmem_empty[2]; then the TMEM address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (prm.stages + s); };
  const uint32_t a_full_bar = bar_base + 8u * (2 * prm.stages);
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * prm.stages + 1 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * prm.stages + 3 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * prm.stages + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work decomposition: blockIdx.x -> (b, rb, split)
  int t = blockIdx.x;
  const int split = t % prm.psplit; t /= prm.psplit;
  const int rb = t % RB;
  const int b = t / RB;
  const int per = (prm.blocks_total + prm.psplit - 1) / prm.psplit;
  const int blk0 = split * per;
  const int blk1 = min(prm.blocks_total, blk0 + per);
  const int nblk = max(0, blk1 - blk0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < prm.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(a_full_bar, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);     // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0 && nblk > 0) {
      if (A_RESIDENT) {
        mbar_expect_tx(a_full_bar, a_bytes);
        for (int kb = 0; kb < KB; ++kb)
          for (int hl = 0; hl < 2; ++hl)
            bulk_g2s(a_base + (uint32_t)(kb * 2 + hl) * kTileBytes,
                     prm.r_tiles + tile_offset_bytes(b, kb, hl, rb, KB, RB), kTileBytes, a_full_bar);
      }
      int it = 0;
      for (int blk = blk0; blk < blk1; ++blk) {
        const int cb = (prm.col_begin >> 7) + blk * kBTiles;     // first 128-row bank tile of the block
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % prm.stages;
          const uint32_t ph = (uint32_t)(it / prm.stages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), kStageBytes);
          uint32_t dst = ring_base + (uint32_t)s * kStageBytes;
          if (!A_RESIDENT) {
            for (int hl = 0; hl < 2; ++hl, dst += kTileBytes)
              bulk_g2s(dst, prm.r_tiles + tile_offset_bytes(b, kb, hl, rb, KB, RB), kTileBytes, full_bar(s));
          }
          for (int hl = 0; hl < 2; ++hl)
            for (int tl = 0; tl < kBTiles; ++tl, dst += kTileBytes)
              bulk_g2s(dst, prm.x_tiles + tile_offset_bytes(b, kb, hl, cb + tl, KB, RB), kTileBytes, full_bar(s));
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && nblk > 0) {
      const uint32_t idesc = umma_idesc_bf16(128, BLOCK_N);
      if (A_RESIDENT) {
        mbar_wait(a_full_bar, 0);
        tc_fence_after();
      }
      int it = 0;
      for (int j = 0; j < nblk; ++j) {
        const int as = j & 1;
        const uint32_t aph = (uint32_t)(j >> 1) & 1u;
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % prm.stages;
          const uint32_t ph = (uint32_t)(it / prm.stages) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t stage = ring_base + (uint32_t)s * kStageBytes;
          const uint32_t a_hi = A_RESIDENT ? a_base + (uint32_t)(kb * 2) * kTileBytes : stage;
          const uint32_t a_lo = a_hi + kTileBytes;
          const uint32_t b_hi = stage + (A_RESIDENT ? 0u : 2u * kTileBytes);
          const uint32_t b_lo = b_hi + kBTiles * kTileBytes;
          const uint64_t da[3] = {umma_desc_k_sw128(a_hi), umma_desc_k_sw128(a_hi), umma_desc_k_sw128(a_lo)};
          const uint64_t db[3] = {umma_desc_k_sw128(b_lo), umma_desc_k_sw128(b_hi), umma_desc_k_sw128(b_hi)};
          // small cross terms first, hi*hi last
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const int pp = (pass == 0) ? 0 : (pass == 1 ? 2 : 1);   // hi*lo, lo*hi, hi*hi
#pragma unroll
            for (int k = 0; k < kTileK / 16; ++k) {
              // +32 bytes per 16-element K step inside the 128-byte swizzle row (descriptor units of 16 B)
              umma_bf16(d_tmem, da[pp] + (uint64_t)(2 * k), db[pp] + (uint64_t)(2 * k), idesc,
                        (kb > 0 || pass > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(s));          // ring slot reusable once these MMAs retire
        }
        umma_commit(tfull_bar(as));           // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (thread == row)
    const int quad = warp & 3;                           // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    const int q = rb * 128 + row;
    float best = -INFINITY, second = -INFINITY;
    int bidx = prm.col_begin + blk0 * BLOCK_N;
    float* dump = prm.s_dump ? prm.s_dump + ((size_t)b * prm.N + q) * prm.N : nullptr;
    for (int j = 0; j < nblk; ++j) {
      const int as = j & 1;
      const uint32_t aph = (uint32_t)(j >> 1) & 1u;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const int colb = prm.col_begin + (blk0 + j) * BLOCK_N;
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BLOCK_N + ch * 32), r);
        tmem_ld_wait();
        const int c0 = colb + ch * 32;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float v = __uint_as_float(r[e]);
          const bool gt = v > best;                       // strict: the lowest column wins ties
          second = gt ? best : fmaxf(second, v);
          bidx = gt ? (c0 + e) : bidx;
          best = gt ? v : best;
        }
        if (dump) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dump + c0 + e) =
                make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]),
                            __uint_as_float(r[e + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
    const size_t o = ((size_t)split * prm.B + b) * prm.N + q;
    prm.part_best[o] = best;
    prm.part_idx[o] = bidx;
    prm.part_second[o] = second;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

// Merge the per-split triples and decide which rows the tensor result can be trusted for.
__global__ void finalize_kernel(const float* __restrict__ part_best, const int* __restrict__ part_idx,
                                const float* __restrict__ part_second, int psplit, const float* __restrict__ rnorm,
                                const int* __restrict__ nonfinite, int B, int N, float tol_rel, float tol_abs,
                                int* __restrict__ ind, int* __restrict__ list, int* __restrict__ nlist,
                                long long* __restrict__ packed) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= N) return;
  float best = -INFINITY, second = -INFINITY;
  int idx = 0;
  for (int s = 0; s < psplit; ++s) {                     // splits are in ascending column order
    const size_t o = ((size_t)s * B + b) * N + q;
    const float vb = part_best[o], vs = part_second[o];
    if (vb > best) {
      second = fmaxf(best, vs);
      best = vb;
      idx = part_idx[o];
    } else {
      second = fmaxf(second, vb);
    }
  }
  const float tol = fmaf(tol_rel, rnorm[(size_t)b * N + q], tol_abs);
  const bool bad = (nonfinite && nonfinite[b] != 0);
  // !(gap >= tol) also catches NaN and the all -inf row
  if (!bad && (best - second) >= tol) {
    ind[(size_t)b * N + q] = idx;
  } else {
    ind[(size_t)b * N + q] = idx;                        // overwritten by ipsr_apply_recheck
    const int pos = atomicAdd(nlist + b, 1);
    list[(size_t)b * N + pos] = q;
  }
  packed[(size_t)b * N + q] = kPackedIdentity;
}

static int tc_stage_count(int C, bool a_resident, int block_n, size_t* smem_out) {
  const size_t stage = (a_resident ? 0 : 2 * (size_t)kTileBytes) + 2 * (size_t)(block_n / 128) * kTileBytes;
  const size_t a_bytes = a_resident ? (size_t)(C / kTileK) * 2 * kTileBytes : 0;
  const size_t fixed = a_bytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
  const size_t budget = 227 * 1024;
  int stages = (int)((budget - fixed) / stage);
  if (stages > 6) stages = 6;
  *smem_out = fixed + (size_t)stages * stage;
  return stages;
}

}  // namespace ipsr

extern "C" int ipsr_tensor_path_supported(int C, int N) {
  return (C > 0 && N > 0 && C % ipsr::kTileK == 0 && N % ipsr::kTileRows == 0 && N <= 65536) ? 1 : 0;
}

extern "C" int ipsr_correlate_argmax_tc(const void* r_tiles, const void* x_tiles, int B, int C, int N,
                                        int col_begin, int col_end, int psplit,
                                        float* part_best, int32_t* part_idx, float* part_second,
                                        float* s_dump, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(r_tiles && x_tiles && part_best && part_idx && part_second, IPSR_ERR_INVALID_ARG,
               "ipsr_correlate_argmax_tc: null pointer");
  IPSR_REQUIRE(ipsr_tensor_path_supported(C, N), IPSR_ERR_UNSUPPORTED,
               "ipsr_correlate_argmax_tc: shape C=%d N=%d not supported (need C %% 64 == 0, N %% 128 == 0)", C, N);
  IPSR_REQUIRE(B > 0 && col_begin >= 0 && col_end <= N && col_begin < col_end && col_begin % 128 == 0 &&
                   col_end % 128 == 0 && psplit >= 1,
               IPSR_ERR_INVALID_ARG, "ipsr_correlate_argmax_tc: bad column range [%d,%d) psplit=%d", col_begin, col_end, psplit);
  const bool a_res = (C <= 256);
  const int ncols = col_end - col_begin;
  // A resident: BLOCK_N = 128 keeps 3+ ring stages next to the 128 KiB row tile; streaming: 256.
  int block_n = a_res ? 128 : 256;
  if (ncols % block_n != 0) block_n = 128;
  TcParams prm;
  prm.r_tiles = reinterpret_cast<const uint8_t*>(r_tiles);
  prm.x_tiles = reinterpret_cast<const uint8_t*>(x_tiles);
  prm.B = B; prm.KB = C / kTileK; prm.RB = N / kTileRows; prm.N = N;
  prm.col_begin = col_begin;
  prm.blocks_total = ncols / block_n;
  // psplit > blocks_total is allowed: the surplus splits own no column block and report -inf
  prm.psplit = psplit;
  prm.part_best = part_best; prm.part_idx = part_idx; prm.part_second = part_second; prm.s_dump = s_dump;
  size_t smem = 0;
  prm.stages = tc_stage_count(C, a_res, block_n, &smem);
  IPSR_REQUIRE(prm.stages >= 2, IPSR_ERR_UNSUPPORTED, "ipsr_correlate_argmax_tc: C=%d leaves %d pipeline stages", C, prm.stages);
  const long long ctas = (long long)B * prm.RB * psplit;
  IPSR_REQUIRE(ctas <= 0x7FFFFFFFll, IPSR_ERR_UNSUPPORTED, "ipsr_correlate_argmax_tc: grid too large");

  void (*kern)(const TcParams) = nullptr;
  if (a_res && block_n == 128) kern = corr_tc_kernel<128, true>;
  else if (!a_res && block_n == 256) kern = corr_tc_kernel<256, false>;
  else if (!a_res && block_n == 128) kern = corr_tc_kernel<128, false>;
  IPSR_REQUIRE(kern != nullptr, IPSR_ERR_UNSUPPORTED, "ipsr_correlate_argmax_tc: no kernel variant");
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "corr_tc smem attribute (%zu B): %s", smem, cudaGetErrorString(e));
  kern<<<(unsigned)ctas, kTcThreads, smem, as_stream(stream)>>>(prm);
  return check_launch("ipsr_correlate_argmax_tc");
}

extern "C" int ipsr_finalize_argmax(const float* part_best, const int32_t* part_idx, const float* part_second,
                                    int psplit, const float* rnorm, const int32_t* nonfinite,
                                    int B, int N, float tol_rel, float tol_abs,
                                    int32_t* ind, int32_t* recheck_list, int32_t* nrecheck, int64_t* packed,
                                    void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(part_best && part_idx && part_second && rnorm && ind && recheck_list && nrecheck && packed,
               IPSR_ERR_INVALID_ARG, "ipsr_finalize_argmax: null pointer");
  IPSR_REQUIRE(B > 0 && N > 0 && psplit >= 1, IPSR_ERR_INVALID_ARG, "ipsr_finalize_argmax: bad dims");
  finalize_kernel<<<dim3((N + 255) / 256, B), 256, 0, as_stream(stream)>>>(
      part_best, part_idx, part_second, psplit, rnorm, nonfinite, B, N, tol_rel, tol_abs, ind, recheck_list, nrecheck,
      reinterpret_cast<long long*>(packed));
  return check_launch("ipsr_finalize_argmax");
}
