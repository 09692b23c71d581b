// Patch mode (shift_sz = k > 1) with LONG patch rows (K = C*k*k > 1024, e.g. C = 256, k = 3: K = 2304) on the tensor
// cores -- BASELINE.json configs[3].  models/IPSRFunction.py:54-59 with shift_sz = 3: the correlation is
// Conv2d(C, P, k) of ref with the normalised patches of x, i.e. S[q,p] = <R_patch[q], Xn_patch[p]> over K values, for
// P = nH*nW patch positions (P = 3844 for a 64 x 64 map: not a multiple of 128).
//
//   ipsr_patch_rows_stats   (ipsr_patch.cu) position-major patch rows [B][P][K] of x and ref, norms, row maxima
//   ipsr_patch_tiles        rows -> fp16 hi/lo operand images of the tcgen05 GEMM, [B][Kpad/64][2][Ppad/128][128 x 64],
//                           zero rows / columns up to Ppad = ceil128(P), Kpad = ceil64(K); same power-of-two scaling
//                           as the 1 x 1 prep (ipsr_prep.cu): Xs = Xn 2^11, Rs = R 2^s_q with max |Rs| in [2^13, 2^14)
//   ipsr_correlate_argmax_tc_valid (ipsr_corr_tc.cu) the three-pass split hi*lo + lo*hi + hi*hi over every row, both
//                           operands streamed (the row tile of K = 2304 does not fit in shared memory), padding
//                           columns masked to -inf in the epilogue
//   ipsr_finalize_argmax_valid     rows whose top-2 gap is inside the error band of the split go to a list ...
//   ipsr_patch_recheck      ... and are recomputed against every bank column in fp32 from the rows (warp per column,
//                           lanes along K), keys merged by atomicMax exactly like ipsr_correlate_argmax_fp32
//   ipsr_patch_winner_scores exact fp32 score of every row's winner = vmax of the blend (IPSRFunction.py:70) and the
//                           (score, index) key of the bank-sharded exchange
#include <cuda_fp16.h>

#include "ipsr_common.cuh"

namespace ipsr {

// 2^s with max|v| * 2^s in [2^13, 2^14); 1 for a zero / non-finite row  (as ipsr_prep.cu)
__device__ __forceinline__ float patch_row_scale_pow2(float vmax) {
  if (!(vmax > 0.f) || !(vmax <= 3.4028234e38f)) return 1.0f;
  int e;
  frexpf(vmax, &e);
  int s = 14 - e;
  s = max(-100, min(100, s));
  return ldexpf(1.0f, s);
}

// grid = (Ppad / 128, Kpad / 64, B), 256 threads: one (hi, lo) tile pair per CTA; a thread converts four 8-value chunks.
// is_ref = 0: v = fl(fl(x * inv_norm) * 2^11);  is_ref = 1: v = r * 2^s_q, rscale[b][q] = 2^-(s_q + 11) (stride Ppad).
__global__ void __launch_bounds__(256)
patch_tiles_kernel(const float* __restrict__ rows, const float* __restrict__ inv_norm, const float* __restrict__ maxabs,
                   const float* __restrict__ norm, int is_ref, int K, int P, int Kpad, int Ppad, uint8_t* __restrict__ tiles,
                   float* __restrict__ rscale, float* __restrict__ rnorm_pad) {
  const int rb = blockIdx.x, kb = blockIdx.y, b = blockIdx.z;
  const int KB = Kpad / kTileK, RB = Ppad / kTileRows;
  uint8_t* hi_t = tiles + tile_offset_bytes(b, kb, 0, rb, KB, RB);
  uint8_t* lo_t = tiles + tile_offset_bytes(b, kb, 1, rb, KB, RB);
  for (int it = threadIdx.x; it < kTileRows * 8; it += blockDim.x) {
    const int r = it >> 3, chunk = it & 7;
    const int p = rb * kTileRows + r, k0 = kb * kTileK + chunk * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (p < P) {
      float sc;
      if (is_ref) {
        sc = patch_row_scale_pow2(maxabs[(size_t)b * P + p]);
        if (kb == 0 && chunk == 0) {
          if (rscale) rscale[(size_t)b * Ppad + p] = __fdiv_rn(0.00048828125f, sc);   // 2^-11 / 2^s, exact
          if (rnorm_pad) rnorm_pad[(size_t)b * Ppad + p] = norm[(size_t)b * P + p];
        }
      } else {
        sc = inv_norm[(size_t)b * P + p];
      }
      const float* src = rows + ((size_t)b * P + p) * K + k0;
      if (k0 + 8 <= K && (K & 3) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), c = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (k0 + i < K) ? __ldg(src + i) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)      // Xn = fl(X * inv) as the reference's encoder weights hold it (NPS:40), then the exact 2^11
        v[i] = is_ref ? __fmul_rn(v[i], sc) : __fmul_rn(__fmul_rn(v[i], sc), 2048.0f);
    } else if (is_ref && kb == 0 && chunk == 0 && p < Ppad) {
      if (rscale) rscale[(size_t)b * Ppad + p] = 1.0f;
      if (rnorm_pad) rnorm_pad[(size_t)b * Ppad + p] = 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half h0 = __float2half_rn(v[2 * i]), h1 = __float2half_rn(v[2 * i + 1]);
      const float d0 = v[2 * i] - __half2float(h0), d1 = v[2 * i + 1] - __half2float(h1);
      hi[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      lo[i] = (uint32_t)__half_as_ushort(__float2half_rn(d0)) | ((uint32_t)__half_as_ushort(__float2half_rn(d1)) << 16);
    }
    const uint32_t off = tile_chunk_offset(r, chunk);
    *reinterpret_cast<uint4*>(hi_t + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(lo_t + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// exact fp32 <R[q], fl(X[p] * inv[p])> with lanes along K and a fixed xor tree (identical rows give identical sums, so
// exact ties keep the lowest column, util/MaxCoord.py:22)
__device__ __forceinline__ float patch_dot(const float* __restrict__ r, const float* __restrict__ x, float inv, int K, int lane) {
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(r + k), __fmul_rn(__ldg(x + k), inv), acc);
  return warp_sum(acc);
}

// grid = (column chunks of 64, B, row groups), 256 threads = 8 warps; every warp takes 8 columns of the chunk per listed row.
__global__ void __launch_bounds__(256)
patch_recheck_kernel(const float* __restrict__ rows_x, const float* __restrict__ rows_r, const float* __restrict__ inv_norm,
                     int K, int P, int Ppad, int col_begin, int col_end, const int* __restrict__ list,
                     const int* __restrict__ nlist, long long* __restrict__ packed) {
  const int b = blockIdx.y;
  const int n = min(nlist[b], P);
  if (n == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p0 = col_begin + blockIdx.x * 64;
  for (int i = blockIdx.z; i < n; i += gridDim.z) {        // listed rows are dealt to blockIdx.z
    const int q = list[(size_t)b * Ppad + i];
    const float* rq = rows_r + ((size_t)b * P + q) * K;
    long long best = kPackedIdentity;
    for (int j = warp; j < 64; j += 8) {
      const int p = p0 + j;
      if (p >= col_end) break;
      const float sc = patch_dot(rq, rows_x + ((size_t)b * P + p) * K, inv_norm[(size_t)b * P + p], K, lane);
      const long long key = pack_maxidx(sc, p);
      best = key > best ? key : best;
    }
    if (lane == 0 && best != kPackedIdentity) atomicMax(packed + (size_t)b * Ppad + q, best);
  }
}

// Rows with exactly two candidates inside the error band of the split (ipsr_finalize_argmax: pair_list, cand2): one warp
// computes both exact scores and keeps the larger (the lower column on a tie); the key goes to packed (stride Ppad).
__global__ void __launch_bounds__(256)
patch_pairs_kernel(const float* __restrict__ rows_x, const float* __restrict__ rows_r, const float* __restrict__ inv_norm,
                   int K, int P, int Ppad, const int* __restrict__ pair_list, const int* __restrict__ npair,
                   const int* __restrict__ cand2, int* __restrict__ ind_pad, long long* __restrict__ packed) {
  const int b = blockIdx.y;
  const int np = min(npair[b], P);
  const int lane = threadIdx.x & 31;
  for (int j = blockIdx.x * 8 + (threadIdx.x >> 5); j < np; j += gridDim.x * 8) {
    const int q = pair_list[(size_t)b * Ppad + j];
    const int p1 = ind_pad[(size_t)b * Ppad + q], p2 = cand2[(size_t)b * Ppad + q];
    const float* rq = rows_r + ((size_t)b * P + q) * K;
    const float a1 = patch_dot(rq, rows_x + ((size_t)b * P + p1) * K, inv_norm[(size_t)b * P + p1], K, lane);
    const float a2 = patch_dot(rq, rows_x + ((size_t)b * P + p2) * K, inv_norm[(size_t)b * P + p2], K, lane);
    if (lane == 0) {
      const bool second_wins = (a2 > a1) || (a2 == a1 && p2 < p1);
      ind_pad[(size_t)b * Ppad + q] = second_wins ? p2 : p1;
      packed[(size_t)b * Ppad + q] = pack_maxidx(second_wins ? a2 : a1, second_wins ? p2 : p1);
    }
  }
}

// one warp per row q < P: the exact score of its winner -> vmax [B][P], ind_out [B][P], keys [B][P] (optional).
// Rows already holding an exact key from the recheck (packed != identity) keep it.
__global__ void __launch_bounds__(256)
patch_winner_kernel(const float* __restrict__ rows_x, const float* __restrict__ rows_r, const float* __restrict__ inv_norm,
                    const int* __restrict__ ind_pad, const long long* __restrict__ packed_pad, int K, int P, int Ppad,
                    int* __restrict__ ind_out, float* __restrict__ vmax, long long* __restrict__ keys) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= P) return;
  long long key = packed_pad ? packed_pad[(size_t)b * Ppad + q] : kPackedIdentity;
  float v;
  int p;
  if (key != kPackedIdentity) {
    unpack_maxidx(key, &v, &p);
  } else {
    p = ind_pad[(size_t)b * Ppad + q];
    v = patch_dot(rows_r + ((size_t)b * P + q) * K, rows_x + ((size_t)b * P + p) * K, inv_norm[(size_t)b * P + p], K, lane);
    key = pack_maxidx(v, p);
  }
  if (lane == 0) {
    if (ind_out) ind_out[(size_t)b * P + q] = p;
    if (vmax) vmax[(size_t)b * P + q] = v;
    if (keys) keys[(size_t)b * P + q] = key;
  }
}

}  // namespace ipsr

extern "C" int ipsr_patch_tiles(const float* rows, const float* inv_norm, const float* maxabs, const float* norm, int is_ref,
                                int B, int K, int P, void* tiles, float* rscale, float* rnorm_pad, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(rows && tiles && (is_ref ? (maxabs != nullptr) : (inv_norm != nullptr)) && (!rnorm_pad || norm), IPSR_ERR_INVALID_ARG,
               "ipsr_patch_tiles: null pointer");
  IPSR_REQUIRE(B > 0 && B <= 65535 && K > 0 && P > 0, IPSR_ERR_INVALID_ARG, "ipsr_patch_tiles: bad dims");
  const int Kpad = (K + kTileK - 1) / kTileK * kTileK, Ppad = (P + kTileRows - 1) / kTileRows * kTileRows;
  IPSR_REQUIRE(Kpad / kTileK <= 65535, IPSR_ERR_UNSUPPORTED, "ipsr_patch_tiles: K=%d too large", K);
  patch_tiles_kernel<<<dim3(Ppad / kTileRows, Kpad / kTileK, B), 256, 0, as_stream(stream)>>>(
      rows, inv_norm, maxabs, norm, is_ref, K, P, Kpad, Ppad, reinterpret_cast<uint8_t*>(tiles), rscale, rnorm_pad);
  return check_launch("ipsr_patch_tiles");
}

extern "C" int ipsr_patch_recheck(const float* rows_x, const float* rows_r, const float* inv_norm, int B, int K, int P,
                                  int col_begin, int col_end, const int32_t* list, const int32_t* nlist, int64_t* packed,
                                  void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(rows_x && rows_r && inv_norm && list && nlist && packed, IPSR_ERR_INVALID_ARG, "ipsr_patch_recheck: null pointer");
  IPSR_REQUIRE(B > 0 && B <= 65535 && K > 0 && P > 0 && col_begin >= 0 && col_begin < col_end && col_end <= P, IPSR_ERR_INVALID_ARG,
               "ipsr_patch_recheck: bad dims / column range [%d,%d)", col_begin, col_end);
  const int Ppad = (P + kTileRows - 1) / kTileRows * kTileRows;
  patch_recheck_kernel<<<dim3((col_end - col_begin + 63) / 64, B, 16), 256, 0, as_stream(stream)>>>(
      rows_x, rows_r, inv_norm, K, P, Ppad, col_begin, col_end, list, nlist, reinterpret_cast<long long*>(packed));
  return check_launch("ipsr_patch_recheck");
}

extern "C" int ipsr_patch_resolve_pairs(const float* rows_x, const float* rows_r, const float* inv_norm, int B, int K, int P,
                                        const int32_t* pair_list, const int32_t* npair, const int32_t* cand2, int32_t* ind_pad,
                                        int64_t* packed, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(rows_x && rows_r && inv_norm && pair_list && npair && cand2 && ind_pad && packed, IPSR_ERR_INVALID_ARG,
               "ipsr_patch_resolve_pairs: null pointer");
  IPSR_REQUIRE(B > 0 && B <= 65535 && K > 0 && P > 0, IPSR_ERR_INVALID_ARG, "ipsr_patch_resolve_pairs: bad dims");
  const int Ppad = (P + kTileRows - 1) / kTileRows * kTileRows;
  int G = (P + 7) / 8;
  if (G > 296) G = 296;
  patch_pairs_kernel<<<dim3(G, B), 256, 0, as_stream(stream)>>>(rows_x, rows_r, inv_norm, K, P, Ppad, pair_list, npair, cand2, ind_pad,
                                                             reinterpret_cast<long long*>(packed));
  return check_launch("ipsr_patch_resolve_pairs");
}

extern "C" int ipsr_patch_winner_scores(const float* rows_x, const float* rows_r, const float* inv_norm, const int32_t* ind_pad,
                                        const int64_t* packed_pad, int B, int K, int P, int32_t* ind, float* vmax, int64_t* keys,
                                        void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(rows_x && rows_r && inv_norm && ind_pad, IPSR_ERR_INVALID_ARG, "ipsr_patch_winner_scores: null pointer");
  IPSR_REQUIRE(B > 0 && B <= 65535 && K > 0 && P > 0, IPSR_ERR_INVALID_ARG, "ipsr_patch_winner_scores: bad dims");
  const int Ppad = (P + kTileRows - 1) / kTileRows * kTileRows;
  patch_winner_kernel<<<dim3((P + 7) / 8, B), 256, 0, as_stream(stream)>>>(rows_x, rows_r, inv_norm, ind_pad,
                                                                         reinterpret_cast<const long long*>(packed_pad), K, P, Ppad,
                                                                         ind, vmax, reinterpret_cast<long long*>(keys));
  return check_launch("ipsr_patch_winner_scores");
}
