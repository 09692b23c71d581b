// (a) patch extraction + L2 normalisation + fp16 operand staging for the tcgen05 correlation.
//
// Replaces util/NonparametricShift.py:36-40 (per-patch `p * (1/(||p||+1e-8))` in a python loop)
// and :59-73 (unfold/permute/index_select copies) for patch_size = stride = 1.
//
// One CTA = one tensor (x or ref) x one image x 32 consecutive positions x all C channels.
//   phase 1: coalesced 128-byte row reads of the NCHW map into a [C][33] shared tile, running sum of
//            squares (and, for ref, the maximum magnitude) per position;
//   phase 2: every warp owns 8 positions x 32 channels per pass (bank-conflict-free smem reads) and
//            emits 16/32-byte vector stores: the position-major fp32 copy, the masked rows of ref, and
//            the fp16 operands written DIRECTLY in the 128B-swizzled UMMA tile image the tcgen05
//            GEMM streams with bulk copies (no tensor map, no second transposition pass).
//
// fp16 operands (11-bit significands, range 6e-8 .. 65504) are pre-scaled by powers of two, which the
// arg-max does not see:  Xs = Xn * 2^11 (|Xn| <= 1), hi = fp16(Xs), lo = fp16(Xs - hi);
//                        Rs = R[q] * 2^s_q with max_c |Rs| in [2^13, 2^14), hi = fp16(Rs), lo = fp16(Rs - hi).
// true score = tensor score * rscale[q], rscale = 2^-(s_q + 11).  The exact rounding error of the hi parts,
//   rerr[q] = ||R[q] - hi(Rs) 2^-s_q||_2,  xerr[p] = ||Xn[p] - hi(Xs) 2^-11||_2,  xerr_max[b] = max_p xerr[p],
// gives the rigorous error bound of the single-pass correlation: |S~ - S| <= rerr[q] + ||R~[q]|| xerr[p].
// HBM-bound: algorithmic bytes per image = 2*4NC read + (4NC + 2*2*2NC) write (+4MC masked rows).
#include <cuda_fp16.h>

#include "ipsr_common.cuh"

namespace ipsr {

constexpr int kPrepThreads = 256;
constexpr int kPrepPos = 32;

// 8 consecutive channels of one row -> fp16 hi (and lo) 16-byte chunks of the tile image; returns the
// squared rounding error of the hi part in SCALED units.
__device__ __forceinline__ float
store_fp16_chunk(const float (&v)[8], uint8_t* __restrict__ tiles, int nhalf, int b, int c0, int row, int KB, int RB) {
  uint32_t hi[4], lo[4];
  float err2 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half h0 = __float2half_rn(v[2 * i]), h1 = __float2half_rn(v[2 * i + 1]);
    const float d0 = v[2 * i] - __half2float(h0), d1 = v[2 * i + 1] - __half2float(h1);
    err2 = fmaf(d0, d0, err2);
    err2 = fmaf(d1, d1, err2);
    hi[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    lo[i] = (uint32_t)__half_as_ushort(__float2half_rn(d0)) | ((uint32_t)__half_as_ushort(__float2half_rn(d1)) << 16);
  }
  const int kb = c0 / kTileK, chunk = (c0 % kTileK) / 8;
  const int rb = row / kTileRows, r = row % kTileRows;
  const uint32_t off = tile_chunk_offset(r, chunk);
  *reinterpret_cast<uint4*>(tiles + tile_offset_bytes_n(b, kb, 0, rb, KB, RB, nhalf) + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (nhalf == 2)
    *reinterpret_cast<uint4*>(tiles + tile_offset_bytes_n(b, kb, 1, rb, KB, RB, nhalf) + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  return err2;
}

// 2^s with max|v| * 2^s in [2^13, 2^14); 1 for a zero / non-finite row
__device__ __forceinline__ float row_scale_pow2(float vmax) {
  if (!(vmax > 0.f) || !(vmax <= 3.4028234e38f)) return 1.0f;
  int e;
  frexpf(vmax, &e);                            // vmax = m * 2^e, m in [0.5, 1)
  int s = 14 - e;
  s = max(-100, min(100, s));
  return ldexpf(1.0f, s);
}

__global__ void __launch_bounds__(kPrepThreads)
prep_kernel(const float* __restrict__ x, const float* __restrict__ ref, int C, int N,
            const int* __restrict__ rank, int M,
            float* __restrict__ inv_norm, float* __restrict__ rnorm, float* __restrict__ xt,
            float* __restrict__ r_masked, uint8_t* __restrict__ x_tiles, uint8_t* __restrict__ r_tiles,
            int* __restrict__ nonfinite, float* __restrict__ rscale, float* __restrict__ rerr,
            float* __restrict__ xerr, int* __restrict__ xerr_max, int ms) {
  extern __shared__ float smem[];
  pdl_trigger();
  pdl_wait();
  float* tile = smem;                      // [C][33]
  float* part = smem + (size_t)C * 33;     // [8][32] sums of squares, then [8][32] maxima
  float* scale = part + 2 * 8 * 32;        // [32]
  float* errp = scale + 32;                // [C/8][32] squared rounding errors per channel group

  const int is_ref = blockIdx.z;
  const int b = blockIdx.y;
  if (rank) rank += (size_t)b * ms;                  // per-image masks: rank is [B][ms]
  const int p0 = blockIdx.x * kPrepPos;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* src = (is_ref ? ref : x) + (size_t)b * C * N;
  uint8_t* tiles = is_ref ? r_tiles : x_tiles;

  // ---- phase 1: load + sum of squares (+ maximum magnitude) ----
  {
    const int p = p0 + lane;
    const bool ok = p < N;
    float ss = 0.f, mx = 0.f;
#pragma unroll 8
    for (int c = warp; c < C; c += 8) {
      float v = ok ? __ldg(src + (size_t)c * N + p) : 0.f;
      tile[c * 33 + lane] = v;
      ss = fmaf(v, v, ss);
      mx = fmaxf(mx, fabsf(v));
    }
    part[warp * 32 + lane] = ss;
    part[256 + warp * 32 + lane] = mx;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float tot = 0.f, mx = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      tot += part[w * 32 + threadIdx.x];
      mx = fmaxf(mx, part[256 + w * 32 + threadIdx.x]);
    }
    const float nrm = sqrtf(tot);
    const int p = p0 + threadIdx.x;
    if (nonfinite && !(fabsf(tot) <= 3.4028234e38f)) atomicOr(nonfinite + b, 1);
    if (is_ref) {
      const float sc = row_scale_pow2(mx);
      scale[threadIdx.x] = sc;
      if (p < N && rnorm) rnorm[(size_t)b * N + p] = nrm;
      if (p < N && rscale) rscale[(size_t)b * N + p] = __fdiv_rn(0.00048828125f, sc);      // 2^-11 / 2^s, exact
    } else {
      const float inv = __fdiv_rn(1.0f, nrm + 1e-8f);   // NPS:40  1/(norm+1e-8)
      scale[threadIdx.x] = inv;
      if (p < N) inv_norm[(size_t)b * N + p] = inv;
    }
  }
  __syncthreads();

  // ---- phase 2: transposed vector stores ----
  const int pp = lane >> 2, j = lane & 3;
  const int KB = C / kTileK, RB = N / kTileRows;
  const int passes = 4 * ((C + 31) / 32);
  for (int pass = warp; pass < passes; pass += 8) {
    const int g = pass & 3, cg = pass >> 2;
    const int pl = g * 8 + pp;
    const int p = p0 + pl;
    const int c0 = cg * 32 + j * 8;
    if (c0 >= C) continue;
    float e2 = 0.f;
    if (p < N) {
      float raw[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) raw[i] = tile[(c0 + i) * 33 + pl];
      const float sc = scale[pl];
      if (!is_ref) {
        if (xt) {
          float4* dst = reinterpret_cast<float4*>(xt + ((size_t)b * N + p) * C + c0);
          dst[0] = make_float4(raw[0], raw[1], raw[2], raw[3]);
          dst[1] = make_float4(raw[4], raw[5], raw[6], raw[7]);
        }
      } else if (r_masked) {
        const int l = rank[p];
        if (l >= 0) {
          float4* dst = reinterpret_cast<float4*>(r_masked + ((size_t)b * M + l) * C + c0);
          dst[0] = make_float4(raw[0], raw[1], raw[2], raw[3]);
          dst[1] = make_float4(raw[4], raw[5], raw[6], raw[7]);
        }
      }
      if (tiles) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)      // Xn = fl(X * inv) as the reference's encoder weights hold it, then the exact 2^11
          v[i] = is_ref ? __fmul_rn(raw[i], sc) : __fmul_rn(__fmul_rn(raw[i], sc), 2048.0f);
        e2 = store_fp16_chunk(v, tiles, 2, b, c0, p, KB, RB);
      }
    }
    if (tiles) errp[(c0 >> 3) * 32 + pl] = e2;
  }
  if (!tiles) return;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int p = p0 + threadIdx.x;
    float tot = 0.f;
    for (int gch = 0; gch < C / 8; ++gch) tot += errp[gch * 32 + threadIdx.x];          // fixed order: deterministic
    if (p < N) {
      // back to true units, rounded up a little so that the bound stays a bound
      const float unscale = is_ref ? __fdiv_rn(1.0f, scale[threadIdx.x]) : 0.00048828125f;
      const float err = __fmul_rn(__fmul_rn(sqrtf(tot), unscale), 1.0001f);
      if (is_ref) {
        if (rerr) rerr[(size_t)b * N + p] = err;
      } else {
        if (xerr) xerr[(size_t)b * N + p] = err;
        if (xerr_max && err == err) atomicMax(xerr_max + b, __float_as_int(err));          // err >= 0: int order == float order
      }
    }
  }
}

// Rows of ref listed in `list` (n = nlist[b] of them) -> compact fp16 hi AND lo tile images for the 3-pass
// correlation of the ambiguous rows: compact row r of image b = position list[b][r]; rows up to the next
// multiple of 128 are zero-filled.  grid = (N / 32, B); CTAs beyond the list exit at once.
__global__ void __launch_bounds__(kPrepThreads)
compact_rows_kernel(const float* __restrict__ ref, int C, int N, const int* __restrict__ list, const int* __restrict__ nlist,
                    const float* __restrict__ rscale, uint8_t* __restrict__ c_tiles) {
  extern __shared__ float smem[];
  float* tile = smem;                      // [C][33]
  __shared__ int rows[kPrepPos];
  __shared__ float scale[kPrepPos];
  const int b = blockIdx.y;
  const int n = min(nlist[b], N);
  const int r0 = blockIdx.x * kPrepPos;
  if (r0 >= ((n + kTileRows - 1) / kTileRows) * kTileRows) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kPrepPos) {
    const int r = r0 + threadIdx.x;
    const int q = (r < n) ? list[(size_t)b * N + r] : -1;
    rows[threadIdx.x] = q;
    scale[threadIdx.x] = (q >= 0) ? __fdiv_rn(0.00048828125f, rscale[(size_t)b * N + q]) : 0.f;   // 2^s_q
  }
  __syncthreads();
  const float* src = ref + (size_t)b * C * N;
  {
    const int q = rows[lane];
#pragma unroll 8
    for (int c = warp; c < C; c += 8) tile[c * 33 + lane] = (q >= 0) ? __ldg(src + (size_t)c * N + q) : 0.f;
  }
  __syncthreads();
  const int pp = lane >> 2, j = lane & 3;
  const int KB = C / kTileK, RB = N / kTileRows;
  const int passes = 4 * ((C + 31) / 32);
  for (int pass = warp; pass < passes; pass += 8) {
    const int g = pass & 3, cg = pass >> 2;
    const int pl = g * 8 + pp;
    const int c0 = cg * 32 + j * 8;
    if (c0 >= C) continue;
    const float sc = scale[pl];
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __fmul_rn(tile[(c0 + i) * 33 + pl], sc);
    store_fp16_chunk(v, c_tiles, 2, b, c0, r0 + pl, KB, RB);
  }
}

}  // namespace ipsr

extern "C" int ipsr_extract_normalize(const float* x, const float* ref, int B, int C, int N,
                                      const int32_t* rank_i32, int M,
                                      float* inv_norm, float* rnorm, float* xt, float* r_masked,
                                      void* x_tiles, void* r_tiles, int32_t* nonfinite,
                                      float* rscale, float* rerr, float* xerr, float* xerr_max, void* stream) {
  return ipsr::extract_normalize_ex(x, ref, B, C, N, rank_i32, M, inv_norm, rnorm, xt, r_masked, x_tiles, r_tiles, nonfinite, rscale,
                                    rerr, xerr, xerr_max, stream, 0);
}

int ipsr::extract_normalize_ex(const float* x, const float* ref, int B, int C, int N, const int32_t* rank_i32, int M,
                               float* inv_norm, float* rnorm, float* xt, float* r_masked, void* x_tiles, void* r_tiles,
                               int32_t* nonfinite, float* rscale, float* rerr, float* xerr, float* xerr_max, void* stream,
                               int ms) {
  using namespace ipsr;
  IPSR_REQUIRE(x && ref && inv_norm, IPSR_ERR_INVALID_ARG, "ipsr_extract_normalize: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0 && B <= 65535, IPSR_ERR_INVALID_ARG, "ipsr_extract_normalize: bad dims B=%d C=%d N=%d", B, C, N);
  IPSR_REQUIRE(C % 8 == 0, IPSR_ERR_UNSUPPORTED, "ipsr_extract_normalize: C=%d must be a multiple of 8", C);
  IPSR_REQUIRE(!(r_masked && M > 0) || rank_i32, IPSR_ERR_INVALID_ARG, "ipsr_extract_normalize: r_masked needs rank");
  if (x_tiles || r_tiles) {
    IPSR_REQUIRE(C % kTileK == 0 && N % kTileRows == 0, IPSR_ERR_UNSUPPORTED,
                 "ipsr_extract_normalize: tile images need C %% 64 == 0 and N %% 128 == 0 (C=%d N=%d)", C, N);
    IPSR_REQUIRE(!r_tiles || (rscale && rerr), IPSR_ERR_INVALID_ARG, "ipsr_extract_normalize: r_tiles needs rscale and rerr");
    IPSR_REQUIRE(!x_tiles || xerr_max, IPSR_ERR_INVALID_ARG, "ipsr_extract_normalize: x_tiles needs xerr_max");
  }
  const size_t smem = ((size_t)C * 33 + 2 * 8 * 32 + 32 + (size_t)(C / 8) * 32) * sizeof(float);
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_extract_normalize: C=%d too large", C);
  if (smem + 2048 > 48 * 1024) {   // dynamic + static shared memory above the default limit (set per call: the attribute is per device)
    cudaError_t e = cudaFuncSetAttribute(prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "prep smem attribute: %s", cudaGetErrorString(e));
  }
  dim3 grid((N + kPrepPos - 1) / kPrepPos, B, 2);
  {
    cudaError_t le__ = launch_pdl(prep_kernel, grid, dim3(kPrepThreads), smem, as_stream(stream), x, ref, C, N, rank_i32, M, inv_norm, rnorm, xt,
                                  (M > 0 ? r_masked : nullptr), reinterpret_cast<uint8_t*>(x_tiles), reinterpret_cast<uint8_t*>(r_tiles),
                                  nonfinite, rscale, rerr, xerr, reinterpret_cast<int*>(xerr_max), ms);
    IPSR_REQUIRE(le__ == cudaSuccess, IPSR_ERR_CUDA, "ipsr_extract_normalize: launch failed: %s", cudaGetErrorString(le__));
  }
  return check_launch("ipsr_extract_normalize");
}

extern "C" int ipsr_compact_rows(const float* ref, int B, int C, int N, const int32_t* list, const int32_t* nlist,
                                 const float* rscale, void* c_tiles, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(ref && list && nlist && rscale && c_tiles, IPSR_ERR_INVALID_ARG, "ipsr_compact_rows: null pointer");
  IPSR_REQUIRE(B > 0 && B <= 65535 && C > 0 && N > 0 && C % kTileK == 0 && N % kTileRows == 0, IPSR_ERR_UNSUPPORTED,
               "ipsr_compact_rows: need C %% 64 == 0 and N %% 128 == 0 (C=%d N=%d)", C, N);
  const size_t smem = (size_t)C * 33 * sizeof(float);
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_compact_rows: C=%d too large", C);
  if (smem + 2048 > 48 * 1024) {   // dynamic + static shared memory above the default limit (set per call: the attribute is per device)
    cudaError_t e = cudaFuncSetAttribute(compact_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "compact_rows smem attribute: %s", cudaGetErrorString(e));
  }
  compact_rows_kernel<<<dim3(N / kPrepPos, B), kPrepThreads, smem, as_stream(stream)>>>(
      ref, C, N, list, nlist, rscale, reinterpret_cast<uint8_t*>(c_tiles));
  return check_launch("ipsr_compact_rows");
}
