// (a) patch extraction + L2 normalisation + bf16 hi/lo operand staging.
//
// Replaces util/NonparametricShift.py:36-40 (per-patch `p * (1/(||p||+1e-8))` in a python loop)
// and :59-73 (unfold/permute/index_select copies) for patch_size = stride = 1.
//
// One CTA = one tensor (x or ref) x one image x 32 consecutive positions x all C channels.
//   phase 1: coalesced 128-byte row reads of the NCHW map into a [C][33] shared tile, running sum of
//            squares per position;
//   phase 2: every warp owns 8 positions x 32 channels per pass (bank-conflict-free smem reads) and
//            emits 16/32-byte vector stores: the position-major fp32 copy, the masked rows of ref, and
//            the bf16 hi/lo split written DIRECTLY in the 128B-swizzled UMMA tile image the tcgen05
//            GEMM streams with bulk copies (no tensor map, no second transposition pass).
// HBM-bound: algorithmic bytes per image = 2*4NC read + (4NC + 2*2*2NC) write (+4MC masked rows).
#include "ipsr_common.cuh"

namespace ipsr {

constexpr int kPrepThreads = 256;
constexpr int kPrepPos = 32;

__global__ void __launch_bounds__(kPrepThreads)
prep_kernel(const float* __restrict__ x, const float* __restrict__ ref, int C, int N,
            const int* __restrict__ rank, int M,
            float* __restrict__ inv_norm, float* __restrict__ rnorm, float* __restrict__ xt,
            float* __restrict__ r_masked, uint8_t* __restrict__ x_tiles, uint8_t* __restrict__ r_tiles,
            int* __restrict__ nonfinite) {
  extern __shared__ float smem[];
  float* tile = smem;                      // [C][33]
  float* part = smem + (size_t)C * 33;     // [8][32]
  float* scale = part + 8 * 32;            // [32]

  const int is_ref = blockIdx.z;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kPrepPos;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* src = (is_ref ? ref : x) + (size_t)b * C * N;

  // ---- phase 1: load + sum of squares ----
  {
    const int p = p0 + lane;
    const bool ok = p < N;
    float ss = 0.f;
#pragma unroll 8
    for (int c = warp; c < C; c += 8) {
      float v = ok ? __ldg(src + (size_t)c * N + p) : 0.f;
      tile[c * 33 + lane] = v;
      ss = fmaf(v, v, ss);
    }
    part[warp * 32 + lane] = ss;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += part[w * 32 + threadIdx.x];
    const float nrm = sqrtf(tot);
    const int p = p0 + threadIdx.x;
    if (nonfinite && !(fabsf(tot) <= 3.4028234e38f)) atomicOr(nonfinite + b, 1);
    if (is_ref) {
      scale[threadIdx.x] = 1.0f;
      if (p < N && rnorm) rnorm[(size_t)b * N + p] = nrm;
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
  uint8_t* tiles = is_ref ? r_tiles : x_tiles;
  const int passes = 4 * ((C + 31) / 32);
  for (int pass = warp; pass < passes; pass += 8) {
    const int g = pass & 3, cg = pass >> 2;
    const int pl = g * 8 + pp;
    const int p = p0 + pl;
    const int c0 = cg * 32 + j * 8;
    if (p >= N || c0 >= C) continue;
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
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v0 = is_ref ? raw[2 * i] : __fmul_rn(raw[2 * i], sc);
        const float v1 = is_ref ? raw[2 * i + 1] : __fmul_rn(raw[2 * i + 1], sc);
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
        hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lo[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      const int kb = c0 / kTileK, chunk = (c0 % kTileK) / 8;
      const int rb = p / kTileRows, r = p % kTileRows;
      const uint32_t off = tile_chunk_offset(r, chunk);
      *reinterpret_cast<uint4*>(tiles + tile_offset_bytes(b, kb, 0, rb, KB, RB) + off) =
          make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(tiles + tile_offset_bytes(b, kb, 1, rb, KB, RB) + off) =
          make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

}  // namespace ipsr

extern "C" int ipsr_extract_normalize(const float* x, const float* ref, int B, int C, int N,
                                      const int32_t* rank_i32, int M,
                                      float* inv_norm, float* rnorm, float* xt, float* r_masked,
                                      void* x_tiles, void* r_tiles, int32_t* nonfinite, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(x && ref && inv_norm, IPSR_ERR_INVALID_ARG, "ipsr_extract_normalize: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0, IPSR_ERR_INVALID_ARG, "ipsr_extract_normalize: bad dims B=%d C=%d N=%d", B, C, N);
  IPSR_REQUIRE(C % 8 == 0, IPSR_ERR_UNSUPPORTED, "ipsr_extract_normalize: C=%d must be a multiple of 8", C);
  IPSR_REQUIRE(!(r_masked && M > 0) || rank_i32, IPSR_ERR_INVALID_ARG, "ipsr_extract_normalize: r_masked needs rank");
  if (x_tiles || r_tiles)
    IPSR_REQUIRE(C % kTileK == 0 && N % kTileRows == 0, IPSR_ERR_UNSUPPORTED,
                 "ipsr_extract_normalize: tile images need C %% 64 == 0 and N %% 128 == 0 (C=%d N=%d)", C, N);
  const size_t smem = ((size_t)C * 33 + 8 * 32 + 32) * sizeof(float);
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_extract_normalize: C=%d too large", C);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "prep smem attribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  dim3 grid((N + kPrepPos - 1) / kPrepPos, B, 2);
  prep_kernel<<<grid, kPrepThreads, smem, as_stream(stream)>>>(
      x, ref, C, N, rank_i32, M, inv_norm, rnorm, xt, (M > 0 ? r_masked : nullptr),
      reinterpret_cast<uint8_t*>(x_tiles), reinterpret_cast<uint8_t*>(r_tiles), nonfinite);
  return check_launch("ipsr_extract_normalize");
}
