// (e) backward of the shift operator, routed through the saved indices.
//
// Replaces models/IPSRFunction.py:144-178: N row-gathers out of an int64 [N,N,H,W] tensor, a
// zeroed [B,N,N] float matrix and a dense torch.mm(W^T [N,N], g [N,C]).  The reference keeps its
// attention in a LongTensor (:36,134), so W = trunc(A): a 0/1 matrix with one 1 per unmasked row
// (and for the first masked row), plus -- only for ill-conditioned inputs -- the few entries of the
// blended rows whose magnitude reaches 1.  gin = g + triple_w * W^T g is therefore a segment sum
// over "unit routes" plus a short exception list; no N x N object exists anywhere.
#include "ipsr_bookkeeping.cuh"

namespace ipsr {

__global__ void __launch_bounds__(256)
build_routes_kernel(const int* __restrict__ ind, const int* __restrict__ flag, const int* __restrict__ mask_idx,
                    int N, int M, int* __restrict__ route_ptr, int* __restrict__ route_q) {
  extern __shared__ int rsm[];
  build_routes_cta(blockIdx.x, rsm, ind, flag, mask_idx, N, M, route_ptr, route_q);
}

__global__ void __launch_bounds__(kExcThreads)
build_exceptions_kernel(const int* __restrict__ ind, const int* __restrict__ mask_idx, const float* __restrict__ wn,
                        const float* __restrict__ wo, int N, int M, int* __restrict__ exc_start, int* __restrict__ exc_cnt,
                        int* __restrict__ exc_l, float* __restrict__ exc_w, int* __restrict__ exc_total, int exc_cap,
                        int nparts) {
  extern __shared__ int esm[];                             // [N] + 3*kExcChunk words
  build_exceptions_cta(blockIdx.x / nparts, blockIdx.x % nparts, nparts, esm, ind, mask_idx, wn, wo, N, M, exc_start,
                       exc_cnt, exc_l, exc_w, exc_total, exc_cap);
}

// ---------------------------------------------------------------------------------------------
// backward proper
// ---------------------------------------------------------------------------------------------
// grid = (C / CT, B): the CTA stages CT rows of g[b] (C % CT == 0) and the CSR of the image in shared
// memory.  Light bank columns (few routes) are summed by their own thread, in a loop of exactly their
// route count; heavy ones -- a non-negative reference makes a few "hub" patches the best match of
// hundreds of positions -- are queued and summed by whole warps (lane-strided partial sums in ascending q,
// then a fixed xor tree: deterministic).  If the exception lists overflowed (exc_total > exc_cap, chaotic
// inputs only) the column replays the recurrence.
constexpr int kBwdLight = 12;

template <int CT>
__global__ void __launch_bounds__(256)
shift_bwd_kernel(const float* __restrict__ g, int C, int N, int M, const int* __restrict__ route_ptr,
                 const int* __restrict__ route_q, const int* __restrict__ exc_start, const int* __restrict__ exc_cnt,
                 const int* __restrict__ exc_l, const float* __restrict__ exc_w, const int* __restrict__ exc_total,
                 int exc_cap, const int* __restrict__ ind, const int* __restrict__ mask_idx,
                 const float* __restrict__ wn, const float* __restrict__ wo, float triple_w, float* __restrict__ gin) {
  extern __shared__ __align__(16) float grow[];           // [CT][N], heavy-column queue [N], CSR copy [2N+1]
  __shared__ int nheavy;
  int* heavy = reinterpret_cast<int*>(grow + (size_t)CT * N);   // [N] (every column can be heavy when exceptions abound)
  int* ptr = heavy + N;                                          // [N+1]
  int* rq = ptr + (N + 1);                                       // [N]
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * CT;
  const float* gb = g + ((size_t)b * C + c0) * N;
  float* ob = gin + ((size_t)b * C + c0) * N;
  const int total = CT * N;
  if (threadIdx.x == 0) nheavy = 0;
  if ((N & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(gb);
    float4* d4 = reinterpret_cast<float4*>(grow);
#pragma unroll 4
    for (int i = threadIdx.x; i < total / 4; i += 256) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = threadIdx.x; i < total; i += 256) grow[i] = __ldg(gb + i);
  }
  {
    // the index lists are shared by all channels: one coalesced copy replaces dependent global loads
    const int* gptr = route_ptr + (size_t)b * (N + 1);
    const int* grq = route_q + (size_t)b * N;
    for (int i = threadIdx.x; i <= N; i += 256) ptr[i] = __ldg(gptr + i);
    for (int i = threadIdx.x; i < N; i += 256) rq[i] = __ldg(grq + i);
  }
  const bool has_exc = (M > 1) && exc_cnt && exc_total && (exc_total[b] != 0);
  const bool overflow = has_exc && (exc_total[b] > exc_cap);
  const bool lists = has_exc && !overflow;
  const int* ecnt = exc_cnt + (size_t)b * N;
  const int* estart = exc_start + (size_t)b * N;
  const int* el = exc_l + (size_t)b * exc_cap;
  const float* ew = exc_w + (size_t)b * exc_cap;
  __syncthreads();

  for (int p = threadIdx.x; p < N; p += 256) {
    const int r0 = ptr[p], r1 = ptr[p + 1];
    const int ne = lists ? ecnt[p] : 0;
    if ((r1 - r0) + ne > kBwdLight) {
      heavy[atomicAdd(&nheavy, 1)] = p;                    // queue order does not affect any sum
      continue;
    }
    float acc[CT];
#pragma unroll
    for (int ch = 0; ch < CT; ++ch) acc[ch] = 0.f;
    for (int r = r0; r < r1; ++r) {
      const int q = rq[r];
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) acc[ch] += grow[ch * N + q];
    }
    if (ne > 0) {
      const int s = estart[p];
      for (int e = 0; e < ne; ++e) {
        const int q = mask_idx[el[s + e]];
        const float w = ew[s + e];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] = fmaf(w, grow[ch * N + q], acc[ch]);
      }
    }
    if (overflow) {
      float e = (ind[(size_t)b * N + mask_idx[0]] == p) ? 1.f : 0.f;
      for (int l = 1; l < M; ++l) {
        const int ql = mask_idx[l];
        e = __fmul_rn(e, wn[(size_t)b * M + l]);
        if (ind[(size_t)b * N + ql] == p) e = __fadd_rn(e, wo[(size_t)b * M + l]);
        if (!(fabsf(e) < 1.0f)) {
          const float w = trunc_as_reference(e);
#pragma unroll
          for (int ch = 0; ch < CT; ++ch) acc[ch] = fmaf(w, grow[ch * N + ql], acc[ch]);
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < CT; ++ch)                          // g + weighted * triple_w           :173
      ob[(size_t)ch * N + p] = __fadd_rn(grow[ch * N + p], __fmul_rn(acc[ch], triple_w));
  }
  __syncthreads();

  // heavy columns: one warp each
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nh = nheavy;
  for (int h = warp; h < nh; h += 8) {
    const int p = heavy[h];
    const int r0 = ptr[p], r1 = ptr[p + 1];
    float acc[CT];
#pragma unroll
    for (int ch = 0; ch < CT; ++ch) acc[ch] = 0.f;
    for (int r = r0 + lane; r < r1; r += 32) {
      const int q = rq[r];
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) acc[ch] += grow[ch * N + q];
    }
    if (lists) {
      const int ne = ecnt[p], s = estart[p];
      for (int e = lane; e < ne; e += 32) {
        const int q = mask_idx[el[s + e]];
        const float w = ew[s + e];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] = fmaf(w, grow[ch * N + q], acc[ch]);
      }
    }
    if (overflow) {                                          // lane 0 replays (rare path)
      if (lane == 0) {
        float e = (ind[(size_t)b * N + mask_idx[0]] == p) ? 1.f : 0.f;
        for (int l = 1; l < M; ++l) {
          const int ql = mask_idx[l];
          e = __fmul_rn(e, wn[(size_t)b * M + l]);
          if (ind[(size_t)b * N + ql] == p) e = __fadd_rn(e, wo[(size_t)b * M + l]);
          if (!(fabsf(e) < 1.0f)) {
            const float w = trunc_as_reference(e);
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) acc[ch] = fmaf(w, grow[ch * N + ql], acc[ch]);
          }
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < CT; ++ch) acc[ch] = warp_sum(acc[ch]);
    if (lane == 0) {
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) ob[(size_t)ch * N + p] = __fadd_rn(grow[ch * N + p], __fmul_rn(acc[ch], triple_w));
    }
  }
}

}  // namespace ipsr

extern "C" int ipsr_build_routes(const int32_t* ind, const int32_t* flag, const int32_t* mask_idx,
                                 int B, int N, int M, int32_t* route_ptr, int32_t* route_q, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(ind && flag && route_ptr && route_q && (M == 0 || mask_idx), IPSR_ERR_INVALID_ARG, "ipsr_build_routes: null pointer");
  IPSR_REQUIRE(B > 0 && N > 0, IPSR_ERR_INVALID_ARG, "ipsr_build_routes: bad dims");
  IPSR_REQUIRE(N <= 16384, IPSR_ERR_UNSUPPORTED, "ipsr_build_routes: N=%d > 16384", N);
  const size_t smem = (size_t)(2 * N + 1) * sizeof(int);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(build_routes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "build_routes smem attribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  build_routes_kernel<<<B, 256, smem, as_stream(stream)>>>(ind, flag, mask_idx, N, M, route_ptr, route_q);
  return check_launch("ipsr_build_routes");
}

extern "C" int ipsr_build_exceptions(const int32_t* ind, const int32_t* mask_idx, const float* wn, const float* wo,
                                     int B, int N, int M, int32_t* exc_start, int32_t* exc_cnt,
                                     int32_t* exc_l, float* exc_w, int32_t* exc_total, int exc_cap, void* stream) {
  using namespace ipsr;
  if (M <= 1) return IPSR_OK;                               // rows l >= 1 do not exist
  IPSR_REQUIRE(ind && mask_idx && wn && wo && exc_start && exc_cnt && exc_l && exc_w && exc_total, IPSR_ERR_INVALID_ARG,
               "ipsr_build_exceptions: null pointer");
  IPSR_REQUIRE(B > 0 && N > 0 && exc_cap > 0 && B <= 65535, IPSR_ERR_INVALID_ARG, "ipsr_build_exceptions: bad dims");
  const size_t smem = ((size_t)N + 3 * kExcChunk) * sizeof(int);
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_build_exceptions: N=%d too large", N);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(build_exceptions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "build_exceptions smem attribute: %s", cudaGetErrorString(e));
  }
  const int nparts = exc_parts(M);
  build_exceptions_kernel<<<B * nparts, kExcThreads, smem, as_stream(stream)>>>(ind, mask_idx, wn, wo, N, M, exc_start, exc_cnt,
                                                                               exc_l, exc_w, exc_total, exc_cap, nparts);
  return check_launch("ipsr_build_exceptions");
}

extern "C" int ipsr_shift_bwd(const float* g, int B, int C, int N, int M,
                              const int32_t* route_ptr, const int32_t* route_q,
                              const int32_t* exc_start, const int32_t* exc_cnt, const int32_t* exc_l, const float* exc_w,
                              const int32_t* exc_total, int exc_cap,
                              const int32_t* ind, const int32_t* mask_idx, const float* wn, const float* wo,
                              float triple_w, float* gin, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(g && gin && route_ptr && route_q, IPSR_ERR_INVALID_ARG, "ipsr_shift_bwd: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0 && B <= 65535, IPSR_ERR_INVALID_ARG, "ipsr_shift_bwd: bad dims");
  if (M > 1)
    IPSR_REQUIRE(exc_start && exc_cnt && exc_l && exc_w && exc_total && ind && mask_idx && wn && wo, IPSR_ERR_INVALID_ARG,
                 "ipsr_shift_bwd: exception lists / replay operands missing");
  // channel rows per CTA: the largest of 8, 4, 2, 1 that divides C and keeps the CTA near 112 KiB of shared memory
  int CT = 8;
  while (CT > 1 && (C % CT != 0 || (size_t)(CT + 3) * N * sizeof(float) > 112 * 1024)) CT >>= 1;
  const size_t smem = (size_t)CT * N * sizeof(float) + (size_t)(3 * N + 2) * sizeof(int);
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_shift_bwd: N=%d too large", N);
  void (*kern)(const float*, int, int, int, const int*, const int*, const int*, const int*, const int*, const float*,
               const int*, int, const int*, const int*, const float*, const float*, float, float*) = nullptr;
  switch (CT) {
    case 8: kern = shift_bwd_kernel<8>; break;
    case 4: kern = shift_bwd_kernel<4>; break;
    case 2: kern = shift_bwd_kernel<2>; break;
    default: kern = shift_bwd_kernel<1>; break;
  }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "shift_bwd smem attribute: %s", cudaGetErrorString(e));
  }
  kern<<<dim3(C / CT, B), 256, smem, as_stream(stream)>>>(g, C, N, M, route_ptr, route_q, exc_start, exc_cnt,
                                                                     exc_l, exc_w, exc_total, exc_cap, ind, mask_idx, wn,
                                                                     wo, triple_w, gin);
  return check_launch("ipsr_shift_bwd");
}
