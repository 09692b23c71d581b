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
// grid = (C / CT, B): the CTA stages CT rows of g[b] (C % CT == 0) in shared memory with ONE bulk async copy
// (the rows are contiguous in NCHW) while its threads already fetch the CSR entries of their columns.
// Shared memory holds nothing but the rows (+ a small queue), so several CTAs are resident per SM and the
// copy of one overlaps the gather of another.  Light bank columns (few routes / exceptions) are summed
// by their own thread in a loop of exactly their entry count; heavy ones -- a non-negative reference makes
// a few "hub" patches the best match of hundreds of positions -- are queued and summed by whole warps
// (lane-strided partial sums in ascending q, then a fixed xor tree: deterministic).  If the exception lists
// overflowed (exc_total > exc_cap, chaotic inputs only) the column replays the recurrence.
constexpr int kBwdLight = 12;
constexpr int kBwdQueue = 1024;
constexpr int kBwdPre = 4;               // columns per thread whose CSR entries are fetched ahead

template <int CT>
__global__ void __launch_bounds__(256)
shift_bwd_kernel(const float* __restrict__ g, int C, int N, int M, const int* __restrict__ route_ptr,
                 const int* __restrict__ route_q, const int* __restrict__ exc_start, const int* __restrict__ exc_cnt,
                 const int* __restrict__ exc_l, const float* __restrict__ exc_w, const int* __restrict__ exc_total,
                 int exc_cap, const int* __restrict__ ind, const int* __restrict__ mask_idx,
                 const float* __restrict__ wn, const float* __restrict__ wo, float triple_w, float* __restrict__ gin) {
  extern __shared__ __align__(128) float grow[];          // [CT][N]
  __shared__ int heavy[kBwdQueue];
  __shared__ int nheavy;
  __shared__ __align__(8) unsigned long long bar;
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * CT;
  const float* gb = g + ((size_t)b * C + c0) * N;
  float* ob = gin + ((size_t)b * C + c0) * N;
  const int total = CT * N;
  const bool bulk = ((total & 3) == 0) && ((reinterpret_cast<uintptr_t>(gb) & 15) == 0);
  if (threadIdx.x == 0) {
    nheavy = 0;
    if (bulk) {
      mbar_init(smem_u32(&bar), 1);
      mbar_fence_init();
      mbar_expect_tx(smem_u32(&bar), (uint32_t)total * 4u);
      bulk_g2s(smem_u32(grow), gb, (uint32_t)total * 4u, smem_u32(&bar));
    }
  }
  if (!bulk)
    for (int i = threadIdx.x; i < total; i += 256) grow[i] = __ldg(gb + i);
  const int* gptr = route_ptr + (size_t)b * (N + 1);
  const int* grq = route_q + (size_t)b * N;
  const bool has_exc = (M > 1) && exc_cnt && exc_total && (exc_total[b] != 0);
  const bool overflow = has_exc && (exc_total[b] > exc_cap);
  const bool lists = has_exc && !overflow;
  const int* ecnt = exc_cnt + (size_t)b * N;
  const int* estart = exc_start + (size_t)b * N;
  const int* el = exc_l + (size_t)b * exc_cap;
  const float* ew = exc_w + (size_t)b * exc_cap;
  __syncthreads();                                        // barrier initialised / plain copy complete
  bool landed = !bulk;

  auto replay = [&](int p, float (&acc)[CT]) {            // exception lists overflowed: rare, slow, bit-faithful
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
  };

  for (int base = 0; base < N; base += 256 * kBwdPre) {
    int r0[kBwdPre], r1[kBwdPre], ne[kBwdPre], es[kBwdPre], q0[kBwdPre];
#pragma unroll
    for (int i = 0; i < kBwdPre; ++i) {
      const int p = base + i * 256 + threadIdx.x;
      r0[i] = r1[i] = ne[i] = es[i] = 0;
      if (p < N) {
        r0[i] = __ldg(gptr + p);
        r1[i] = __ldg(gptr + p + 1);
        if (lists) {
          ne[i] = __ldg(ecnt + p);
          es[i] = __ldg(estart + p);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kBwdPre; ++i) q0[i] = (r1[i] > r0[i]) ? __ldg(grq + r0[i]) : 0;
    if (!landed) {
      mbar_wait(smem_u32(&bar), 0);
      landed = true;
    }
#pragma unroll
    for (int i = 0; i < kBwdPre; ++i) {
      const int p = base + i * 256 + threadIdx.x;
      if (p >= N) continue;
      const int n = r1[i] - r0[i];
      if (n + ne[i] > kBwdLight) {
        const int slot = atomicAdd(&nheavy, 1);            // queue order does not affect any sum
        if (slot < kBwdQueue) {
          heavy[slot] = p;
          continue;
        }
      }
      float acc[CT];
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) acc[ch] = 0.f;
      if (n > 0) {
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] += grow[ch * N + q0[i]];
        for (int r = r0[i] + 1; r < r1[i]; ++r) {
          const int q = __ldg(grq + r);
#pragma unroll
          for (int ch = 0; ch < CT; ++ch) acc[ch] += grow[ch * N + q];
        }
      }
      if (ne[i] > 0) {
        const int s = es[i];
        for (int e = 0; e < ne[i]; ++e) {
          const int q = __ldg(el + s + e);
          const float w = __ldg(ew + s + e);
#pragma unroll
          for (int ch = 0; ch < CT; ++ch) acc[ch] = fmaf(w, grow[ch * N + q], acc[ch]);
        }
      }
      if (overflow) replay(p, acc);
#pragma unroll
      for (int ch = 0; ch < CT; ++ch)                        // g + weighted * triple_w           :173
        ob[(size_t)ch * N + p] = __fadd_rn(grow[ch * N + p], __fmul_rn(acc[ch], triple_w));
    }
  }
  __syncthreads();

  // heavy columns: one warp each
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nh = min(nheavy, kBwdQueue);
  for (int h = warp; h < nh; h += 8) {
    const int p = heavy[h];
    const int r0 = __ldg(gptr + p), r1 = __ldg(gptr + p + 1);
    float acc[CT];
#pragma unroll
    for (int ch = 0; ch < CT; ++ch) acc[ch] = 0.f;
    for (int r = r0 + lane; r < r1; r += 32) {
      const int q = __ldg(grq + r);
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) acc[ch] += grow[ch * N + q];
    }
    if (lists) {
      const int ne = ecnt[p], s = estart[p];
      for (int e = lane; e < ne; e += 32) {
        const int q = __ldg(el + s + e);
        const float w = __ldg(ew + s + e);
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] = fmaf(w, grow[ch * N + q], acc[ch]);
      }
    }
    if (overflow && lane == 0) replay(p, acc);             // lane 0 replays (rare path)
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
  // channel rows per CTA: the largest of 8, 4, 2, 1 that divides C and keeps the rows within 64 KiB of shared memory
  // (>= 3 CTAs resident per SM: one CTA's bulk copy overlaps the gathers of the others)
  int CT = 8;
  while (CT > 1 && (C % CT != 0 || (size_t)CT * N * sizeof(float) > 64 * 1024)) CT >>= 1;
  const size_t smem = (size_t)CT * N * sizeof(float);
  IPSR_REQUIRE(smem <= 200 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_shift_bwd: N=%d too large", N);
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
