// (e) backward of the shift operator, routed through the saved indices.
//
// Replaces models/IPSRFunction.py:144-178: N row-gathers out of an int64 [N,N,H,W] tensor, a
// zeroed [B,N,N] float matrix and a dense torch.mm(W^T [N,N], g [N,C]).  The reference keeps its
// attention in a LongTensor (:36,134), so W = trunc(A): a 0/1 matrix with one 1 per unmasked row
// (and for the first masked row), plus -- only for ill-conditioned inputs -- the few entries of the
// blended rows whose magnitude reaches 1.  gin = g + triple_w * W^T g is therefore a segment sum
// over "unit routes" plus a short exception list; no N x N object exists anywhere.
#include "ipsr_common.cuh"

namespace ipsr {

// ---------------------------------------------------------------------------------------------
// unit routes: stable counting sort of {q : unmasked or q == q_0} by p = ind[q]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
build_routes_kernel(const int* __restrict__ ind, const int* __restrict__ flag, const int* __restrict__ mask_idx,
                    int N, int M, int* __restrict__ route_ptr, int* __restrict__ route_q) {
  extern __shared__ int rsm[];
  int* cursor = rsm;            // [N+1] counts -> exclusive offsets -> running cursors
  int* key = rsm + (N + 1);     // [N]   p = ind[q] for routed q, -1 otherwise
  __shared__ int warp_tot[8];
  __shared__ int carry;
  const int b = blockIdx.x;
  const int* indb = ind + (size_t)b * N;
  const int q_first = (M > 0) ? mask_idx[0] : -1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int i = threadIdx.x; i <= N; i += blockDim.x) cursor[i] = 0;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int q = threadIdx.x; q < N; q += blockDim.x) {
    const bool routed = (flag[q] == 0) || (q == q_first);
    const int p = routed ? indb[q] : -1;
    key[q] = p;
    if (routed) atomicAdd(&cursor[p], 1);
  }
  __syncthreads();
  // exclusive scan of cursor[0..N) in chunks of 256, cursor[N] = total
  for (int base = 0; base < N; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < N) ? cursor[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int off = carry;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (i < N) cursor[i] = off + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += warp_tot[w];
      carry += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) cursor[N] = carry;
  __syncthreads();
  int* ptr = route_ptr + (size_t)b * (N + 1);
  for (int i = threadIdx.x; i <= N; i += blockDim.x) ptr[i] = cursor[i];
  __syncthreads();
  // stable fill by one warp: ascending q, duplicates inside a warp step ranked by lane
  if (warp == 0) {
    int* rq = route_q + (size_t)b * N;
    for (int base = 0; base < N; base += 32) {
      const int q = base + lane;
      const int p = (q < N) ? key[q] : -1;
      const bool active = p >= 0;
      const int mkey = active ? p : -1 - lane;              // inactive lanes never match anybody
      const unsigned peers = __match_any_sync(0xffffffffu, mkey);
      const int rnk = __popc(peers & ((1u << lane) - 1u));
      int start = 0;
      if (active && rnk == 0) {
        start = cursor[p];
        cursor[p] = start + __popc(peers);
      }
      start = __shfl_sync(0xffffffffu, start, __ffs(peers) - 1);
      if (active) rq[start + rnk] = q;
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// exceptions: thread per bank column replays row_l[p] = row_{l-1}[p]*wn_l (+ wo_l if p == p_l)
// ---------------------------------------------------------------------------------------------
constexpr int kExcChunk = 512;

template <bool WRITE>
__device__ __forceinline__ int replay_column(int p, int M, const int* __restrict__ ind_b, const int* __restrict__ mask_idx,
                                             const float* __restrict__ wn, const float* __restrict__ wo,
                                             float* s_wn, float* s_wo, int* s_p, int* __restrict__ out_l,
                                             float* __restrict__ out_w, bool valid) {
  float e = 0.f;
  int cnt = 0;
  for (int base = 0; base < M; base += kExcChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < kExcChunk && base + i < M; i += blockDim.x) {
      s_wn[i] = wn[base + i];
      s_wo[i] = wo[base + i];
      s_p[i] = ind_b[mask_idx[base + i]];
    }
    __syncthreads();
    if (!valid) continue;
    const int n = min(kExcChunk, M - base);
    for (int i = 0; i < n; ++i) {
      const int l = base + i;
      if (l == 0) {
        e = (s_p[0] == p) ? 1.f : 0.f;                     // in_attention[0, p_0] = 1     :100
        continue;
      }
      e = __fmul_rn(e, s_wn[i]);                            // row * wn                      :123
      if (s_p[i] == p) e = __fadd_rn(e, s_wo[i]);           // row[p_l] += wo                :124
      if (!(fabsf(e) < 1.0f)) {                             // survives the int64 store      :134
        if (WRITE) {
          out_l[cnt] = l;
          out_w[cnt] = trunc_as_reference(e);
        }
        ++cnt;
      }
    }
  }
  return cnt;
}

__global__ void __launch_bounds__(128)
build_exceptions_kernel(const int* __restrict__ ind, const int* __restrict__ mask_idx, const float* __restrict__ wn,
                        const float* __restrict__ wo, int N, int M, int* __restrict__ exc_start, int* __restrict__ exc_cnt,
                        int* __restrict__ exc_l, float* __restrict__ exc_w, int* __restrict__ exc_total, int exc_cap) {
  __shared__ float s_wn[kExcChunk], s_wo[kExcChunk];
  __shared__ int s_p[kExcChunk];
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = p < N;
  const int* ind_b = ind + (size_t)b * N;
  const float* wnb = wn + (size_t)b * M;
  const float* wob = wo + (size_t)b * M;
  const int cnt = replay_column<false>(p, M, ind_b, mask_idx, wnb, wob, s_wn, s_wo, s_p, nullptr, nullptr, valid);
  int start = 0;
  bool fits = false;
  if (valid && cnt > 0) {
    start = atomicAdd(exc_total + b, cnt);
    fits = (start + cnt <= exc_cap);
  }
  if (valid) {
    exc_start[(size_t)b * N + p] = fits ? start : 0;
    exc_cnt[(size_t)b * N + p] = fits ? cnt : 0;
  }
  // second pass only when somebody in the CTA has something to write (uniform decision)
  if (__syncthreads_or(fits ? 1 : 0)) {
    replay_column<true>(p, M, ind_b, mask_idx, wnb, wob, s_wn, s_wo, s_p,
                        exc_l + (size_t)b * exc_cap + start, exc_w + (size_t)b * exc_cap + start, valid && fits);
  }
}

// ---------------------------------------------------------------------------------------------
// backward proper
// ---------------------------------------------------------------------------------------------
// grid = (C / CT, B): the CTA stages CT rows of g[b] in shared memory; thread p sums its routes
// (ascending q: deterministic) and exceptions for the CT channels.  If the exception lists
// overflowed (exc_total > exc_cap, chaotic inputs only) the column replays the recurrence instead.
template <int CT>
__global__ void __launch_bounds__(256)
shift_bwd_kernel(const float* __restrict__ g, int C, int N, int M, const int* __restrict__ route_ptr,
                 const int* __restrict__ route_q, const int* __restrict__ exc_start, const int* __restrict__ exc_cnt,
                 const int* __restrict__ exc_l, const float* __restrict__ exc_w, const int* __restrict__ exc_total,
                 int exc_cap, const int* __restrict__ ind, const int* __restrict__ mask_idx,
                 const float* __restrict__ wn, const float* __restrict__ wo, float triple_w, float* __restrict__ gin) {
  extern __shared__ __align__(16) float grow[];           // [CT][N]
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * CT;
  const int ct = min(CT, C - c0);
  const float* gb = g + ((size_t)b * C + c0) * N;
  float* ob = gin + ((size_t)b * C + c0) * N;
  const int total = ct * N;
  if ((N & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(gb);
    float4* d4 = reinterpret_cast<float4*>(grow);
    for (int i = threadIdx.x; i < total / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) grow[i] = __ldg(gb + i);
  }
  __syncthreads();
  const int* ptr = route_ptr + (size_t)b * (N + 1);
  const int* rq = route_q + (size_t)b * N;
  const bool overflow = (M > 1) && exc_total && (exc_total[b] > exc_cap);
  for (int p = threadIdx.x; p < N; p += blockDim.x) {
    float acc[CT];
#pragma unroll
    for (int ch = 0; ch < CT; ++ch) acc[ch] = 0.f;
    const int r0 = ptr[p], r1 = ptr[p + 1];
    for (int r = r0; r < r1; ++r) {
      const int q = rq[r];
#pragma unroll
      for (int ch = 0; ch < CT; ++ch)
        if (ch < ct) acc[ch] += grow[ch * N + q];
    }
    if (M > 1 && !overflow && exc_cnt) {
      const int n = exc_cnt[(size_t)b * N + p];
      if (n > 0) {
        const int s = exc_start[(size_t)b * N + p];
        for (int e = 0; e < n; ++e) {
          const int q = mask_idx[exc_l[(size_t)b * exc_cap + s + e]];
          const float w = exc_w[(size_t)b * exc_cap + s + e];
#pragma unroll
          for (int ch = 0; ch < CT; ++ch)
            if (ch < ct) acc[ch] = fmaf(w, grow[ch * N + q], acc[ch]);
        }
      }
    } else if (overflow) {
      float e = (ind[(size_t)b * N + mask_idx[0]] == p) ? 1.f : 0.f;
      for (int l = 1; l < M; ++l) {
        const int ql = mask_idx[l];
        e = __fmul_rn(e, wn[(size_t)b * M + l]);
        if (ind[(size_t)b * N + ql] == p) e = __fadd_rn(e, wo[(size_t)b * M + l]);
        if (!(fabsf(e) < 1.0f)) {
          const float w = trunc_as_reference(e);
#pragma unroll
          for (int ch = 0; ch < CT; ++ch)
            if (ch < ct) acc[ch] = fmaf(w, grow[ch * N + ql], acc[ch]);
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < CT; ++ch)                          // g + weighted * triple_w           :173
      if (ch < ct) ob[(size_t)ch * N + p] = __fadd_rn(grow[ch * N + p], __fmul_rn(acc[ch], triple_w));
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
  build_exceptions_kernel<<<dim3((N + 127) / 128, B), 128, 0, as_stream(stream)>>>(ind, mask_idx, wn, wo, N, M, exc_start,
                                                                                  exc_cnt, exc_l, exc_w, exc_total, exc_cap);
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
  // channel rows per CTA: 8 while 8*N floats fit in ~64 KiB, else 4, 2, 1
  int CT = 8;
  while (CT > 1 && (size_t)CT * N * sizeof(float) > 64 * 1024) CT >>= 1;
  const size_t smem = (size_t)CT * N * sizeof(float);
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
  kern<<<dim3((C + CT - 1) / CT, B), 256, smem, as_stream(stream)>>>(g, C, N, M, route_ptr, route_q, exc_start, exc_cnt,
                                                                     exc_l, exc_w, exc_total, exc_cap, ind, mask_idx, wn,
                                                                     wo, triple_w, gin);
  return check_launch("ipsr_shift_bwd");
}
