// (e) backward of the shift operator, routed through the saved indices.
//
// Replaces models/IPSRFunction.py:144-178: N row-gathers out of an int64 [N,N,H,W] tensor, a
// zeroed [B,N,N] float matrix and a dense torch.mm(W^T [N,N], g [N,C]).  The reference keeps its
// attention in a LongTensor (:36,134), so W = trunc(A): a 0/1 matrix with one 1 per unmasked row
// (and for the first masked row), plus -- only for ill-conditioned inputs -- the few entries of the
// blended rows whose magnitude reaches 1.  gin = g + triple_w * W^T g is therefore a segment sum
// over "unit routes" plus a short exception list; no N x N object exists anywhere.
#include <stdlib.h>

#include "ipsr_bookkeeping.cuh"

namespace ipsr {

__global__ void __launch_bounds__(256)
build_routes_kernel(const int* __restrict__ ind, const int* __restrict__ flag, const int* __restrict__ mask_idx,
                    int N, int M, int* __restrict__ route_ptr, int* __restrict__ route_q, int ms, const int* __restrict__ mcount) {
  extern __shared__ int rsm[];
  build_routes_cta(blockIdx.x, rsm, ind, flag, mask_idx, N, M, route_ptr, route_q, ms, mcount);
}

__global__ void __launch_bounds__(kExcThreads)
build_exceptions_kernel(const int* __restrict__ ind, const int* __restrict__ mask_idx, const float* __restrict__ wn,
                        const float* __restrict__ wo, int B, int N, int M, int* __restrict__ exc_start, int* __restrict__ exc_cnt,
                        int* __restrict__ exc_l, float* __restrict__ exc_w, int* __restrict__ exc_state, int exc_cap,
                        int ms, const int* __restrict__ mcount) {
  extern __shared__ __align__(16) int esm[];               // exc_smem_words(N, M)
  build_exceptions_cta(blockIdx.x, B, esm, ind, mask_idx, wn, wo, N, M, exc_start, exc_cnt, exc_l, exc_w, exc_state, exc_cap, ms,
                       mcount);
}

// ---------------------------------------------------------------------------------------------
// backward proper
// ---------------------------------------------------------------------------------------------
// grid = (parts, B): a CTA owns a contiguous range of channel tiles (CT rows each, C % CT == 0) of ONE image and
// streams them through a two-stage ring of bulk async copies (the CT rows of a tile are contiguous in NCHW), so
// the copy of tile t+1 overlaps the work on tile t.  Most bank columns receive nothing (a non-negative reference
// concentrates the matches on a few hundred patches), hence per tile:
//   (1) copy-out   gin tile = g tile, 16-byte vector stores straight from shared memory;
//   (2) correct    the columns that do receive something -- a compact list built ONCE per CTA from the CSR and
//                  the exception directory, so every lane has work -- are recomputed as
//                  g[:,p] + triple_w (sum of routed rows + weighted exception rows) by one thread each, hub
//                  columns (hundreds of routes) by whole warps (lane-strided partial sums in ascending q, then a
//                  fixed xor tree): deterministic.
// If the exception lists of an image are unusable (exc_total[b] >= kExcReplay: the batch's pool is exhausted or a blend
// weight is non-finite, chaotic inputs only) every column of that image replays the recurrence.
constexpr int kBwdLight = 24;            // entries a single thread sums; heavier columns go to a warp
constexpr int kBwdQueue = 512;

template <int CT>
__global__ void __launch_bounds__(CT >= 8 ? 512 : 1024, CT >= 8 ? 2 : 1)
shift_bwd_kernel(const float* __restrict__ g, int C, int N, int M, int tiles_per_cta, const int* __restrict__ route_ptr,
                 const int* __restrict__ route_q, const int* __restrict__ exc_start, const int* __restrict__ exc_cnt,
                 const int* __restrict__ exc_l, const float* __restrict__ exc_w, const int* __restrict__ exc_total,
                 int exc_cap, const int* __restrict__ ind, const int* __restrict__ mask_idx,
                 const float* __restrict__ wn, const float* __restrict__ wo, float triple_w, float* __restrict__ gin,
                 int ninfo, int nexc_s, int ms, const int* __restrict__ mcount) {
  extern __shared__ __align__(128) float bwd_smem[];      // rows[2][CT*N] | spec[N] | info[N] int4 | rq[N] | el[E] | ew[E]
  __shared__ int heavy[kBwdQueue];
  __shared__ int4 heavy_info[kBwdQueue];
  __shared__ int nheavy, nspec_s;
  __shared__ __align__(8) unsigned long long bars[2];
  const int b = blockIdx.y;
  // per-image masks: mask_idx is [B][ms], mcount[b] steps; M stays the row stride of wn / wo
  const int Mc = mcount ? mcount[b] : M;
  if (mask_idx) mask_idx += (size_t)b * ms;
  const int ntiles = C / CT;
  const int t0 = blockIdx.x * tiles_per_cta;
  const int t1 = min(ntiles, t0 + tiles_per_cta);
  if (t0 >= t1) return;
  const int tile_elems = CT * N;
  const uint32_t tile_bytes = (uint32_t)tile_elems * 4u;
  float* rows0 = bwd_smem;
  int* spec = reinterpret_cast<int*>(bwd_smem + 2 * (size_t)tile_elems);
  int4* info = reinterpret_cast<int4*>(spec + ((N + 3) & ~3));      // per listed column: first route, routes, first exception, exceptions
  int* rq_s = reinterpret_cast<int*>(info + ninfo);                 // the CSR's row list
  int* el_s = rq_s + N;                                             // exception entries (when they fit)
  float* ew_s = reinterpret_cast<float*>(el_s + nexc_s);
  const int nthreads = blockDim.x;
  const float* gimg = g + (size_t)b * C * N;
  float* oimg = gin + (size_t)b * C * N;
  const bool vec = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(gimg) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(oimg) & 15) == 0);

  auto load_tile = [&](int t, int buf) {                  // executed by thread 0 (bulk) or by everybody (fallback)
    const float* src = gimg + (size_t)t * tile_elems;
    float* dst = rows0 + (size_t)buf * tile_elems;
    if (vec) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(&bars[buf]), tile_bytes);
        bulk_g2s(smem_u32(dst), src, tile_bytes, smem_u32(&bars[buf]));
      }
    } else {
      for (int i = threadIdx.x; i < tile_elems; i += nthreads) dst[i] = __ldg(src + i);
    }
  };

  if (threadIdx.x == 0) {
    nheavy = 0;
    nspec_s = 0;
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_fence_init();
  }
  __syncthreads();
  load_tile(t0, 0);
  if (t0 + 1 < t1) load_tile(t0 + 1, 1);

  // exc_total is the exception state of ipsr_build_exceptions: [b] entries of image b (>= kExcReplay: lists unusable,
  // replay), [B + b] first entry of image b in the pool shared by the batch
  const int ecount = ((M > 1) && exc_cnt && exc_total) ? exc_total[b] : 0;
  const bool has_exc = ecount != 0;
  const bool overflow = ecount >= kExcReplay;
  const bool lists = has_exc && !overflow;
  const size_t ebase = lists ? (size_t)exc_total[gridDim.y + b] : 0;
  const int* gptr = route_ptr + (size_t)b * (N + 1);
  const int* grq = route_q + (size_t)b * N;
  const int* ecnt = exc_cnt + (size_t)b * N;
  const int* estart = exc_start + (size_t)b * N;
  const int* el = exc_l + ebase;
  const float* ew = exc_w + ebase;
  // the columns that receive something, once per CTA (list order does not affect any sum); their CSR entries and
  // the image's route / exception lists are staged in shared memory, so that the per-tile work below never waits
  // on global memory for an index
  const int etotal = lists ? ecount : 0;
  const bool exc_in_smem = etotal <= nexc_s;
  for (int p = threadIdx.x; p < N; p += nthreads) {
    const int r0 = __ldg(gptr + p), n = __ldg(gptr + p + 1) - r0;
    const int ne = lists ? __ldg(ecnt + p) : 0;
    const int work = n + ne;
    if (overflow || work > 0) {
      const int4 rec = make_int4(r0, n, lists ? __ldg(estart + p) : 0, ne);
      bool queued = false;
      if (work > kBwdLight && !overflow) {
        const int slot = atomicAdd(&nheavy, 1);
        if (slot < kBwdQueue) {
          heavy[slot] = p;
          heavy_info[slot] = rec;
          queued = true;
        }
      }
      if (!queued) {
        const int k = atomicAdd(&nspec_s, 1);
        spec[k] = p;
        if (k < ninfo) info[k] = rec;
      }
    }
  }
  for (int i = threadIdx.x; i < N; i += nthreads) rq_s[i] = __ldg(grq + i);
  if (exc_in_smem)
    for (int i = threadIdx.x; i < etotal; i += nthreads) {
      el_s[i] = __ldg(el + i);
      ew_s[i] = __ldg(ew + i);
    }
  __syncthreads();
  const int* elx = exc_in_smem ? el_s : el;
  const float* ewx = exc_in_smem ? ew_s : ew;
  const int nspec = nspec_s;
  const int nh = min(nheavy, kBwdQueue);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = nthreads >> 5;

  for (int t = t0; t < t1; ++t) {
    const int buf = (t - t0) & 1;
    const float* grow = rows0 + (size_t)buf * tile_elems;
    float* ob = oimg + (size_t)t * tile_elems;
    if (vec) mbar_wait(smem_u32(&bars[buf]), (uint32_t)((t - t0) >> 1) & 1u);
    else __syncthreads();

    // (1) copy-out: g + triple_w * 0
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(grow);
      float4* d4 = reinterpret_cast<float4*>(ob);
      for (int i = threadIdx.x; i < tile_elems / 4; i += nthreads) d4[i] = s4[i];
    } else {
      for (int i = threadIdx.x; i < tile_elems; i += nthreads) ob[i] = grow[i];
    }
    __syncthreads();                                         // the corrections below overwrite some of these stores

    // (2) corrections
    for (int k = threadIdx.x; k < nspec; k += nthreads) {
      const int p = spec[k];
      int4 rec;
      if (k < ninfo) rec = info[k];
      else rec = make_int4(__ldg(gptr + p), __ldg(gptr + p + 1) - __ldg(gptr + p), lists ? __ldg(estart + p) : 0,
                           lists ? __ldg(ecnt + p) : 0);
      const int r0 = rec.x, r1 = rec.x + rec.y, es = rec.z, ne = rec.w;
      float acc[CT];
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) acc[ch] = 0.f;
      for (int r = r0; r < r1; ++r) {
        const int q = rq_s[r];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] += grow[ch * N + q];
      }
      for (int e = 0; e < ne; ++e) {
        const int q = elx[es + e];
        const float w = ewx[es + e];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] = fmaf(w, grow[ch * N + q], acc[ch]);
      }
      if (overflow) {                                        // rare, slow, bit-faithful replay of the recurrence
        float e = (ind[(size_t)b * N + mask_idx[0]] == p) ? 1.f : 0.f;
        for (int l = 1; l < Mc; ++l) {
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
      for (int ch = 0; ch < CT; ++ch)                        // g + weighted * triple_w           :173
        ob[(size_t)ch * N + p] = __fadd_rn(grow[ch * N + p], __fmul_rn(acc[ch], triple_w));
    }
    for (int h = warp; h < nh; h += nwarps) {                // hub columns: one warp each
      const int p = heavy[h];
      const int4 rec = heavy_info[h];
      const int r0 = rec.x, r1 = rec.x + rec.y;
      float acc[CT];
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) acc[ch] = 0.f;
      for (int r = r0 + lane; r < r1; r += 32) {
        const int q = rq_s[r];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] += grow[ch * N + q];
      }
      if (lists) {
        const int ne = rec.w, es = rec.z;
        for (int e = lane; e < ne; e += 32) {
          const int q = elx[es + e];
          const float w = ewx[es + e];
#pragma unroll
          for (int ch = 0; ch < CT; ++ch) acc[ch] = fmaf(w, grow[ch * N + q], acc[ch]);
        }
      }
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) acc[ch] = warp_sum(acc[ch]);
      if (lane == 0) {
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) ob[(size_t)ch * N + p] = __fadd_rn(grow[ch * N + p], __fmul_rn(acc[ch], triple_w));
      }
    }
    __syncthreads();                                         // everybody is done reading this buffer
    if (t + 2 < t1) load_tile(t + 2, buf);
  }
}

}  // namespace ipsr

extern "C" int ipsr_build_routes(const int32_t* ind, const int32_t* flag, const int32_t* mask_idx,
                                 int B, int N, int M, int32_t* route_ptr, int32_t* route_q, void* stream) {
  return ipsr::build_routes_ex(ind, flag, mask_idx, B, N, M, route_ptr, route_q, stream, 0, nullptr);
}

int ipsr::build_routes_ex(const int32_t* ind, const int32_t* flag, const int32_t* mask_idx, int B, int N, int M,
                          int32_t* route_ptr, int32_t* route_q, void* stream, int ms, const int32_t* mcount) {
  using namespace ipsr;
  IPSR_REQUIRE(ind && flag && route_ptr && route_q && (M == 0 || mask_idx), IPSR_ERR_INVALID_ARG, "ipsr_build_routes: null pointer");
  IPSR_REQUIRE(B > 0 && N > 0, IPSR_ERR_INVALID_ARG, "ipsr_build_routes: bad dims");
  IPSR_REQUIRE(N <= 16384, IPSR_ERR_UNSUPPORTED, "ipsr_build_routes: N=%d > 16384", N);
  const size_t smem = (size_t)(2 * N + 1) * sizeof(int);
  if (smem + 2048 > 48 * 1024) {   // dynamic + static shared memory above the default limit (set per call: the attribute is per device)
    cudaError_t e = cudaFuncSetAttribute(build_routes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "build_routes smem attribute: %s", cudaGetErrorString(e));
  }
  build_routes_kernel<<<B, 256, smem, as_stream(stream)>>>(ind, flag, mask_idx, N, M, route_ptr, route_q, ms, mcount);
  return check_launch("ipsr_build_routes");
}

extern "C" int ipsr_build_exceptions(const int32_t* ind, const int32_t* mask_idx, const float* wn, const float* wo,
                                     int B, int N, int M, int32_t* exc_start, int32_t* exc_cnt,
                                     int32_t* exc_l, float* exc_w, int32_t* exc_total, int exc_cap, void* stream) {
  return ipsr::build_exceptions_ex(ind, mask_idx, wn, wo, B, N, M, exc_start, exc_cnt, exc_l, exc_w, exc_total, exc_cap, stream, 0,
                                   nullptr);
}

int ipsr::build_exceptions_ex(const int32_t* ind, const int32_t* mask_idx, const float* wn, const float* wo, int B, int N, int M,
                              int32_t* exc_start, int32_t* exc_cnt, int32_t* exc_l, float* exc_w, int32_t* exc_state,
                              int exc_cap, void* stream, int ms, const int32_t* mcount) {
  using namespace ipsr;
  if (M <= 1) return IPSR_OK;                               // rows l >= 1 do not exist
  IPSR_REQUIRE(ind && mask_idx && wn && wo && exc_start && exc_cnt && exc_l && exc_w && exc_state, IPSR_ERR_INVALID_ARG,
               "ipsr_build_exceptions: null pointer");
  IPSR_REQUIRE(B > 0 && N > 0 && exc_cap > 0 && B <= 65535, IPSR_ERR_INVALID_ARG, "ipsr_build_exceptions: bad dims");
  const size_t smem = exc_smem_words(N, M) * sizeof(int);
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_build_exceptions: N=%d / M=%d too large", N, M);
  if (smem + 1024 > 48 * 1024) {                            // + the kernel's static shared memory
    cudaError_t e = cudaFuncSetAttribute(build_exceptions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "build_exceptions smem attribute: %s", cudaGetErrorString(e));
  }
  // one owner thread per distinct matched column: as many threads as masked steps (whole warps, 128..1024)
  int threads = (M + 31) & ~31;
  if (threads < 128) threads = 128;
  if (threads > kExcThreads) threads = kExcThreads;
  build_exceptions_kernel<<<B, threads, smem, as_stream(stream)>>>(ind, mask_idx, wn, wo, B, N, M, exc_start, exc_cnt, exc_l, exc_w,
                                                                  exc_state, exc_cap, ms, mcount);
  return check_launch("ipsr_build_exceptions");
}

extern "C" int ipsr_shift_bwd(const float* g, int B, int C, int N, int M,
                              const int32_t* route_ptr, const int32_t* route_q,
                              const int32_t* exc_start, const int32_t* exc_cnt, const int32_t* exc_l, const float* exc_w,
                              const int32_t* exc_total, int exc_cap,
                              const int32_t* ind, const int32_t* mask_idx, const float* wn, const float* wo,
                              float triple_w, float* gin, void* stream) {
  return ipsr_shift_bwd_masks(g, B, C, N, M, route_ptr, route_q, exc_start, exc_cnt, exc_l, exc_w, exc_total, exc_cap, ind, mask_idx,
                              wn, wo, triple_w, gin, 0, nullptr, stream);
}

extern "C" int ipsr_shift_bwd_masks(const float* g, int B, int C, int N, int M,
                                    const int32_t* route_ptr, const int32_t* route_q,
                                    const int32_t* exc_start, const int32_t* exc_cnt, const int32_t* exc_l, const float* exc_w,
                                    const int32_t* exc_total, int exc_cap,
                                    const int32_t* ind, const int32_t* mask_idx, const float* wn, const float* wo,
                                    float triple_w, float* gin, int mask_stride, const int32_t* m_count, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(mask_stride == 0 || (mask_stride == N && m_count), IPSR_ERR_INVALID_ARG,
               "ipsr_shift_bwd_masks: per-image masks need mask_stride == N and m_count");
  IPSR_REQUIRE(g && gin && route_ptr && route_q, IPSR_ERR_INVALID_ARG, "ipsr_shift_bwd: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0 && B <= 65535, IPSR_ERR_INVALID_ARG, "ipsr_shift_bwd: bad dims");
  if (M > 1)
    IPSR_REQUIRE(exc_start && exc_cnt && exc_l && exc_w && exc_total && ind && mask_idx && wn && wo, IPSR_ERR_INVALID_ARG,
                 "ipsr_shift_bwd: exception lists / replay operands missing");
  // channel rows per tile: the largest of 8, 4, 2, 1 that divides C and keeps a tile within 32 KiB (N > 2048: 64 KiB)
  int CT = 8;
  const size_t tile_cap = (N <= 2048 ? 32 : 64) * 1024;
  while (CT > 1 && (C % CT != 0 || (size_t)CT * N * sizeof(float) > tile_cap)) CT >>= 1;
  // shared memory: two tiles + the column list [N] + the CSR row list [N], then as much of the per-column records
  // (16 B each) and of the exception entries (8 B each) as fits: what does not fit is read from global memory
  const size_t base_smem = 2 * (size_t)CT * N * sizeof(float) + 2 * (size_t)((N + 3) & ~3) * sizeof(int);
  IPSR_REQUIRE(base_smem <= 227 * 1024 - 12 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_shift_bwd: N=%d too large", N);
  const size_t room = (N <= 2048 ? 96 : 215) * 1024 - 12 * 1024 > base_smem ? (N <= 2048 ? 96 : 215) * 1024 - 12 * 1024 - base_smem : 0;
  int ninfo = (int)((room / 2) / 16);
  if (ninfo > N) ninfo = N;
  int nexc_s = (int)((room - (size_t)ninfo * 16) / 8) & ~3;
  if (nexc_s > exc_cap) nexc_s = (exc_cap + 3) & ~3;
  if (M <= 1) nexc_s = 0;
  const size_t smem = base_smem + (size_t)ninfo * 16 + (size_t)nexc_s * 8;
  const int threads = N > 2048 ? 1024 : 512;
  // Tiles per CTA: every CTA pays a fixed set-up (staging the image's index lists) and the per-image work is uneven
  // (hub columns, long exception runs), so the grid is cut into about 1.75 CTAs per SM -- measured best on the B200 for
  // 32x32 (B = 16) and 64x64 (B = 64) maps alike -- with at least two tiles per CTA to keep the two-stage ring busy.
  const int ntiles = C / CT;
  int tiles_per_cta = (int)(((long long)B * ntiles + 258) / 259);
  if (tiles_per_cta < 2) tiles_per_cta = 2;
  if (tiles_per_cta > ntiles) tiles_per_cta = ntiles;
  const int parts = (ntiles + tiles_per_cta - 1) / tiles_per_cta;
  void (*kern)(const float*, int, int, int, int, const int*, const int*, const int*, const int*, const int*, const float*,
               const int*, int, const int*, const int*, const float*, const float*, float, float*, int, int, int, const int*) = nullptr;
  switch (CT) {
    case 8: kern = shift_bwd_kernel<8>; break;
    case 4: kern = shift_bwd_kernel<4>; break;
    case 2: kern = shift_bwd_kernel<2>; break;
    default: kern = shift_bwd_kernel<1>; break;
  }
  if (smem + 12 * 1024 > 48 * 1024) {                       // static (queues) + dynamic shared memory above the default limit
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "shift_bwd smem attribute: %s", cudaGetErrorString(e));
  }
  kern<<<dim3(parts, B), threads, smem, as_stream(stream)>>>(g, C, N, M, tiles_per_cta, route_ptr, route_q, exc_start, exc_cnt,
                                                            exc_l, exc_w, exc_total, exc_cap, ind, mask_idx, wn, wo, triple_w,
                                                            gin, ninfo, nexc_s, mask_stride, m_count);
  return check_launch("ipsr_shift_bwd");
}
