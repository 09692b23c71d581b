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
// gin = g + triple_w * W^T g is a COPY of g plus corrections in the few hundred bank columns that receive anything
// (a non-negative reference concentrates the matches on a few hundred patches).  The kernel therefore streams:
//   * g is cut into tiles of CT channel rows (CT * N contiguous floats in NCHW; image-major, so the whole tensor is
//     one sequence of tiles) and every persistent CTA owns a contiguous run of them;
//   * a tile arrives by ONE bulk async copy (TMA engine) into a ring of shared-memory stages, is corrected IN PLACE
//     and leaves by ONE bulk async store -- no thread ever copies a byte, loads and stores of neighbouring tiles
//     overlap the corrections;
//   * corrections: per image (once per CTA and image) the columns that receive something are listed with their CSR
//     / exception ranges in shared memory; per tile, phase A sums the routed rows of every (column, channel) pair
//     into a delta buffer -- one thread per pair for short columns, one warp per pair for hub columns (lane-strided
//     partial sums in ascending q, then a fixed xor tree): deterministic -- and phase B adds the deltas into the tile
//     (a column can also be a SOURCE of another column, hence two phases).
// If the exception lists of an image are unusable (exc_state[b] >= kExcReplay: pool exhausted or non-finite weights,
// chaotic inputs only) every column of that image replays the recurrence.
constexpr int kBwdThreads = 512;
constexpr int kBwdLight = 12;            // entries a single thread sums; heavier columns go to a warp
constexpr int kBwdQueue = 512;
constexpr int kBwdMaxStages = 8;

__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct BwdArgs {
  const float* g; float* gin;
  int B, C, N, M, CT, ntiles, stages, tiles_per_cta, vec;
  const int* route_ptr; const int* route_q;
  const int* exc_start; const int* exc_cnt; const int* exc_l; const float* exc_w; const int* exc_state;
  const int* ind; const int* mask_idx; const float* wn; const float* wo;
  float triple_w;
  int ninfo, nexc_s, ms; const int* mcount;
};

__global__ void __launch_bounds__(kBwdThreads, 1) shift_bwd_kernel(const BwdArgs a) {
  extern __shared__ __align__(128) float bwd_smem[];      // ring[stages][CT*N] | delta[CT*N] | spec[N] | rq[N] | info[ninfo] | el[E] | ew[E]
  __shared__ int heavy[kBwdQueue];
  __shared__ int4 heavy_info[kBwdQueue];
  __shared__ int nheavy, nspec_s;
  __shared__ __align__(8) unsigned long long bars[kBwdMaxStages];
  const int N = a.N, CT = a.CT, S = a.stages, B = a.B;
  const int N4 = (N + 3) & ~3;
  const int tile_elems = CT * N;
  const uint32_t tile_bytes = (uint32_t)tile_elems * 4u;
  float* ring = bwd_smem;
  float* delta = ring + (size_t)S * tile_elems;
  int* spec = reinterpret_cast<int*>(delta + (((size_t)tile_elems + 3) & ~(size_t)3));
  int* rq_s = spec + N4;
  int4* info = reinterpret_cast<int4*>(rq_s + N4);       // per listed column: first route, routes, first exception, exceptions
  int* el_s = reinterpret_cast<int*>(info + a.ninfo);
  float* ew_s = reinterpret_cast<float*>(el_s + a.nexc_s);
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;

  const long long total_tiles = (long long)B * a.ntiles;
  const long long g0 = (long long)blockIdx.x * a.tiles_per_cta;
  const int n = (int)max(0ll, min(total_tiles, g0 + a.tiles_per_cta) - g0);
  if (n <= 0) return;
  const float* gsrc = a.g + (size_t)g0 * tile_elems;
  float* gdst = a.gin + (size_t)g0 * tile_elems;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(smem_u32(&bars[s]), 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue_load = [&](int i) {                          // thread 0 only
    const int s = i % S;
    mbar_expect_tx(smem_u32(&bars[s]), tile_bytes);
    bulk_g2s(smem_u32(ring + (size_t)s * tile_elems), gsrc + (size_t)i * tile_elems, tile_bytes, smem_u32(&bars[s]));
  };
  if (a.vec && tid == 0)
    for (int i = 0; i < min(S - 1, n); ++i) issue_load(i);

  // per-image correction program
  int cur_b = -1, nspec = 0, nh = 0, Mc = 0;
  bool overflow = false, lists = false;
  const int* gptr = nullptr; const int* estart = nullptr; const int* ecnt = nullptr;
  const int* elx = nullptr; const float* ewx = nullptr; const int* mask_idx = nullptr;
  auto setup = [&](int b) {
    if (tid == 0) {
      nheavy = 0;
      nspec_s = 0;
    }
    __syncthreads();
    Mc = a.mcount ? a.mcount[b] : a.M;
    mask_idx = a.mask_idx ? a.mask_idx + (size_t)b * a.ms : nullptr;
    const int ecount = (a.M > 1 && a.exc_state) ? a.exc_state[b] : 0;
    overflow = ecount >= kExcReplay;
    lists = ecount > 0 && !overflow;
    gptr = a.route_ptr + (size_t)b * (N + 1);
    const int* grq = a.route_q + (size_t)b * N;
    ecnt = a.exc_cnt ? a.exc_cnt + (size_t)b * N : nullptr;
    estart = a.exc_start ? a.exc_start + (size_t)b * N : nullptr;
    const size_t ebase = lists ? (size_t)a.exc_state[B + b] : 0;
    const int* el = a.exc_l ? a.exc_l + ebase : nullptr;
    const float* ew = a.exc_w ? a.exc_w + ebase : nullptr;
    const int etotal = lists ? ecount : 0;
    const bool exc_in_smem = etotal <= a.nexc_s;
    for (int p = tid; p < N; p += nthreads) {
      const int r0 = __ldg(gptr + p), nr = __ldg(gptr + p + 1) - r0;
      const int ne = lists ? __ldg(ecnt + p) : 0;
      const int work = nr + ne;
      if (overflow || work > 0) {
        const int4 rec = make_int4(r0, nr, lists ? __ldg(estart + p) : 0, ne);
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
          if (k < a.ninfo) info[k] = rec;
        }
      }
    }
    for (int i = tid; i < N; i += nthreads) rq_s[i] = __ldg(grq + i);
    if (exc_in_smem)
      for (int i = tid; i < etotal; i += nthreads) {
        el_s[i] = __ldg(el + i);
        ew_s[i] = __ldg(ew + i);
      }
    __syncthreads();
    elx = exc_in_smem ? el_s : el;
    ewx = exc_in_smem ? ew_s : ew;
    nspec = nspec_s;
    nh = min(nheavy, kBwdQueue);
    cur_b = b;
  };

  for (int i = 0; i < n; ++i) {
    const long long gid = g0 + i;
    const int b = (int)(gid / a.ntiles);
    if (b != cur_b) setup(b);
    const int s = a.vec ? i % S : 0;
    float* grow = ring + (size_t)s * tile_elems;
    if (a.vec) {
      mbar_wait(smem_u32(&bars[s]), (uint32_t)(i / S) & 1u);
    } else {                                               // unaligned tensors: plain loads, one stage
      for (int e = tid; e < tile_elems; e += nthreads) grow[e] = __ldg(gsrc + (size_t)i * tile_elems + e);
      __syncthreads();
    }

    // ---- phase A: deltas (reads the unmodified tile only) ----
    const int nlight = nspec * CT;
    for (int it = tid; it < nlight; it += nthreads) {
      const int ch = it / nspec, k = it - ch * nspec;
      const int p = spec[k];
      int4 rec;
      if (k < a.ninfo) rec = info[k];
      else rec = make_int4(__ldg(gptr + p), __ldg(gptr + p + 1) - __ldg(gptr + p), lists ? __ldg(estart + p) : 0,
                           lists ? __ldg(ecnt + p) : 0);
      const float* row = grow + (size_t)ch * N;
      float acc = 0.f;
      for (int r = rec.x; r < rec.x + rec.y; ++r) acc += row[rq_s[r]];
      for (int e = rec.z; e < rec.z + rec.w; ++e) acc = fmaf(ewx[e], row[elx[e]], acc);
      if (overflow) {                                      // rare, slow, bit-faithful replay of the recurrence
        float e = (a.ind[(size_t)b * N + mask_idx[0]] == p) ? 1.f : 0.f;
        for (int l = 1; l < Mc; ++l) {
          const int ql = mask_idx[l];
          e = __fmul_rn(e, a.wn[(size_t)b * a.M + l]);
          if (a.ind[(size_t)b * N + ql] == p) e = __fadd_rn(e, a.wo[(size_t)b * a.M + l]);
          if (!(fabsf(e) < 1.0f)) acc = fmaf(trunc_as_reference(e), row[ql], acc);
        }
      }
      delta[it] = acc;
    }
    const int nhw = nh * CT;
    for (int hw = warp; hw < nhw; hw += nwarps) {          // hub columns: one warp per (column, channel)
      const int ch = hw / nh, h = hw - ch * nh;
      const int4 rec = heavy_info[h];
      const float* row = grow + (size_t)ch * N;
      float acc = 0.f;
      for (int r = rec.x + lane; r < rec.x + rec.y; r += 32) acc += row[rq_s[r]];
      for (int e = rec.z + lane; e < rec.z + rec.w; e += 32) acc = fmaf(ewx[e], row[elx[e]], acc);
      acc = warp_sum(acc);
      if (lane == 0) delta[nlight + hw] = acc;
    }
    __syncthreads();

    // ---- phase B: g + weighted * triple_w in place                                               :173
    for (int it = tid; it < nlight; it += nthreads) {
      const int ch = it / nspec, k = it - ch * nspec;
      float* cell = grow + (size_t)ch * N + spec[k];
      *cell = __fadd_rn(*cell, __fmul_rn(delta[it], a.triple_w));
    }
    for (int hw = tid; hw < nhw; hw += nthreads) {
      const int ch = hw / nh, h = hw - ch * nh;
      float* cell = grow + (size_t)ch * N + heavy[h];
      *cell = __fadd_rn(*cell, __fmul_rn(delta[nlight + hw], a.triple_w));
    }
    if (a.vec) {
      fence_proxy_async();                                 // the in-place writes become visible to the bulk store
      __syncthreads();
      if (tid == 0) {
        bulk_s2g(gdst + (size_t)i * tile_elems, smem_u32(grow), tile_bytes);
        bulk_commit();
        if (i + S - 1 < n) {
          bulk_wait_read<1>();                             // the store of tile i-1 has left its stage
          issue_load(i + S - 1);
        }
      }
    } else {
      __syncthreads();
      for (int e = tid; e < tile_elems; e += nthreads) gdst[(size_t)i * tile_elems + e] = grow[e];
      __syncthreads();
    }
  }
  if (a.vec && tid == 0) bulk_wait_all();                  // shared memory must outlive the last stores
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
  // channel rows per tile: a power of two <= 16 that keeps a tile near the target size (IPSR_BWD_TILE_KB, default
  // 32 KiB); N % 4 != 0 needs CT % 4 == 0 so that every tile stays a whole number of 16-byte units for the bulk copies
  static const int tile_kb = [] {
    const char* e = getenv("IPSR_BWD_TILE_KB");
    const int v = e ? atoi(e) : 32;
    return v >= 4 && v <= 64 ? v : 32;
  }();
  int CT = 16;
  while (CT > 1 && (C % CT != 0 || (size_t)CT * N * sizeof(float) > (size_t)tile_kb * 1024)) CT >>= 1;
  if ((N & 3) != 0 && CT < 4 && C % 4 == 0) CT = 4;
  const size_t tile_bytes = (size_t)CT * N * sizeof(float);
  const bool vec = (tile_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(gin) & 15) == 0);
  // shared memory: the ring, the delta buffer (one tile), the column list [N] + the CSR row list [N], then up to 1024
  // per-column records (16 B each) and 2048 exception entries (8 B each): what does not fit is read from global memory
  const size_t N4 = (size_t)((N + 3) & ~3);
  int ninfo = N < 1024 ? N : 1024;
  int nexc_s = (M > 1) ? 2048 : 0;
  const size_t delta_bytes = (((size_t)CT * N + 3) & ~(size_t)3) * sizeof(float);
  const size_t fixed = delta_bytes + 2 * N4 * sizeof(int) + (size_t)ninfo * 16 + (size_t)nexc_s * 8;
  const size_t budget = 227 * 1024 - 12 * 1024;            // static: the hub queue and the barriers
  IPSR_REQUIRE(fixed + (vec ? 2 : 1) * tile_bytes <= budget, IPSR_ERR_UNSUPPORTED, "ipsr_shift_bwd: N=%d too large", N);
  int stages = vec ? (int)((budget - fixed) / tile_bytes) : 1;
  if (stages > kBwdMaxStages) stages = kBwdMaxStages;
  const size_t smem = fixed + (size_t)stages * tile_bytes;
  // persistent CTAs, one per SM, each owning a contiguous run of tiles (image-major: a CTA sees 1-2 images)
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = C / CT;
  const long long total_tiles = (long long)B * ntiles;
  long long tpc = (total_tiles + sms - 1) / sms;
  if (tpc < 1) tpc = 1;
  const long long ctas = (total_tiles + tpc - 1) / tpc;
  IPSR_REQUIRE(tpc <= 0x7FFFFFFFll, IPSR_ERR_UNSUPPORTED, "ipsr_shift_bwd: too many tiles");
  cudaError_t e = cudaFuncSetAttribute(shift_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "shift_bwd smem attribute: %s", cudaGetErrorString(e));
  BwdArgs a;
  a.g = g; a.gin = gin;
  a.B = B; a.C = C; a.N = N; a.M = M; a.CT = CT; a.ntiles = ntiles; a.stages = stages; a.tiles_per_cta = (int)tpc; a.vec = vec ? 1 : 0;
  a.route_ptr = route_ptr; a.route_q = route_q;
  a.exc_start = exc_start; a.exc_cnt = exc_cnt; a.exc_l = exc_l; a.exc_w = exc_w; a.exc_state = exc_total;
  a.ind = ind; a.mask_idx = mask_idx; a.wn = wn; a.wo = wo;
  a.triple_w = triple_w;
  a.ninfo = ninfo; a.nexc_s = nexc_s; a.ms = mask_stride; a.mcount = m_count;
  shift_bwd_kernel<<<(unsigned)ctas, kBwdThreads, smem, as_stream(stream)>>>(a);
  return check_launch("ipsr_shift_bwd");
}
