// (e) backward of the shift operator, routed through the saved indices.
//
// Replaces models/IPSRFunction.py:144-178: N row-gathers out of an int64 [N,N,H,W] tensor, a
// zeroed [B,N,N] float matrix and a dense torch.mm(W^T [N,N], g [N,C]).  The reference keeps its
// attention in a LongTensor (:36,134), so W = trunc(A): a 0/1 matrix with one 1 per unmasked row
// (and for the first masked row), plus -- only for ill-conditioned inputs -- the few entries of the
// blended rows whose magnitude reaches 1.  gin = g + triple_w * W^T g is therefore a segment sum
// over "unit routes" plus a short exception list; no N x N object exists anywhere.
#include <stdlib.h>

#include <algorithm>

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
// grid = (parts, B): a CTA owns a contiguous range of channel tiles (CT rows each, C % CT == 0) of ONE image.  Most bank
// columns receive nothing (a non-negative reference concentrates the matches on a few hundred patches), so gin is a copy
// of g plus corrections in the columns that do.
//
// Warp 0 is the DMA warp: one thread issues the bulk async copy of every tile into a two-stage ring (the CT rows of a tile
// are contiguous in NCHW) and -- when the tile was corrected in place -- the bulk async STORE of the finished tile, so no
// compute thread ever copies a byte and the store of tile t drains while tile t+1 is corrected.  The other warps, per tile:
//   phase A  the columns that receive something -- a compact list built ONCE per CTA from the CSR and the exception
//            directory, staged in shared memory -- are summed: one thread per column (all CT channels in registers;
//            the sum stays in the thread's registers), hub columns (hundreds of routes) by whole warps (lane-strided partial
//            sums in ascending q, then a fixed xor tree) into a small shared buffer: deterministic;
//   phase B  (after a barrier of the compute warps: a column can also be a SOURCE of another column)
//            tile[ch][p] = g + triple_w * sum in place; fence; every warp arrives on the stage's `done` barrier.
// Images with more listed columns than compute threads (the self-match case: every column receives one route) or whose
// tensors are not 16-byte aligned take the round-1 path in the same kernel: the compute warps copy the tile out themselves
// and write the corrected columns straight to global memory.
// If the exception lists of an image are unusable (exc_total[b] >= kExcReplay: the batch's pool is exhausted or a blend
// weight is non-finite, chaotic inputs only) every column of that image replays the recurrence.
constexpr int kBwdLight = 24;            // entries a single thread sums; heavier columns go to a warp (IPSR_BWD_LIGHT: A/B runs
                                         //   at 64 x 64 x 256: 4 -> 248 us, 8 -> 204 us, 16 -> 189 us, 24 -> 184 us)
constexpr int kBwdPiece = 256;           // entries one warp sums: longer columns (chaotic images: thousands of entries) are cut in pieces
constexpr int kBwdQueue = 512;
constexpr int kBwdHubInplace = 256;     // hub pieces whose sums fit the in-place path's shared buffer

__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void compute_barrier(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

template <int CT>
__global__ void __launch_bounds__(CT >= 8 ? 512 : 1024, CT >= 8 ? 2 : 1)
shift_bwd_kernel(const float* __restrict__ g, int C, int N, int M, int tiles_per_cta, const int* __restrict__ route_ptr,
                 const int* __restrict__ route_q, const int* __restrict__ exc_start, const int* __restrict__ exc_cnt,
                 const int* __restrict__ exc_l, const float* __restrict__ exc_w, const int* __restrict__ exc_total,
                 int exc_cap, const int* __restrict__ ind, const int* __restrict__ mask_idx,
                 const float* __restrict__ wn, const float* __restrict__ wo, float triple_w, float* __restrict__ gin,
                 int ninfo, int nexc_s, int ms, const int* __restrict__ mcount, int S, int light, int qcap) {
  extern __shared__ __align__(128) float bwd_smem[];      // rows[S][CT*N] | spec[N] | info[N] int4 | rq[N] | el[E] | ew[E]
  __shared__ int heavy[kBwdQueue];
  __shared__ int4 heavy_info[kBwdQueue];
  __shared__ float hdelta[2][kBwdHubInplace][CT];            // hub sums of the tile in flight (parity of the tile)
  __shared__ int nheavy, nspec_s, inplace_s;
  __shared__ __align__(8) unsigned long long full_bar[4], done_bar[4];   // S <= 4 stages
  pdl_trigger();
  pdl_wait();
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
  int* spec = reinterpret_cast<int*>(bwd_smem + (size_t)S * tile_elems);
  int4* info = reinterpret_cast<int4*>(spec + ((N + 3) & ~3));      // per listed column: first route, routes, first exception, exceptions
  int* rq_s = reinterpret_cast<int*>(info + ninfo);                 // the CSR's row list
  int* el_s = rq_s + N;                                             // exception entries (when they fit)
  float* ew_s = reinterpret_cast<float*>(el_s + nexc_s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncompute = blockDim.x - 32, ncwarps = ncompute >> 5;    // warp 0 is the DMA warp
  const int ctid = threadIdx.x - 32, cwarp = warp - 1;
  const float* gimg = g + (size_t)b * C * N;
  float* oimg = gin + (size_t)b * C * N;
  const bool vec = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(gimg) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(oimg) & 15) == 0);

  if (threadIdx.x == 0) {
    nheavy = 0;
    nspec_s = 0;
    inplace_s = 0;
    for (int s2 = 0; s2 < S; ++s2) {
      mbar_init(smem_u32(&full_bar[s2]), 1);
      mbar_init(smem_u32(&done_bar[s2]), (uint32_t)ncwarps);
    }
    mbar_fence_init();
  }
  // A column whose pieces do not fit the queue any more goes to the light list, but it has already advanced the queue
  // cursor: the slots it leaves behind must read as "later piece, nothing to sum".
  for (int i = threadIdx.x; i < kBwdQueue; i += blockDim.x) {
    heavy[i] = -1;
    heavy_info[i] = make_int4(0, 0, 0, 0);
  }
  __syncthreads();
  auto issue_load = [&](int t) {                          // DMA thread
    const int s2 = (t - t0) % S;
    mbar_expect_tx(smem_u32(&full_bar[s2]), tile_bytes);
    bulk_g2s(smem_u32(rows0 + (size_t)s2 * tile_elems), gimg + (size_t)t * tile_elems, tile_bytes, smem_u32(&full_bar[s2]));
  };
  if (vec && threadIdx.x == 0)
    for (int t = t0; t < min(t1, t0 + S); ++t) issue_load(t);

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
  for (int p = threadIdx.x; p < N; p += blockDim.x) {
    const int r0 = __ldg(gptr + p), n = __ldg(gptr + p + 1) - r0;
    const int ne = lists ? __ldg(ecnt + p) : 0;
    const int work = n + ne;
    if (overflow || work > 0) {
      const int4 rec = make_int4(r0, n, lists ? __ldg(estart + p) : 0, ne);
      bool queued = false;
      if (work > light && !overflow) {
        // a warp per piece of <= kBwdPiece entries of the column's list (routes, then exceptions); heavy[] holds the column
        // for the first piece (with the piece count in the high bits) and -1 for the others
        const int np = (work + kBwdPiece - 1) / kBwdPiece;
        const int slot = atomicAdd(&nheavy, np);
        if (slot + np <= qcap) {
          for (int j = 0; j < np; ++j) {
            const int lo = j * kBwdPiece, hi = min(work, lo + kBwdPiece);
            const int ra = min(lo, n), rb = min(hi, n);                    // routes [ra, rb)
            const int ea = max(lo - n, 0), eb = max(hi - n, 0);             // exceptions [ea, eb)
            heavy[slot + j] = (j == 0) ? (p | (np << 20)) : -1;
            heavy_info[slot + j] = make_int4(r0 + ra, rb - ra, rec.z + ea, eb - ea);
          }
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
  for (int i = threadIdx.x; i < N; i += blockDim.x) rq_s[i] = __ldg(grq + i);
  if (exc_in_smem)
    for (int i = threadIdx.x; i < etotal; i += blockDim.x) {
      el_s[i] = __ldg(el + i);
      ew_s[i] = __ldg(ew + i);
    }
  __syncthreads();
  const int* elx = exc_in_smem ? el_s : el;
  const float* ewx = exc_in_smem ? ew_s : ew;
  const int nspec = nspec_s;
  const int nh = min(nheavy, qcap);                     // (a column whose pieces did not fit the queue went to the light list)
  // in place (DMA warp stores the finished tile) when every listed column has its own compute thread, so that the sums of
  // phase A can wait in registers for phase B
  const bool inplace = vec && !overflow && nspec <= ncompute && nspec <= ninfo && nh <= kBwdHubInplace;

  if (warp == 0) {
    // ------------------------------------------------------------------ DMA warp
    if (lane == 0 && vec) {
      for (int t = t0; t < t1; ++t) {
        const int s2 = (t - t0) % S;
        mbar_wait(smem_u32(&done_bar[s2]), (uint32_t)((t - t0) / S) & 1u);    // the compute warps are done with this stage
        if (inplace) {
          bulk_s2g(oimg + (size_t)t * tile_elems, smem_u32(rows0 + (size_t)s2 * tile_elems), tile_bytes);
          bulk_commit();
        }
        if (t + S < t1) {
          if (inplace) bulk_wait_read_all();               // the store has left the stage: refill it
          issue_load(t + S);
        }
      }
      if (inplace) bulk_wait_all();                        // shared memory must outlive the last stores
    }
    return;
  }

  // ------------------------------------------------------------------ compute warps
  for (int t = t0; t < t1; ++t) {
    const int buf = vec ? (t - t0) % S : 0;
    float* grow = rows0 + (size_t)buf * tile_elems;
    float* ob = oimg + (size_t)t * tile_elems;
    if (vec) {
      mbar_wait(smem_u32(&full_bar[buf]), (uint32_t)((t - t0) / S) & 1u);
    } else {                                               // unaligned tensors: plain loads by the compute warps
      for (int i = ctid; i < tile_elems; i += ncompute) grow[i] = __ldg(gimg + (size_t)t * tile_elems + i);
      compute_barrier(ncompute);
    }

    if (inplace) {
      // ---- phase A: sums (the unmodified tile is read only) ----
      float acc[CT];
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) acc[ch] = 0.f;
      const int k = ctid;                                  // nspec <= ncompute: one column per thread
      int p = -1;
      if (k < nspec) {
        p = spec[k];
        const int4 rec = info[k];
        const int r0 = rec.x, r1 = rec.x + rec.y, es = rec.z, ne = rec.w;
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
      }
      float (*hd)[CT] = hdelta[(t - t0) & 1];
      for (int h = cwarp; h < nh; h += ncwarps) {          // hub columns: one warp per piece
        const int4 rec = heavy_info[h];
        const int r0 = rec.x, r1 = rec.x + rec.y;
        float hacc[CT];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) hacc[ch] = 0.f;
        for (int r = r0 + lane; r < r1; r += 32) {
          const int q = rq_s[r];
#pragma unroll
          for (int ch = 0; ch < CT; ++ch) hacc[ch] += grow[ch * N + q];
        }
        if (lists) {
          const int ne = rec.w, es = rec.z;
          for (int e = lane; e < ne; e += 32) {
            const int q = elx[es + e];
            const float w = ewx[es + e];
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) hacc[ch] = fmaf(w, grow[ch * N + q], hacc[ch]);
          }
        }
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) hacc[ch] = warp_sum(hacc[ch]);
        if (lane == 0) {
#pragma unroll
          for (int ch = 0; ch < CT; ++ch) hd[h][ch] = hacc[ch];
        }
      }
      compute_barrier(ncompute);                           // every read of the tile is done
      // ---- phase B: g + weighted * triple_w in place                                                :173
      if (p >= 0) {
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) grow[ch * N + p] = __fadd_rn(grow[ch * N + p], __fmul_rn(acc[ch], triple_w));
      }
      for (int i = ctid; i < nh * CT; i += ncompute) {
        const int h = i / CT, ch = i - h * CT;
        const int code = heavy[h];
        if (code < 0) continue;                            // a later piece: added by the column's first piece
        const int pcol = code & 0xFFFFF, np = code >> 20;
        float sum = hd[h][ch];
        for (int j = 1; j < np; ++j) sum += hd[h + j][ch];   // piece order: deterministic
        float* cell = grow + ch * N + pcol;
        *cell = __fadd_rn(*cell, __fmul_rn(sum, triple_w));
      }
      fence_proxy_async();                                 // the corrected tile becomes visible to the bulk store
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&done_bar[buf]));
      continue;
    }

    // ---- round-1 path: (1) copy-out g + triple_w * 0, (2) corrected columns straight to global memory ----
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(grow);
      float4* d4 = reinterpret_cast<float4*>(ob);
      for (int i = ctid; i < tile_elems / 4; i += ncompute) d4[i] = s4[i];
    } else {
      for (int i = ctid; i < tile_elems; i += ncompute) ob[i] = grow[i];
    }
    compute_barrier(ncompute);                             // the corrections below overwrite some of these stores
    for (int k = ctid; k < nspec; k += ncompute) {
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
    for (int h = cwarp; h < nh; h += ncwarps) {              // hub columns: one warp each (all pieces of the column)
      const int code = heavy[h];
      if (code < 0) continue;
      const int p = code & 0xFFFFF, np = code >> 20;
      float acc[CT];
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) acc[ch] = 0.f;
      for (int j = 0; j < np; ++j) {
        const int4 rec = heavy_info[h + j];
        for (int r = rec.x + lane; r < rec.x + rec.y; r += 32) {
          const int q = rq_s[r];
#pragma unroll
          for (int ch = 0; ch < CT; ++ch) acc[ch] += grow[ch * N + q];
        }
        if (lists) {
          for (int e = rec.z + lane; e < rec.z + rec.w; e += 32) {
            const int q = elx[e];
            const float w = ewx[e];
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) acc[ch] = fmaf(w, grow[ch * N + q], acc[ch]);
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
    if (vec) {
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&done_bar[buf]));  // this warp is done reading the stage
    } else {
      compute_barrier(ncompute);
    }
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
  // channel rows per tile: the largest of 8, 4, 2, 1 that divides C and keeps a tile within 32 KiB (N > 2048: 64 KiB);
  // ring depth 2.  IPSR_BWD_TILE_KB / IPSR_BWD_STAGES override both (A/B runs: at 64 x 64 x 256 the kernel takes 197 us
  // with 64 KiB tiles, 255 us with 32 KiB tiles whatever the ring depth, 421 us with 16 KiB tiles).
  static const int env_tile_kb = [] { const char* e = getenv("IPSR_BWD_TILE_KB"); return e ? atoi(e) : 0; }();
  static const int env_stages = [] { const char* e = getenv("IPSR_BWD_STAGES"); return e ? atoi(e) : 0; }();
  int CT = 8;
  const size_t tile_cap = (size_t)(env_tile_kb > 0 ? env_tile_kb : (N <= 2048 ? 32 : 64)) * 1024;
  while (CT > 1 && (C % CT != 0 || (size_t)CT * N * sizeof(float) > tile_cap)) CT >>= 1;
  // shared memory: the ring + the column list [N] + the CSR row list [N], then as much of the per-column records
  // (16 B each) and of the exception entries (8 B each) as fits: what does not fit is read from global memory
  const size_t tile_bytes = (size_t)CT * N * sizeof(float);
  const size_t lists_bytes = 2 * (size_t)((N + 3) & ~3) * sizeof(int);
  const size_t budget = (N <= 2048 ? 96 : 215) * 1024 - 28 * 1024;      // minus the static queues / hub sums
  int S = 2;                                               // (measured: deeper rings of smaller tiles lose -- the index walk
  if (env_stages >= 2 && env_stages <= 4) S = env_stages;   //  of the corrections is paid per tile, whatever its channel count)
  const size_t base_smem = (size_t)S * tile_bytes + lists_bytes;
  IPSR_REQUIRE(base_smem <= 227 * 1024 - 28 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_shift_bwd: N=%d too large", N);
  const size_t room = budget > base_smem ? budget - base_smem : 0;
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
               const int*, int, const int*, const int*, const float*, const float*, float, float*, int, int, int, const int*, int, int, int) = nullptr;
  switch (CT) {
    case 8: kern = shift_bwd_kernel<8>; break;
    case 4: kern = shift_bwd_kernel<4>; break;
    case 2: kern = shift_bwd_kernel<2>; break;
    default: kern = shift_bwd_kernel<1>; break;
  }
  if (smem + 28 * 1024 > 48 * 1024) {                       // static (queues, hub sums) + dynamic shared memory above the default limit
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "shift_bwd smem attribute: %s", cudaGetErrorString(e));
  }
  // IPSR_BWD_QUEUE (tests; read per call): a smaller hub queue, so that a modest case exercises its overflow
  int qcap = kBwdQueue;
  if (const char* e = getenv("IPSR_BWD_QUEUE")) qcap = std::max(1, std::min(kBwdQueue, atoi(e)));
  static const int light = [] { const char* e = getenv("IPSR_BWD_LIGHT"); return e ? atoi(e) : kBwdLight; }();
  cudaError_t le = launch_pdl(kern, dim3(parts, B), dim3(threads), smem, as_stream(stream), g, C, N, M, tiles_per_cta, route_ptr, route_q,
                              exc_start, exc_cnt, exc_l, exc_w, exc_total, exc_cap, ind, mask_idx, wn, wo, triple_w, gin, ninfo, nexc_s,
                              mask_stride, m_count, S, light, qcap);
  IPSR_REQUIRE(le == cudaSuccess, IPSR_ERR_CUDA, "ipsr_shift_bwd: launch failed: %s", cudaGetErrorString(le));
  return check_launch("ipsr_shift_bwd");
}
