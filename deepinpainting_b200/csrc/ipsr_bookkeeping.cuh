// Backward bookkeeping built in the forward: unit routes (stable counting sort by arg-max column)
// and exception lists (attention entries of blended rows that survive the reference's int64 store).
// Device bodies live here so that they can run as stand-alone kernels (ipsr_build_routes /
// ipsr_build_exceptions) and as extra CTAs of the fused paste launch of ipsr_shift_forward.
#pragma once
#include "ipsr_common.cuh"

namespace ipsr {

// ---------------------------------------------------------------------------------------------
// unit routes: stable counting sort of {q : unmasked or q == q_0} by p = ind[q]
// ---------------------------------------------------------------------------------------------
// one CTA (any multiple of 32 threads up to 1024) per image; rsm: (2N+1) ints of shared memory
__device__ __forceinline__ void
build_routes_cta(int b, int* rsm, const int* __restrict__ ind, const int* __restrict__ flag, const int* __restrict__ mask_idx,
                 int N, int M, int* __restrict__ route_ptr, int* __restrict__ route_q,
                 int ms = 0, const int* __restrict__ mcount = nullptr) {
  // per-image masks: flag / mask_idx are [B][ms] (ms = N), mcount[b] masked positions; ms = 0: one mask for the batch
  flag += (size_t)b * ms;
  mask_idx += (size_t)b * ms;
  if (mcount) M = mcount[b];
  int* cursor = rsm;            // [N+1] counts -> exclusive offsets -> running cursors
  int* key = rsm + (N + 1);     // [N]   p = ind[q] for routed q, -1 otherwise
  __shared__ int warp_tot[32];
  __shared__ int carry;
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
  // exclusive scan of cursor[0..N) in chunks of blockDim.x, cursor[N] = total
  const int nwarps = blockDim.x >> 5;
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
      for (int w = 0; w < nwarps; ++w) t += warp_tot[w];
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
// exceptions: replay row_l[p] = row_{l-1}[p]*wn_l (+ wo_l if p == p_l) -- only columns that are the
// match of some masked position can ever be non-zero, so the work is one thread per masked position
// l0 that is the FIRST occurrence of its column p_{l0}, walking l = l0+1 .. M-1.
// One CTA per image.  Entries go into ONE POOL shared by the whole batch (exc_l / exc_w [exc_cap]): an image takes
// the contiguous range [base, base + count) it reserves with one atomicAdd on the pool cursor, so a chaotic image
// (tens of thousands of surviving entries with signed inputs) borrows the room the well-behaved images do not use.
// exc_state is int32 [2B + 2]: [b] = entries of image b (>= kExcReplay: the lists are unusable -- pool exhausted or
// non-finite weights -- and the backward replays the recurrence), [B + b] = base of image b, [2B] = pool cursor.
// ---------------------------------------------------------------------------------------------
constexpr int kExcChunk = 1024;
constexpr int kExcThreads = 1024;
constexpr int kExcReplay = 0x3FFFFFFF;

// One owner = one bank column p, identified by the first masked step l0 whose match is p.  Walks
// row_l[p] = row_{l-1}[p] * wn_l (+ wo_l if p_l == p) for l = l0 .. M-1 in the reference's operation order and
// counts (WRITE: stores) every entry that survives the float -> int64 store (l >= 1, |e| >= 1): the position
// q_l = mask_idx[l] of the row and its truncated weight.  The walk is a chain of M dependent steps executed by a
// lone warp, so the step is kept to a handful of instructions: operands come as 16-byte shared-memory vectors, the
// addend is selected off the dependent chain (adding +0 when p_l != p leaves every non-zero value untouched).
template <bool WRITE>
__device__ __forceinline__ int replay_owner(int l0, int p, int base, int n, const float* s_wn, const float* s_wo,
                                            const int* s_p, const int* __restrict__ mask_idx, float& e, int cnt,
                                            int* __restrict__ out_q, float* __restrict__ out_w) {
  // processes steps l in [max(base, l0), base+n) of this chunk
  auto step = [&](int i, float wn_i, float wo_i, int p_i) {
    e = __fadd_rn(__fmul_rn(e, wn_i), (p_i == p) ? wo_i : 0.f);     // row * wn; row[p_l] += wo        :123-124
    if (!(fabsf(e) < 1.0f)) {                                        // survives the int64 store        :134
      if (WRITE) {
        out_q[cnt] = mask_idx[base + i];
        out_w[cnt] = trunc_as_reference(e);
      }
      ++cnt;
    }
  };
  int i = l0 - base;
  if (i >= n) return cnt;
  if (i >= 0) {                                             // first appearance: row[p] = 0*wn + wo (or 1 at l = 0)
    e = (l0 == 0) ? 1.f : s_wo[i];
    if (l0 >= 1 && !(fabsf(e) < 1.0f)) {
      if (WRITE) {
        out_q[cnt] = mask_idx[l0];
        out_w[cnt] = trunc_as_reference(e);
      }
      ++cnt;
    }
    ++i;
  } else {
    i = 0;
  }
  for (; (i & 3) != 0 && i < n; ++i) step(i, s_wn[i], s_wo[i], s_p[i]);
  for (; i + 4 <= n; i += 4) {
    const float4 a = *reinterpret_cast<const float4*>(s_wn + i);
    const float4 c = *reinterpret_cast<const float4*>(s_wo + i);
    const int4 pp = *reinterpret_cast<const int4*>(s_p + i);
    step(i, a.x, c.x, pp.x);
    step(i + 1, a.y, c.y, pp.y);
    step(i + 2, a.z, c.z, pp.z);
    step(i + 3, a.w, c.w, pp.w);
  }
  for (; i < n; ++i) step(i, s_wn[i], s_wo[i], s_p[i]);
  return cnt;
}

// shared memory of one exceptions CTA, in 4-byte words
__host__ __device__ inline size_t exc_smem_words(int N, int M) {
  return (size_t)((N + 3) & ~3) + 2 * (size_t)((M + 3) & ~3) + 3 * kExcChunk;
}

// fsm: first-occurrence table [N] ints, owner list [M] ints, slot offsets [M] ints (both padded to 4), then
// 3*kExcChunk staging words (16-byte aligned)
__device__ __forceinline__ void
build_exceptions_cta(int b, int B, void* fsm, const int* __restrict__ ind, const int* __restrict__ mask_idx,
                     const float* __restrict__ wn, const float* __restrict__ wo, int N, int M,
                     int* __restrict__ exc_start, int* __restrict__ exc_cnt, int* __restrict__ exc_l,
                     float* __restrict__ exc_w, int* __restrict__ exc_state, int exc_cap,
                     int ms = 0, const int* __restrict__ mcount = nullptr) {
  int* first = reinterpret_cast<int*>(fsm);                 // [N] first masked step whose match is p, or INT_MAX
  int* owners = first + ((N + 3) & ~3);                     // [M] compact list of the owner steps
  int* slot = owners + ((M + 3) & ~3);                      // [M] offset of owner k's entries inside the image's range
  float* s_wn = reinterpret_cast<float*>(slot + ((M + 3) & ~3));
  const int Mstride = M;                                    // rows of wn / wo are M (the batch maximum) apart
  mask_idx += (size_t)b * ms;                               // per-image masks (see build_routes_cta)
  if (mcount) M = mcount[b];
  float* s_wo = s_wn + kExcChunk;
  int* s_p = reinterpret_cast<int*>(s_wo + kExcChunk);
  __shared__ int nonfinite_w, nown_s, total_s, base_s;
  const int* ind_b = ind + (size_t)b * N;
  const float* wnb = wn + (size_t)b * Mstride;
  const float* wob = wo + (size_t)b * Mstride;
  for (int p = threadIdx.x; p < N; p += blockDim.x) {
    first[p] = 0x7FFFFFFF;
    exc_start[(size_t)b * N + p] = 0;                       // default: no exceptions in this column
    exc_cnt[(size_t)b * N + p] = 0;
  }
  if (threadIdx.x == 0) {
    nonfinite_w = 0;
    nown_s = 0;
    total_s = 0;
    base_s = 0;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < M; l += blockDim.x) atomicMin(&first[ind_b[mask_idx[l]]], l);
  __syncthreads();
  // compact owner list: only the first occurrence of a column walks it, and only a fraction of the masked steps
  // are first occurrences -- compaction keeps every lane of the walking warps busy
  for (int l0 = threadIdx.x; l0 < M; l0 += blockDim.x)
    if (first[ind_b[mask_idx[l0]]] == l0) owners[atomicAdd(&nown_s, 1)] = l0;
  // A non-finite weight turns EVERY column of the later rows into NaN (0 * inf), which the sparse
  // "first occurrence" walk below cannot represent: flag the image so that the backward replays the full
  // recurrence per column (bit-faithful, slow, chaotic inputs only).
  int cur_base = -1;                                        // chunk currently staged (uniform)
  auto stage = [&](int base) {
    if (base == cur_base) return;
    __syncthreads();
    const int n = min(kExcChunk, M - base);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float a = wnb[base + i], c = wob[base + i];
      s_wn[i] = a;
      s_wo[i] = c;
      s_p[i] = ind_b[mask_idx[base + i]];
      if (!(fabsf(a) <= 3.4028234e38f) || !(fabsf(c) <= 3.4028234e38f)) nonfinite_w = 1;
    }
    __syncthreads();
    cur_base = base;
  };
  stage(0);                                                 // also publishes the owner list
  const int nown = nown_s;
  // walk 1: count the surviving entries of every owned column and hand out slots inside the image's range
  // (placement between columns is arbitrary, the order inside a column is ascending l)
  for (int round = 0; round < nown; round += blockDim.x) {
    const int k = round + threadIdx.x;
    const int l0 = (k < nown) ? owners[k] : M;
    const int p = (k < nown) ? ind_b[mask_idx[l0]] : -1;
    float e = 0.f;
    int found = 0;
    for (int base = 0; base < M; base += kExcChunk) {
      stage(base);
      if (p >= 0) found = replay_owner<false>(l0, p, base, min(kExcChunk, M - base), s_wn, s_wo, s_p, mask_idx, e, found, nullptr,
                                              nullptr);
    }
    if (k < nown) slot[k] = found > 0 ? atomicAdd(&total_s, found) : -1;
  }
  __syncthreads();
  const int total = total_s;
  if (threadIdx.x == 0) {
    int base = 0;
    bool ok = !nonfinite_w;
    if (ok && total > 0) {
      base = atomicAdd(exc_state + 2 * B, total);           // one reservation per image
      ok = (long long)base + total <= (long long)exc_cap;
    }
    base_s = ok ? base : -1;
    exc_state[b] = ok ? total : kExcReplay;
    exc_state[B + b] = ok ? base : 0;
  }
  __syncthreads();
  const int ibase = base_s;
  if (ibase < 0 || total == 0) return;
  // walk 2: write
  for (int round = 0; round < nown; round += blockDim.x) {
    const int k = round + threadIdx.x;
    const int l0 = (k < nown) ? owners[k] : M;
    const int p = (k < nown) ? ind_b[mask_idx[l0]] : -1;
    const int sl = (k < nown) ? slot[k] : -1;
    float e = 0.f;
    int cnt = 0;
    for (int base = 0; base < M; base += kExcChunk) {
      stage(base);
      if (sl >= 0)
        cnt = replay_owner<true>(l0, p, base, min(kExcChunk, M - base), s_wn, s_wo, s_p, mask_idx, e, cnt,
                                 exc_l + (size_t)ibase + sl, exc_w + (size_t)ibase + sl);
    }
    if (sl >= 0) {
      exc_start[(size_t)b * N + p] = sl;                    // relative to the image's base
      exc_cnt[(size_t)b * N + p] = cnt;
    }
  }
}

}  // namespace ipsr
