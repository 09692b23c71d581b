// Exact fp32 correlation + arg-max (the recheck path of the tensor mode and the whole of
// IPSR_MODE_EXACT), (max, idx) key packing for the bank-sharded exchange, and MaxCoord on a
// materialised score tensor.
//
// Replaces models/IPSRFunction.py:59 (conv_enc(ref): S[q,p] = <R[q], Xn[p]>) followed by
// util/MaxCoord.py:22 (torch.max over the bank axis) without writing S.
#include "ipsr_common.cuh"

namespace ipsr {

constexpr int kFpTile = 64;     // rows x cols per CTA tile
constexpr int kFpKc = 32;       // channels per shared-memory slab
constexpr int kFpThreads = 256;
constexpr int kFpLd = kFpKc * kFpTile / kFpThreads;   // elements per thread per slab and operand (8)

// grid = (column tiles, row_ctas, B).  A CTA walks the row list of its image in chunks of 64 rows
// and, for each chunk, computes the 64 x 64 scores against its column tile with register-tiled
// FFMA (4 x 4 per thread, channels in ascending order), then folds them into packed[b,q] with
// one 64-bit atomicMax per (row, CTA).  The next 32-channel slab is fetched into registers while
// the current one is multiplied (the sparse recheck lists make this kernel latency-bound).
__global__ void __launch_bounds__(kFpThreads)
corr_fp32_kernel(const float* __restrict__ x, const float* __restrict__ ref, const float* __restrict__ inv_norm,
                 int C, int N, int col_begin, int col_end,
                 const int* __restrict__ list, const int* __restrict__ nlist, long long* __restrict__ packed) {
  __shared__ __align__(16) float Rs[kFpKc][kFpTile];
  __shared__ __align__(16) float Xs[kFpKc][kFpTile];
  __shared__ int rows[kFpTile];

  const int b = blockIdx.z;
  const int nrows = min(nlist[b], N);
  if ((int)blockIdx.y * kFpTile >= nrows) return;
  const int col0 = col_begin + blockIdx.x * kFpTile;
  const float* xb = x + (size_t)b * C * N;
  const float* rb = ref + (size_t)b * C * N;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // slab element e of this thread: channel kk = (tid + e*256) / 64 = tid/64 + 4e, column/row jj = tid % 64
  const int jj = threadIdx.x & 63, kk0 = threadIdx.x >> 6;
  const int pcol = col0 + jj;
  const bool pok = pcol < col_end;
  const float invp = pok ? inv_norm[(size_t)b * N + pcol] : 0.f;

  for (int rc = blockIdx.y; rc * kFpTile < nrows; rc += gridDim.y) {
    __syncthreads();
    if (threadIdx.x < kFpTile) {
      const int i = rc * kFpTile + threadIdx.x;
      rows[threadIdx.x] = (i < nrows) ? list[(size_t)b * N + i] : -1;
    }
    __syncthreads();
    const int qrow = rows[jj];

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float rr[kFpLd], xr[kFpLd];
    auto fetch = [&](int c0) {
#pragma unroll
      for (int e = 0; e < kFpLd; ++e) {
        const int c = c0 + kk0 + 4 * e;
        rr[e] = (c < C && qrow >= 0) ? __ldg(rb + (size_t)c * N + qrow) : 0.f;
        xr[e] = (c < C && pok) ? __ldg(xb + (size_t)c * N + pcol) : 0.f;
      }
    };
    fetch(0);
    for (int c0 = 0; c0 < C; c0 += kFpKc) {
#pragma unroll
      for (int e = 0; e < kFpLd; ++e) {
        Rs[kk0 + 4 * e][jj] = rr[e];
        // Xn = fl(X * inv_norm): the reference normalises the patch first (NPS:40), then correlates
        Xs[kk0 + 4 * e][jj] = __fmul_rn(xr[e], invp);
      }
      __syncthreads();
      if (c0 + kFpKc < C) fetch(c0 + kFpKc);
#pragma unroll
      for (int kk = 0; kk < kFpKc; ++kk) {
        const float4 rv = *reinterpret_cast<const float4*>(&Rs[kk][ty * 4]);
        const float4 xv = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
        const float r4[4] = {rv.x, rv.y, rv.z, rv.w};
        const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(r4[i], x4[j], acc[i][j]);
      }
      __syncthreads();
    }

    // per row: best of this thread's 4 columns, then across the 16 threads sharing the row
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      long long key = kPackedIdentity;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = col0 + tx * 4 + j;
        if (p < col_end) {
          const long long k2 = pack_maxidx(acc[i][j], p);
          key = k2 > key ? k2 : key;
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
      }
      const int q = rows[ty * 4 + i];
      if (tx == 0 && q >= 0 && key != kPackedIdentity) atomicMax(packed + (size_t)b * N + q, key);
    }
  }
}

// Sparse variant for the recheck lists of the tensor mode (a handful of rows per image): the dense
// kernel above would multiply 64-row tiles that are 95 % padding.  grid = (column tiles of 64, B);
// 256 threads = 64 bank columns x 4 channel quarters; rows are taken 8 at a time (R rows staged in
// shared memory, broadcast reads), the four quarter sums are added in fixed order.
constexpr int kSkCols = 64;
constexpr int kSkRows = 8;

// one CTA: the listed rows of image b against bank columns [col_begin + tile * 64, + 64)
__device__ __forceinline__ void
sparse_cta(float* sk_smem, int* rows, int tile, int b, int nrows, const float* __restrict__ x, const float* __restrict__ ref,
           const float* __restrict__ inv_norm, int C, int N, int col_begin, int col_end, const int* __restrict__ list,
           long long* __restrict__ packed) {
  float* Rs = sk_smem;                                   // [C][8]
  float* red = sk_smem + (size_t)C * kSkRows;            // [4][64][8]
  const int j = threadIdx.x & 63, kg = threadIdx.x >> 6;
  const int p = col_begin + tile * kSkCols + j;
  const bool pok = p < col_end;
  const float* xb = x + (size_t)b * C * N;
  const float* rb = ref + (size_t)b * C * N;
  const float invp = pok ? inv_norm[(size_t)b * N + p] : 0.f;
  const int cq = (C + 3) / 4;                            // channels per quarter
  const int cbeg = kg * cq, cend = min(C, cbeg + cq);

  for (int r0 = 0; r0 < nrows; r0 += kSkRows) {
    __syncthreads();
    if (threadIdx.x < kSkRows) rows[threadIdx.x] = (r0 + threadIdx.x < nrows) ? list[(size_t)b * N + r0 + threadIdx.x] : -1;
    __syncthreads();
    for (int idx = threadIdx.x; idx < C * kSkRows; idx += blockDim.x) {
      const int c = idx >> 3, i = idx & 7;
      const int q = rows[i];
      Rs[idx] = (q >= 0) ? __ldg(rb + (size_t)c * N + q) : 0.f;
    }
    __syncthreads();
    float acc[kSkRows];
#pragma unroll
    for (int i = 0; i < kSkRows; ++i) acc[i] = 0.f;
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
      float xv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) xv[u] = (pok && c0 + u < cend) ? __ldg(xb + (size_t)(c0 + u) * N + p) : 0.f;
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        if (c0 + u < cend) {
          const float xn = __fmul_rn(xv[u], invp);       // Xn = fl(X * inv_norm)   (NPS:40)
          const float4 ra = *reinterpret_cast<const float4*>(Rs + (size_t)(c0 + u) * kSkRows);
          const float4 rc = *reinterpret_cast<const float4*>(Rs + (size_t)(c0 + u) * kSkRows + 4);
          acc[0] = fmaf(ra.x, xn, acc[0]); acc[1] = fmaf(ra.y, xn, acc[1]);
          acc[2] = fmaf(ra.z, xn, acc[2]); acc[3] = fmaf(ra.w, xn, acc[3]);
          acc[4] = fmaf(rc.x, xn, acc[4]); acc[5] = fmaf(rc.y, xn, acc[5]);
          acc[6] = fmaf(rc.z, xn, acc[6]); acc[7] = fmaf(rc.w, xn, acc[7]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kSkRows; ++i) red[((size_t)kg * kSkCols + j) * kSkRows + i] = acc[i];
    __syncthreads();
    if (kg == 0) {
#pragma unroll
      for (int i = 0; i < kSkRows; ++i) {
        float sum = red[(size_t)j * kSkRows + i];
        sum += red[((size_t)1 * kSkCols + j) * kSkRows + i];
        sum += red[((size_t)2 * kSkCols + j) * kSkRows + i];
        sum += red[((size_t)3 * kSkCols + j) * kSkRows + i];
        long long key = pok ? pack_maxidx(sum, p) : kPackedIdentity;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const long long other = __shfl_xor_sync(0xffffffffu, key, o);
          key = other > key ? other : key;
        }
        const int q = rows[i];
        if ((threadIdx.x & 31) == 0 && q >= 0 && key != kPackedIdentity) atomicMax(packed + (size_t)b * N + q, key);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
corr_fp32_sparse_kernel(const float* __restrict__ x, const float* __restrict__ ref, const float* __restrict__ inv_norm,
                        int C, int N, int col_begin, int col_end,
                        const int* __restrict__ list, const int* __restrict__ nlist, long long* __restrict__ packed) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float sk_smem[];
  __shared__ int rows[kSkRows];
  const int b = blockIdx.y;
  const int nrows = min(nlist[b], N);
  if (nrows == 0) return;
  sparse_cta(sk_smem, rows, blockIdx.x, b, nrows, x, ref, inv_norm, C, N, col_begin, col_end, list, packed);
}

__global__ void select_all_kernel(int N, int* __restrict__ list, int* __restrict__ nlist, long long* __restrict__ packed) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < N) {
    list[(size_t)b * N + q] = q;
    packed[(size_t)b * N + q] = kPackedIdentity;
  }
  if (q == 0) nlist[b] = N;
}

__global__ void apply_recheck_kernel(const long long* __restrict__ packed, const int* __restrict__ list,
                                     const int* __restrict__ nlist, int N, int* __restrict__ ind, float* __restrict__ vmax) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= min(nlist[b], N)) return;
  const int q = list[(size_t)b * N + i];
  float v;
  int idx;
  unpack_maxidx(packed[(size_t)b * N + q], &v, &idx);
  ind[(size_t)b * N + q] = idx;
  if (vmax) vmax[(size_t)b * N + q] = v;
}

// rows of `pair_list` (exactly two candidates inside the error band): one warp per row computes both exact fp32 scores
__device__ __forceinline__ void
resolve_pairs_cta(int part, int nparts, int b, const int* __restrict__ pair_list, const int* __restrict__ npair,
                  const int* __restrict__ cand2, const float* __restrict__ xt, const float* __restrict__ ref,
                  const float* __restrict__ inv_norm, int C, int N, int* __restrict__ ind, float* __restrict__ vmax) {
  const int np = min(npair[b], N);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = part * 8 + warp; j < np; j += nparts * 8) {
    const int q = pair_list[(size_t)b * N + j];
    const int p1 = ind[(size_t)b * N + q], p2 = cand2[(size_t)b * N + q];
    const float i1 = inv_norm[(size_t)b * N + p1], i2 = inv_norm[(size_t)b * N + p2];
    const float* x1 = xt + ((size_t)b * N + p1) * C;
    const float* x2 = xt + ((size_t)b * N + p2) * C;
    const float* rr = ref + (size_t)b * C * N + q;
    float a1 = 0.f, a2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float r = __ldg(rr + (size_t)c * N);
      a1 = fmaf(r, __fmul_rn(__ldg(x1 + c), i1), a1);      // Xn = fl(X * inv_norm)   (NPS:40)
      a2 = fmaf(r, __fmul_rn(__ldg(x2 + c), i2), a2);
    }
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (lane == 0) {
      const bool second_wins = (a2 > a1) || (a2 == a1 && p2 < p1);
      ind[(size_t)b * N + q] = second_wins ? p2 : p1;
      if (vmax) vmax[(size_t)b * N + q] = second_wins ? a2 : a1;
    }
  }
}


// Settles what the tensor passes left open, in one launch:
//   (1) rows of `list` (recomputed in exact fp32 by ipsr_correlate_argmax_fp32): ind[b,q] <- packed[b,q];
//   (2) rows of `pair_list` (exactly two candidates, ind[b,q] and cand2[b,q], inside the error band of the
//       three-pass split): one warp computes both exact fp32 scores <R[q], fl(X[p] inv_norm[p])> and keeps the
//       larger one (the lower column on a tie, torch.max).
// grid = (G, B), 256 threads.
__global__ void __launch_bounds__(256)
resolve_kernel(const long long* __restrict__ packed, const int* __restrict__ list, const int* __restrict__ nlist,
               const int* __restrict__ pair_list, const int* __restrict__ npair, const int* __restrict__ cand2,
               const float* __restrict__ xt, const float* __restrict__ ref, const float* __restrict__ inv_norm,
               int C, int N, int* __restrict__ ind, float* __restrict__ vmax, const int* __restrict__ npass2,
               int* __restrict__ nrecheck_out, int* __restrict__ npass2_out) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  if (blockIdx.x == 0 && threadIdx.x == 0) {               // the counters the caller keeps (diagnostics): no copy nodes
    if (nrecheck_out) nrecheck_out[b] = nlist[b];
    if (npass2_out && npass2) npass2_out[b] = npass2[b];
  }
  const int nl = min(nlist[b], N);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += gridDim.x * blockDim.x) {
    const int q = list[(size_t)b * N + i];
    float v;
    int idx;
    unpack_maxidx(packed[(size_t)b * N + q], &v, &idx);
    ind[(size_t)b * N + q] = idx;
    if (vmax) vmax[(size_t)b * N + q] = v;
  }
  if (!pair_list) return;
  resolve_pairs_cta(blockIdx.x, gridDim.x, b, pair_list, npair, cand2, xt, ref, inv_norm, C, N, ind, vmax);
}

// The exact fp32 recheck of the listed rows over the WHOLE bank and everything resolve_kernel does, as one launch
// (ipsr_shift_forward, tensor mode): grid = (N / 64 column tiles + gpairs, B).  The pair CTAs need nothing from the
// recheck; the listed rows take their keys from the image's LAST column tile to finish (ticket in done[b], zero on entry).
__global__ void __launch_bounds__(256)
recheck_resolve_kernel(const float* __restrict__ x, const float* __restrict__ ref, const float* __restrict__ inv_norm,
                       const float* __restrict__ xt, int C, int N, const int* __restrict__ list, const int* __restrict__ nlist,
                       long long* __restrict__ packed, const int* __restrict__ pair_list, const int* __restrict__ npair,
                       const int* __restrict__ cand2, int* __restrict__ ind, int* __restrict__ done,
                       const int* __restrict__ npass2, int* __restrict__ nrecheck_out, int* __restrict__ npass2_out,
                       int ntiles, int gpairs) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float sk_smem[];
  __shared__ int rows[kSkRows];
  __shared__ int last;
  const int b = blockIdx.y;
  if ((int)blockIdx.x >= ntiles) {
    if ((int)blockIdx.x == ntiles && threadIdx.x == 0) {
      if (nrecheck_out) nrecheck_out[b] = nlist[b];
      if (npass2_out && npass2) npass2_out[b] = npass2[b];
    }
    resolve_pairs_cta((int)blockIdx.x - ntiles, gpairs, b, pair_list, npair, cand2, xt, ref, inv_norm, C, N, ind, nullptr);
    return;
  }
  const int nrows = min(nlist[b], N);
  if (nrows == 0) return;
  sparse_cta(sk_smem, rows, blockIdx.x, b, nrows, x, ref, inv_norm, C, N, 0, N, list, packed);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                                       // this CTA's keys before its ticket
    last = (atomicAdd(done + b, 1) == ntiles - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
    const int q = list[(size_t)b * N + i];
    float v;
    int idx;
    unpack_maxidx(__ldcg(packed + (size_t)b * N + q), &v, &idx);
    ind[(size_t)b * N + q] = idx;
  }
}

// one warp per (b, q): exact score of the already chosen winner, as an exchange key.
__global__ void __launch_bounds__(256)
pack_winner_kernel(const float* __restrict__ xt, const float* __restrict__ ref, const float* __restrict__ inv_norm,
                   const int* __restrict__ ind, int C, int N, long long* __restrict__ packed) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= N) return;
  long long* slot = packed + (size_t)b * N + q;
  if (*slot != kPackedIdentity) return;
  const int p = ind[(size_t)b * N + q];
  const float inv = inv_norm[(size_t)b * N + p];
  const float* xr = xt + ((size_t)b * N + p) * C;
  const float* rr = ref + (size_t)b * C * N + q;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(rr + (size_t)c * N), __fmul_rn(__ldg(xr + c), inv), acc);
  acc = warp_sum(acc);
  if (lane == 0) *slot = pack_maxidx(acc, p);
}

__global__ void pack_kernel(const float* __restrict__ v, const int* __restrict__ idx, long long n, long long* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = pack_maxidx(v[i], idx[i]);
}
__global__ void unpack_kernel(const long long* __restrict__ in, long long n, float* __restrict__ v, int* __restrict__ idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float vv;
  int ii;
  unpack_maxidx(in[i], &vv, &ii);
  if (v) v[i] = vv;
  if (idx) idx[i] = ii;
}

// MaxCoord on a materialised [P, L] score tensor: thread per location, coalesced over l.
__global__ void maxcoord_kernel(const float* __restrict__ s, int P, int L, long long* __restrict__ ind, float* __restrict__ vmax) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  long long key = kPackedIdentity;
  for (int p = 0; p < P; ++p) {
    const long long k2 = pack_maxidx(__ldg(s + (size_t)p * L + l), p);
    key = k2 > key ? k2 : key;
  }
  float v;
  int idx;
  unpack_maxidx(key, &v, &idx);
  // the key canonicalises -0 -> +0 and NaN payloads; report the stored element itself
  ind[l] = idx;
  vmax[l] = s[(size_t)idx * L + l];
}

}  // namespace ipsr

extern "C" int ipsr_select_all_rows(int B, int N, int32_t* recheck_list, int32_t* nrecheck, int64_t* packed, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(recheck_list && nrecheck && packed && B > 0 && N > 0, IPSR_ERR_INVALID_ARG, "ipsr_select_all_rows: bad arguments");
  select_all_kernel<<<dim3((N + 255) / 256, B), 256, 0, as_stream(stream)>>>(N, recheck_list, nrecheck,
                                                                             reinterpret_cast<long long*>(packed));
  return check_launch("ipsr_select_all_rows");
}

extern "C" int ipsr_correlate_argmax_fp32(const float* x, const float* ref, const float* inv_norm,
                                          int B, int C, int N, int col_begin, int col_end,
                                          const int32_t* recheck_list, const int32_t* nrecheck, int row_ctas,
                                          int64_t* packed, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(x && ref && inv_norm && recheck_list && nrecheck && packed, IPSR_ERR_INVALID_ARG,
               "ipsr_correlate_argmax_fp32: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0 && col_begin >= 0 && col_end <= N && col_begin < col_end, IPSR_ERR_INVALID_ARG,
               "ipsr_correlate_argmax_fp32: bad dims B=%d C=%d N=%d cols=[%d,%d)", B, C, N, col_begin, col_end);
  IPSR_REQUIRE(B <= 65535, IPSR_ERR_UNSUPPORTED, "ipsr_correlate_argmax_fp32: B=%d > 65535", B);
  if (row_ctas <= 0) {
    // sparse row lists (tensor-mode recheck)
    const size_t smem = ((size_t)C * kSkRows + 4 * (size_t)kSkCols * kSkRows) * sizeof(float);
    IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_correlate_argmax_fp32: C=%d too large", C);
    if (smem + 2048 > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(corr_fp32_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "corr_fp32_sparse smem attribute: %s", cudaGetErrorString(e));
    }
    dim3 grid((col_end - col_begin + kSkCols - 1) / kSkCols, B);
    {
      cudaError_t le__ = launch_pdl(corr_fp32_sparse_kernel, grid, dim3(256), smem, as_stream(stream), x, ref, inv_norm, C, N, col_begin,
                                    col_end, recheck_list, nrecheck, reinterpret_cast<long long*>(packed));
      IPSR_REQUIRE(le__ == cudaSuccess, IPSR_ERR_CUDA, "ipsr_correlate_argmax_fp32: launch failed: %s", cudaGetErrorString(le__));
    }
    return check_launch("ipsr_correlate_argmax_fp32");
  }
  const int max_ctas = (N + kFpTile - 1) / kFpTile;
  if (row_ctas > max_ctas) row_ctas = max_ctas;
  dim3 grid((col_end - col_begin + kFpTile - 1) / kFpTile, row_ctas, B);
  corr_fp32_kernel<<<grid, kFpThreads, 0, as_stream(stream)>>>(x, ref, inv_norm, C, N, col_begin, col_end,
                                                                recheck_list, nrecheck,
                                                                reinterpret_cast<long long*>(packed));
  return check_launch("ipsr_correlate_argmax_fp32");
}

int ipsr::recheck_resolve_ex(const float* x, const float* ref, const float* inv_norm, const float* xt, int B, int C, int N,
                             const int32_t* list, const int32_t* nlist, int64_t* packed, const int32_t* pair_list,
                             const int32_t* npair, const int32_t* cand2, int32_t* ind, int32_t* done, const int32_t* npass2,
                             int32_t* nrecheck_out, int32_t* npass2_out, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(x && ref && inv_norm && xt && list && nlist && packed && pair_list && npair && cand2 && ind && done,
               IPSR_ERR_INVALID_ARG, "recheck_resolve: null pointer");
  IPSR_REQUIRE(B > 0 && B <= 65535 && C > 0 && N > 0, IPSR_ERR_INVALID_ARG, "recheck_resolve: bad dims");
  const size_t smem = ((size_t)C * kSkRows + 4 * (size_t)kSkCols * kSkRows) * sizeof(float);
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "recheck_resolve: C=%d too large", C);
  if (smem + 2048 > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(recheck_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "recheck_resolve smem attribute: %s", cudaGetErrorString(e));
  }
  const int ntiles = (N + kSkCols - 1) / kSkCols, gpairs = 4;
  cudaError_t le = launch_pdl(recheck_resolve_kernel, dim3(ntiles + gpairs, B), dim3(256), smem, as_stream(stream), x, ref, inv_norm, xt, C, N,
                              list, nlist, reinterpret_cast<long long*>(packed), pair_list, npair, cand2, ind, done, npass2, nrecheck_out,
                              npass2_out, ntiles, gpairs);
  IPSR_REQUIRE(le == cudaSuccess, IPSR_ERR_CUDA, "recheck_resolve: launch failed: %s", cudaGetErrorString(le));
  return check_launch("recheck_resolve");
}

extern "C" int ipsr_apply_recheck(const int64_t* packed, const int32_t* recheck_list, const int32_t* nrecheck,
                                  int B, int N, int32_t* ind, float* vmax, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(packed && recheck_list && nrecheck && ind && B > 0 && N > 0, IPSR_ERR_INVALID_ARG,
               "ipsr_apply_recheck: bad arguments");
  apply_recheck_kernel<<<dim3((N + 255) / 256, B), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const long long*>(packed), recheck_list, nrecheck, N, ind, vmax);
  return check_launch("ipsr_apply_recheck");
}

extern "C" int ipsr_resolve_rows(const int64_t* packed, const int32_t* recheck_list, const int32_t* nrecheck,
                                 const int32_t* pair_list, const int32_t* npair, const int32_t* cand2,
                                 const float* xt, const float* ref, const float* inv_norm,
                                 int B, int C, int N, int32_t* ind, float* vmax, void* stream) {
  return ipsr::resolve_rows_ex(packed, recheck_list, nrecheck, pair_list, npair, cand2, xt, ref, inv_norm, B, C, N, ind, vmax, nullptr,
                               nullptr, nullptr, stream);
}

int ipsr::resolve_rows_ex(const int64_t* packed, const int32_t* recheck_list, const int32_t* nrecheck, const int32_t* pair_list,
                          const int32_t* npair, const int32_t* cand2, const float* xt, const float* ref, const float* inv_norm,
                          int B, int C, int N, int32_t* ind, float* vmax, const int32_t* npass2, int32_t* nrecheck_out,
                          int32_t* npass2_out, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(packed && recheck_list && nrecheck && ind && B > 0 && N > 0 && C > 0 && B <= 65535, IPSR_ERR_INVALID_ARG,
               "ipsr_resolve_rows: bad arguments");
  IPSR_REQUIRE(!pair_list || (npair && cand2 && xt && ref && inv_norm), IPSR_ERR_INVALID_ARG,
               "ipsr_resolve_rows: the pair list needs npair, cand2, xt, ref and inv_norm");
  int G = (N + 255) / 256;
  if (G > 8) G = 8;
  {
    cudaError_t le__ = launch_pdl(resolve_kernel, dim3(G, B), dim3(256), 0, as_stream(stream), reinterpret_cast<const long long*>(packed),
                                  recheck_list, nrecheck, pair_list, npair, cand2, xt, ref, inv_norm, C, N, ind, vmax, npass2, nrecheck_out,
                                  npass2_out);
    IPSR_REQUIRE(le__ == cudaSuccess, IPSR_ERR_CUDA, "ipsr_resolve_rows: launch failed: %s", cudaGetErrorString(le__));
  }
  return check_launch("ipsr_resolve_rows");
}

extern "C" int ipsr_pack_winner_scores(const float* xt, const float* ref, const float* inv_norm, const int32_t* ind,
                                       int B, int C, int N, int64_t* packed, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(xt && ref && inv_norm && ind && packed && B > 0 && C > 0 && N > 0, IPSR_ERR_INVALID_ARG,
               "ipsr_pack_winner_scores: bad arguments");
  pack_winner_kernel<<<dim3((N + 7) / 8, B), 256, 0, as_stream(stream)>>>(xt, ref, inv_norm, ind, C, N,
                                                                          reinterpret_cast<long long*>(packed));
  return check_launch("ipsr_pack_winner_scores");
}

extern "C" int ipsr_pack_maxidx(const float* v, const int32_t* idx, int64_t n, int64_t* packed, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(v && idx && packed && n >= 0, IPSR_ERR_INVALID_ARG, "ipsr_pack_maxidx: bad arguments");
  if (n == 0) return IPSR_OK;
  pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(v, idx, n, reinterpret_cast<long long*>(packed));
  return check_launch("ipsr_pack_maxidx");
}

extern "C" int ipsr_unpack_maxidx(const int64_t* packed, int64_t n, float* v, int32_t* idx, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(packed && n >= 0, IPSR_ERR_INVALID_ARG, "ipsr_unpack_maxidx: bad arguments");
  if (n == 0) return IPSR_OK;
  unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(packed), n, v, idx);
  return check_launch("ipsr_unpack_maxidx");
}

extern "C" int ipsr_maxcoord(const float* s, int P, int L, int64_t* ind_i64, float* vmax, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(s && ind_i64 && vmax && P > 0 && L > 0, IPSR_ERR_INVALID_ARG, "ipsr_maxcoord: bad arguments");
  maxcoord_kernel<<<(L + 127) / 128, 128, 0, as_stream(stream)>>>(s, P, L, reinterpret_cast<long long*>(ind_i64), vmax);
  return check_launch("ipsr_maxcoord");
}
