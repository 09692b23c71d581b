// shift_sz = k > 1 / stride = s > 1 ("patch mode"), FORWARD ONLY: the reference computes the whole output for these
// settings (models/IPSRFunction.py:46-133) and then fails storing the attention for its backward (:134).
//
// A k x k patch over C channels is a row of K = C*k*k values in the reference's unfold order (c, dy, dx)
// (util/NonparametricShift.py:65-68).  Two routes:
//   K <= 1024: unfold x and ref into patch maps [B][Kpad][nH*nW] -- "feature maps" with Kpad channels on the nH x nW
//     grid of patch positions -- and run the 1 x 1 pipeline (ipsr_shift_forward, tensor path included) on them
//     unchanged; fold_cols_kernel then sums the overlapping patches (ConvTranspose2d, IPSRFunction.py:131).
//   K  > 1024 (e.g. C = 256, k = 3: K = 2304): the shared-memory tiles of the 1 x 1 prep / blend kernels no longer
//     fit; patch_rows_kernel writes position-major rows + norms, the exact fp32 correlation runs on the patch maps,
//     blend_wide_kernel runs the recurrence with y in registers (one CTA per image) and fold_rows_kernel gathers
//     and sums.
#include <stdlib.h>

#include "ipsr_common.cuh"

namespace ipsr {

struct PatchGeom {
  int C, H, W, k, s, nH, nW, K, Kpad;
};

// cols[b][kk][q] = x[b][c][i*s+dy][j*s+dx], kk = (c*k + dy)*k + dx, q = i*nW + j; rows kk >= K are zero
__global__ void __launch_bounds__(256) unfold_cols_kernel(const float* __restrict__ x, PatchGeom g, long long total,
                                                          float* __restrict__ cols) {
  const int P = g.nH * g.nW, kk2 = g.k * g.k;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t % P);
    const long long r = t / P;
    const int kk = (int)(r % g.Kpad);
    const int b = (int)(r / g.Kpad);
    float v = 0.f;
    if (kk < g.K) {
      const int c = kk / kk2, d = kk - c * kk2, dy = d / g.k, dx = d - dy * g.k;
      const int i = q / g.nW, j = q - i * g.nW;
      v = __ldg(x + (((size_t)b * g.C + c) * g.H + (i * g.s + dy)) * g.W + (j * g.s + dx));
    }
    cols[t] = v;
  }
}

// out[b][c][Y][X] = sum over the patches (i, j) covering (Y, X) of cols[b][(c, Y - i*s, X - j*s)][i*nW + j]
__global__ void __launch_bounds__(256) fold_cols_kernel(const float* __restrict__ cols, PatchGeom g, long long total,
                                                        float* __restrict__ out) {
  const int P = g.nH * g.nW, kk2 = g.k * g.k;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(t % g.W);
    long long r = t / g.W;
    const int Y = (int)(r % g.H);
    r /= g.H;
    const int c = (int)(r % g.C);
    const int b = (int)(r / g.C);
    const float* src = cols + ((size_t)b * g.Kpad + (size_t)c * kk2) * P;
    float acc = 0.f;
    for (int dy = 0; dy < g.k; ++dy) {
      const int yy = Y - dy;
      if (yy < 0 || yy % g.s != 0 || yy / g.s >= g.nH) continue;
      for (int dx = 0; dx < g.k; ++dx) {
        const int xx = X - dx;
        if (xx < 0 || xx % g.s != 0 || xx / g.s >= g.nW) continue;
        acc += __ldg(src + (size_t)(dy * g.k + dx) * P + (yy / g.s) * g.nW + xx / g.s);
      }
    }
    out[t] = acc;
  }
}

// rows[b][q][kk] (position-major raw patches = decoder weights, NonparametricShift.py:54) and
// inv_norm[b][q] = 1 / (||patch||_2 + 1e-8) (:40).  One CTA per patch position.
__global__ void __launch_bounds__(256) patch_rows_kernel(const float* __restrict__ x, PatchGeom g, float* __restrict__ rows,
                                                         float* __restrict__ inv_norm, float* __restrict__ norm_out,
                                                         float* __restrict__ max_out) {
  __shared__ float part[8], partm[8];
  const int q = blockIdx.x, b = blockIdx.y, P = g.nH * g.nW, kk2 = g.k * g.k;
  const int i = q / g.nW, j = q - i * g.nW;
  const float* xb = x + (size_t)b * g.C * g.H * g.W + (size_t)(i * g.s) * g.W + j * g.s;
  float* dst = rows + ((size_t)b * P + q) * g.K;
  float ss = 0.f, mx = 0.f;
  for (int kk = threadIdx.x; kk < g.K; kk += blockDim.x) {
    const int c = kk / kk2, d = kk - c * kk2, dy = d / g.k, dx = d - dy * g.k;
    const float v = __ldg(xb + ((size_t)c * g.H + dy) * g.W + dx);
    dst[kk] = v;
    ss = fmaf(v, v, ss);
    mx = fmaxf(mx, fabsf(v));
  }
  if (!inv_norm && !norm_out && !max_out) return;
  ss = warp_sum(ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) {
    part[threadIdx.x >> 5] = ss;
    partm[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f, m = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      tot += part[w];
      m = fmaxf(m, partm[w]);
    }
    if (inv_norm) inv_norm[(size_t)b * P + q] = 1.0f / (sqrtf(tot) + 1e-8f);
    if (norm_out) norm_out[(size_t)b * P + q] = sqrtf(tot);
    if (max_out) max_out[(size_t)b * P + q] = m;
  }
}

// The coherent blend (IPSRFunction.py:82-126) on rows of K values, one CTA of 1024 threads per image, y_{l-1} in
// registers (E values per thread).  Step l: a = <u_l, y_{l-1}> (block reduction, fixed order), wn = a/(a+v),
// wo = v/(a+v), y_l = wn*y_{l-1} + wo*X[p_l]; the operand rows of step l+1 are loaded before the reduction of step l.
constexpr int kWideThreads = 1024;
template <int E>
__global__ void __launch_bounds__(kWideThreads) blend_wide_kernel(const float* __restrict__ rows, const float* __restrict__ inv_norm,
                                                                   const float* __restrict__ vmax, const int32_t* __restrict__ ind,
                                                                   const int32_t* __restrict__ mask_idx, int K, int P, int M,
                                                                   float* __restrict__ y, float* __restrict__ wn_out,
                                                                   float* __restrict__ wo_out) {
  __shared__ float part[2][32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* R = rows + (size_t)b * P * K;
  const int32_t* indb = ind + (size_t)b * P;
  float* yb = y + (size_t)b * M * K;
  float yv[E], un[E], kn[E];
  {
    const int p0 = indb[mask_idx[0]];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int kk = tid + e * kWideThreads;
      yv[e] = kk < K ? R[(size_t)p0 * K + kk] : 0.f;
      if (kk < K) yb[kk] = yv[e];
    }
    if (tid == 0) {
      wn_out[(size_t)b * M] = 0.f;
      wo_out[(size_t)b * M] = 1.f;
    }
  }
  auto load_step = [&](int l, float& invq, float& v) {
    const int q = mask_idx[l], p = indb[q];
    invq = inv_norm[(size_t)b * P + q];
    v = vmax[(size_t)b * P + q];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int kk = tid + e * kWideThreads;
      un[e] = kk < K ? R[(size_t)q * K + kk] : 0.f;
      kn[e] = kk < K ? R[(size_t)p * K + kk] : 0.f;
    }
  };
  float invq = 0.f, v = 0.f;
  if (M > 1) load_step(1, invq, v);
  for (int l = 1; l < M; ++l) {
    float acc = 0.f;
    float kc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      acc = fmaf(__fmul_rn(un[e], invq), yv[e], acc);   // u_l = little * (1/(norm + 1e-8)) (:109), then the dot (:116)
      kc[e] = kn[e];
    }
    const float vc = v;
    if (l + 1 < M) load_step(l + 1, invq, v);
    acc = warp_sum(acc);
    if (lane == 0) part[l & 1][warp] = acc;
    __syncthreads();
    const float a = warp_sum(part[l & 1][lane]);
    const float den = a + vc;
    const float wn = a / den, wo = vc / den;            // no clamping (:120-121)
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int kk = tid + e * kWideThreads;
      yv[e] = __fadd_rn(__fmul_rn(wn, yv[e]), __fmul_rn(wo, kc[e]));   // :122
      if (kk < K) yb[(size_t)l * K + kk] = yv[e];
    }
    if (tid == 0) {
      wn_out[(size_t)b * M + l] = wn;
      wo_out[(size_t)b * M + l] = wo;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Blocked version of the recurrence (the scheme of ipsr_blend.cu's scan, for rows too long for its shared-memory
// tiles): only ONE scalar per step is sequential.  By linearity z_i = <u_i, y_cur> for the steps i still ahead obeys
// z_i <- wn_l z_i + wo_l <u_i, X[p_l]>, so inside a block of T = 32 steps the chain runs on scalars and the in-block
// Gram matrix Gt[j][i] = <u_i, X[p_j]>, which wide_gram_kernel computes for all blocks in parallel.  Per block the scan
// CTA then needs one real reduction (z_i = <u_i, y_prev> for the block's T rows: re-anchors z on the real y, so that
// rounding differences to the reference's dot product cannot accumulate beyond T steps) and one channel-parallel update
// of y.  The rows of u and X[p] are read in place from `rows` (position-major): nothing is staged.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWideT = 32;
constexpr int kWideKS = 8;               // K ranges of the Gram kernel (partial matrices, summed in fixed order by the scan)
constexpr int kWideCluster = 8;          // CTAs per image of the scan: each owns K / 8 of every row
constexpr int kWideScanThreads = 256;

// grid = (blocks of T steps, kWideKS, B), 256 threads: thread (ti, tj) owns the 2 x 2 outputs (ti, ti+16) x (tj, tj+16) of
// the partial Gram matrix over its K range: gram[b][block][ks][j][i] = sum over that range of u_i[k] X[p_j][k]
__global__ void __launch_bounds__(256) wide_gram_kernel(const float* __restrict__ rows, const float* __restrict__ inv_norm,
                                                        const int32_t* __restrict__ ind, const int32_t* __restrict__ mask_idx,
                                                        int K, int P, int M, float* __restrict__ gram) {
  constexpr int T = kWideT, KC = 64;
  __shared__ float Us[T][KC + 4], Ks[T][KC + 4];
  __shared__ int qs[T], ps[T];
  __shared__ float invs[T];
  const int kb = blockIdx.x, ks = blockIdx.y, b = blockIdx.z, nblk = gridDim.x;
  const int l0 = kb * T;
  const float* R = rows + (size_t)b * P * K;
  if (threadIdx.x < T) {
    const int l = l0 + threadIdx.x;
    const int q = l < M ? mask_idx[l] : -1;
    qs[threadIdx.x] = q;
    ps[threadIdx.x] = q >= 0 ? ind[(size_t)b * P + q] : -1;
    invs[threadIdx.x] = q >= 0 ? inv_norm[(size_t)b * P + q] : 0.f;
  }
  __syncthreads();
  const int kper = ((K + kWideKS - 1) / kWideKS + KC - 1) / KC * KC;      // whole chunks per range
  const int kbeg = ks * kper, kend = min(K, kbeg + kper);
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  float g00 = 0.f, g01 = 0.f, g10 = 0.f, g11 = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += KC) {
    for (int it = threadIdx.x; it < T * KC; it += 256) {        // coalesced 256-byte row segments
      const int r = it / KC, c = it - r * KC;
      const bool ok = qs[r] >= 0 && k0 + c < kend;
      Us[r][c] = ok ? __fmul_rn(__ldg(R + (size_t)qs[r] * K + k0 + c), invs[r]) : 0.f;   // u = little * (1/(norm+1e-8))  :109
      Ks[r][c] = ok ? __ldg(R + (size_t)ps[r] * K + k0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < KC; ++c) {
      const float u0 = Us[ti][c], u1 = Us[ti + 16][c], x0 = Ks[tj][c], x1 = Ks[tj + 16][c];
      g00 = fmaf(u0, x0, g00);
      g01 = fmaf(u0, x1, g01);
      g10 = fmaf(u1, x0, g10);
      g11 = fmaf(u1, x1, g11);
    }
    __syncthreads();
  }
  float* G = gram + (((size_t)b * nblk + kb) * kWideKS + ks) * T * T;          // Gt[j][i] = <u_i, X[p_j]> over this range
  G[tj * T + ti] = g00;
  G[(tj + 16) * T + ti] = g01;
  G[tj * T + ti + 16] = g10;
  G[(tj + 16) * T + ti + 16] = g11;
}

__device__ __forceinline__ float wide_rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t wide_map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void wide_st_cluster(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// One CLUSTER of kWideCluster CTAs per image: a single SM cannot pull the 2 * M * K * 4 bytes of u / X[p] rows (21 MB at
// configs[3]) fast enough -- the one-CTA version of this kernel was bound by exactly that, at 1.1 us per step -- so every
// CTA owns 1/8 of the columns of every row.  Per block of T steps: (B) each CTA reduces its part of z_i = <u_i, y_prev>,
// writes the T partial sums into the shared memory of ALL CTAs of the cluster (distributed shared memory) and, after ONE
// cluster barrier, adds the eight partials in rank order; (C) every CTA runs the same T scalar steps on warp 0 (bit-identical
// in all of them); (D) each CTA updates its columns of y.  y lives in registers (E values per thread).
template <int E>
__global__ void __launch_bounds__(kWideScanThreads)
wide_scan_kernel(const float* __restrict__ rows, const float* __restrict__ inv_norm, const float* __restrict__ vmax,
                 const int32_t* __restrict__ ind, const int32_t* __restrict__ mask_idx, const float* __restrict__ gram,
                 int K, int P, int M, int kper, float* __restrict__ y, float* __restrict__ wn_out, float* __restrict__ wo_out) {
  constexpr int T = kWideT, NW = kWideScanThreads / 32;
  __shared__ float red[NW][T + 1];
  __shared__ float zpart[2][kWideCluster][T];              // [parity][source CTA][row]: written by every CTA of the cluster
  __shared__ float zs[T], wn_s[T], wo_s[T], vs[T], invs[T];
  __shared__ int qs[T], ps[T];
  const uint32_t crank = cluster_ctarank();
  const int b = blockIdx.x / kWideCluster, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* R = rows + (size_t)b * P * K;
  const int32_t* indb = ind + (size_t)b * P;
  float* yb = y + (size_t)b * M * K;
  const int nblk = (M + T - 1) / T;
  const int kbeg = (int)crank * kper, kend = min(K, kbeg + kper);      // this CTA's columns
  float yv[E];
#pragma unroll
  for (int e = 0; e < E; ++e) yv[e] = 0.f;
  cluster_sync_all();                                      // every CTA of the cluster runs before anybody writes into it
  for (int kb = 0; kb < nblk; ++kb) {
    const int l0 = kb * T;
    const int nvalid = min(T, M - l0);
    __syncthreads();                                       // the previous block's lists are no longer read
    if (tid < T) {
      const int l = l0 + tid;
      const int q = l < M ? mask_idx[l] : -1;
      qs[tid] = q;
      ps[tid] = q >= 0 ? indb[q] : -1;
      invs[tid] = q >= 0 ? inv_norm[(size_t)b * P + q] : 0.f;
      vs[tid] = q >= 0 ? vmax[(size_t)b * P + q] : 1.f;
    }
    __syncthreads();
    // ---- B: this CTA's part of z_i = <u_i, y_prev> for the block's rows, eight rows at a time ----
    for (int i0 = 0; i0 < T; i0 += 8) {
      float acc[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        acc[r] = 0.f;
        const int q = qs[i0 + r];
        if (q >= 0) {
          const float inv = invs[i0 + r];
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int kk = kbeg + tid + e * kWideScanThreads;
            if (kk < kend) acc[r] = fmaf(__fmul_rn(__ldg(R + (size_t)q * K + kk), inv), yv[e], acc[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = warp_sum(acc[r]);
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) red[warp][i0 + r] = acc[r];
      }
    }
    __syncthreads();
    if (tid < T) {
      float t = 0.f;
      for (int w = 0; w < NW; ++w) t += red[w][tid];        // fixed order
      const uint32_t mine = smem_u32(&zpart[kb & 1][crank][tid]);
#pragma unroll
      for (uint32_t dst = 0; dst < (uint32_t)kWideCluster; ++dst) wide_st_cluster(wide_map_to_cta(mine, dst), t);
    }
    cluster_sync_all();                                    // release / acquire: all partials of this block have landed
    // ---- C: T scalar steps on warp 0 (the same arithmetic in every CTA of the cluster) ----
    if (warp == 0) {
      float z = 0.f;
#pragma unroll
      for (int c = 0; c < kWideCluster; ++c) z += zpart[kb & 1][c][lane];      // rank order
      const float* Gt = gram + ((size_t)b * nblk + kb) * kWideKS * T * T;
      const float v = vs[lane];
      float my_wn = 0.f, my_wo = 1.f;
      float g[T], vj[T];
#pragma unroll
      for (int j = 0; j < T; ++j) {
        float gs = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < kWideKS; ++s2) gs += __ldg(Gt + (size_t)s2 * T * T + j * T + lane);   // range order
        g[j] = gs;
        vj[j] = __shfl_sync(0xffffffffu, v, j);
      }
#pragma unroll
      for (int j = 0; j < T; ++j) {
        const float zj = __shfl_sync(0xffffffffu, z, j);
        const float r = wide_rcp_approx(__fadd_rn(zj, vj[j]));   // no clamp: inf / nan propagate      :120
        float wn = __fmul_rn(zj, r);
        float wo = __fmul_rn(vj[j], r);                          //                                     :121
        if (kb == 0 && j == 0) {                                 // first masked patch: plain copy      :98-101
          wn = 0.f;
          wo = 1.f;
        }
        z = fmaf(wn, z, __fmul_rn(wo, g[j]));
        if (lane == j) {
          my_wn = wn;
          my_wo = wo;
        }
      }
      wn_s[lane] = my_wn;
      wo_s[lane] = my_wo;
      if (crank == 0 && l0 + lane < M) {
        wn_out[(size_t)b * M + l0 + lane] = my_wn;
        wo_out[(size_t)b * M + l0 + lane] = my_wo;
      }
    }
    __syncthreads();
    // ---- D: this CTA's columns of the block's y rows ----
    for (int j0 = 0; j0 < nvalid; j0 += 8) {
      float kv[8][E];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int p = j0 + r < nvalid ? ps[j0 + r] : -1;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int kk = kbeg + tid + e * kWideScanThreads;
          kv[r][e] = (p >= 0 && kk < kend) ? __ldg(R + (size_t)p * K + kk) : 0.f;
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (j0 + r < nvalid) {
          const float wn = wn_s[j0 + r], wo = wo_s[j0 + r];
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int kk = kbeg + tid + e * kWideScanThreads;
            yv[e] = __fadd_rn(__fmul_rn(wn, yv[e]), __fmul_rn(wo, kv[r][e]));          // :122
            if (kk < kend) yb[(size_t)(l0 + j0 + r) * K + kk] = yv[e];
          }
        }
      }
    }
  }
  cluster_sync_all();                                      // nobody leaves while a peer may still write into it
}

// ---------------------------------------------------------------------------------------------------------------
// The same cluster scan with everything a block needs ALREADY IN SHARED MEMORY when the block starts: the version above
// pays eight dependent round trips to L2 per block (index lists, u rows, Gram entries, X[p] rows: ~15 us per block of 32
// steps, latency not bandwidth).  Here the warps that idle during the scalar steps of block k fetch block k+1 with
// cp.async (this CTA's column slice of the 2 T rows, the index lists of block k+2, the Gram matrix of block k+1, summed
// over its K ranges), so that a block costs its own arithmetic, one cluster barrier and T scalar steps.
// Needs 4 * T * kper floats of shared memory (K <= ~3000 with 8 CTAs); longer rows take the kernel above.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int E>
__global__ void __launch_bounds__(kWideScanThreads)
wide_scan_smem_kernel(const float* __restrict__ rows, const float* __restrict__ inv_norm, const float* __restrict__ vmax,
                      const int32_t* __restrict__ ind, const int32_t* __restrict__ mask_idx, const float* __restrict__ gram,
                      int K, int P, int M, int kper, float* __restrict__ y, float* __restrict__ wn_out, float* __restrict__ wo_out) {
  constexpr int T = kWideT, NW = kWideScanThreads / 32;
  extern __shared__ __align__(16) float wsm[];              // U[2][T][kper] | X[2][T][kper] | G[2][T][T]
  float* Ubuf = wsm;
  float* Xbuf = Ubuf + 2 * (size_t)T * kper;
  float* Gbuf = Xbuf + 2 * (size_t)T * kper;
  __shared__ float red[NW][T + 1];
  __shared__ float zpart[2][kWideCluster][T];
  __shared__ float wn_s[T], wo_s[T];
  __shared__ float vs[3][T], invs[3][T];                    // index lists of three consecutive blocks
  __shared__ int qs[3][T], ps[3][T];
  const uint32_t crank = cluster_ctarank();
  const int b = blockIdx.x / kWideCluster, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* R = rows + (size_t)b * P * K;
  const int32_t* indb = ind + (size_t)b * P;
  float* yb = y + (size_t)b * M * K;
  const int nblk = (M + T - 1) / T;
  const int kbeg = (int)crank * kper, kend = min(K, kbeg + kper);
  const int klen = max(0, kend - kbeg);
  const bool vec16 = ((K & 3) == 0) && ((kper & 3) == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0);

  auto fetch_lists = [&](int kb) {                          // by one warp: T lanes, three dependent loads
    const int s = kb % 3, l = kb * T + lane;
    const int q = (kb < nblk && l < M) ? mask_idx[l] : -1;
    qs[s][lane] = q;
    ps[s][lane] = q >= 0 ? indb[q] : -1;
    invs[s][lane] = q >= 0 ? inv_norm[(size_t)b * P + q] : 0.f;
    vs[s][lane] = q >= 0 ? vmax[(size_t)b * P + q] : 1.f;
  };
  auto fetch_rows = [&](int kb, int t0, int nt) {           // threads t0 .. t0+nt-1: this CTA's slice of the block's 2 T rows
    if (kb >= nblk) return;
    const int s = kb % 3, buf = kb & 1;
    float* U = Ubuf + (size_t)buf * T * kper;
    float* X = Xbuf + (size_t)buf * T * kper;
    const int t = tid - t0;
    if (vec16) {
      const int per_row = klen >> 2;                        // 16-byte pieces per row slice
      for (int it = t; it < 2 * T * per_row; it += nt) {
        const int r = it / per_row, c = (it - r * per_row) << 2;
        const int rr = r < T ? r : r - T;
        const int src = r < T ? qs[s][rr] : ps[s][rr];
        float* dst = (r < T ? U : X) + (size_t)rr * kper + c;
        if (src >= 0) cp_async16(dst, R + (size_t)src * K + kbeg + c);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      for (int it = t; it < 2 * T * klen; it += nt) {
        const int r = it / klen, c = it - r * klen;
        const int rr = r < T ? r : r - T;
        const int src = r < T ? qs[s][rr] : ps[s][rr];
        float* dst = (r < T ? U : X) + (size_t)rr * kper + c;
        if (src >= 0) cp_async4(dst, R + (size_t)src * K + kbeg + c);
        else *dst = 0.f;
      }
    }
  };
  auto fetch_gram = [&](int kb, int t0, int nt) {           // Gram of block kb, its K ranges added in range order
    if (kb >= nblk) return;
    const float* Gt = gram + ((size_t)b * nblk + kb) * kWideKS * T * T;
    float* G = Gbuf + (size_t)(kb & 1) * T * T;
    for (int e = tid - t0; e < T * T; e += nt) {
      float gs = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < kWideKS; ++s2) gs += __ldg(Gt + (size_t)s2 * T * T + e);
      G[e] = gs;
    }
  };

  float yv[E];
#pragma unroll
  for (int e = 0; e < E; ++e) yv[e] = 0.f;
  // prologue: lists of blocks 0 and 1, rows and Gram of block 0
  if (warp == 0) fetch_lists(0);
  if (warp == 1) fetch_lists(1);
  __syncthreads();
  fetch_rows(0, 0, kWideScanThreads);
  fetch_gram(0, 0, kWideScanThreads);
  cp_async_wait_all();
  cluster_sync_all();                                      // (also a CTA barrier) every CTA of the cluster runs before anybody writes into it

  for (int kb = 0; kb < nblk; ++kb) {
    const int l0 = kb * T, s = kb % 3, buf = kb & 1;
    const int nvalid = min(T, M - l0);
    const float* U = Ubuf + (size_t)buf * T * kper;
    const float* X = Xbuf + (size_t)buf * T * kper;
    // ---- B: this CTA's part of z_i = <u_i, y_prev> for the block's T rows; 31-shuffle transpose-reduce per warp ----
    {
      float acc[T];
#pragma unroll
      for (int i = 0; i < T; ++i) {
        const float inv = invs[s][i];
        float a = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int c = tid + e * kWideScanThreads;
          if (c < klen) a = fmaf(__fmul_rn(U[(size_t)i * kper + c], inv), yv[e], a);     // u = little * (1/(norm+1e-8))  :109
        }
        acc[i] = a;
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int j = 0; j < o; ++j) {
          const float send = up ? acc[j] : acc[j + o];
          const float keep = up ? acc[j + o] : acc[j];
          acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      red[warp][lane] = acc[0];                            // lane L holds the warp's sum of row L
    }
    __syncthreads();
    if (tid < T) {
      float t = 0.f;
      for (int w = 0; w < NW; ++w) t += red[w][tid];        // fixed order
      const uint32_t mine = smem_u32(&zpart[buf][crank][tid]);
#pragma unroll
      for (uint32_t dst = 0; dst < (uint32_t)kWideCluster; ++dst) wide_st_cluster(wide_map_to_cta(mine, dst), t);
    }
    cluster_sync_all();                                    // all partials of this block have landed everywhere
    if (warp == 0) {
      // ---- C: T scalar steps (the same arithmetic in every CTA of the cluster) ----
      float z = 0.f;
#pragma unroll
      for (int c = 0; c < kWideCluster; ++c) z += zpart[buf][c][lane];         // rank order
      const float* G = Gbuf + (size_t)buf * T * T;
      const float v = vs[s][lane];
      float my_wn = 0.f, my_wo = 1.f;
      float g[T], vj[T];
#pragma unroll
      for (int j = 0; j < T; ++j) {
        g[j] = G[j * T + lane];
        vj[j] = __shfl_sync(0xffffffffu, v, j);
      }
#pragma unroll
      for (int j = 0; j < T; ++j) {
        const float zj = __shfl_sync(0xffffffffu, z, j);
        const float r = wide_rcp_approx(__fadd_rn(zj, vj[j]));   // no clamp: inf / nan propagate      :120
        float wn = __fmul_rn(zj, r);
        float wo = __fmul_rn(vj[j], r);                          //                                     :121
        if (kb == 0 && j == 0) {                                 // first masked patch: plain copy      :98-101
          wn = 0.f;
          wo = 1.f;
        }
        z = fmaf(wn, z, __fmul_rn(wo, g[j]));
        if (lane == j) {
          my_wn = wn;
          my_wo = wo;
        }
      }
      wn_s[lane] = my_wn;
      wo_s[lane] = my_wo;
      if (crank == 0 && l0 + lane < M) {
        wn_out[(size_t)b * M + l0 + lane] = my_wn;
        wo_out[(size_t)b * M + l0 + lane] = my_wo;
      }
    } else {
      // ---- meanwhile: everything block kb+1 needs (its lists arrived one block ago), and the lists of block kb+2 ----
      if (warp == 1) fetch_lists(kb + 2);
      fetch_rows(kb + 1, 32, kWideScanThreads - 32);
      fetch_gram(kb + 1, 32, kWideScanThreads - 32);
      cp_async_wait_all();
    }
    __syncthreads();
    // ---- D: this CTA's columns of the block's y rows ----
#pragma unroll 4
    for (int j = 0; j < nvalid; ++j) {
      const float wn = wn_s[j], wo = wo_s[j];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int c = tid + e * kWideScanThreads;
        if (c < klen) {
          yv[e] = __fadd_rn(__fmul_rn(wn, yv[e]), __fmul_rn(wo, X[(size_t)j * kper + c]));          // :122
          yb[(size_t)(l0 + j) * K + kbeg + c] = yv[e];
        }
      }
    }
    __syncthreads();                                       // this block's buffers may be refilled (two blocks from now)
  }
  cluster_sync_all();                                      // nobody leaves while a peer may still write into it
}

// out[b][c][Y][X] = sum over the patches q covering (Y, X) of src(q)[(c, dy, dx)], src(q) = y[rank[q]] for masked
// patch positions, rows[ind[q]] otherwise (IPSRFunction.py:129-131)
__global__ void __launch_bounds__(256) fold_rows_kernel(const float* __restrict__ rows, const float* __restrict__ y,
                                                        const int32_t* __restrict__ ind, const int32_t* __restrict__ rank,
                                                        PatchGeom g, int M, long long total, float* __restrict__ out) {
  const int P = g.nH * g.nW, kk2 = g.k * g.k;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(t % g.W);
    long long r = t / g.W;
    const int Y = (int)(r % g.H);
    r /= g.H;
    const int c = (int)(r % g.C);
    const int b = (int)(r / g.C);
    float acc = 0.f;
    for (int dy = 0; dy < g.k; ++dy) {
      const int yy = Y - dy;
      if (yy < 0 || yy % g.s != 0 || yy / g.s >= g.nH) continue;
      for (int dx = 0; dx < g.k; ++dx) {
        const int xx = X - dx;
        if (xx < 0 || xx % g.s != 0 || xx / g.s >= g.nW) continue;
        const int q = (yy / g.s) * g.nW + xx / g.s;
        const int rk = rank[q];
        const float* src = rk >= 0 ? y + ((size_t)b * M + rk) * g.K : rows + ((size_t)b * P + __ldg(ind + (size_t)b * P + q)) * g.K;
        acc += __ldg(src + c * kk2 + dy * g.k + dx);
      }
    }
    out[t] = acc;
  }
}

static int make_geom(const char* who, int B, int C, int H, int W, int k, int s, int Kpad, PatchGeom* g) {
  IPSR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && k > 0 && s > 0 && k <= H && k <= W, IPSR_ERR_INVALID_ARG,
               "%s: bad dims B=%d C=%d H=%d W=%d patch=%d stride=%d", who, B, C, H, W, k, s);
  g->C = C; g->H = H; g->W = W; g->k = k; g->s = s;
  g->nH = (H - k) / s + 1;
  g->nW = (W - k) / s + 1;
  g->K = C * k * k;
  g->Kpad = Kpad > 0 ? Kpad : g->K;
  IPSR_REQUIRE(g->Kpad >= g->K, IPSR_ERR_INVALID_ARG, "%s: Kpad=%d < C*k*k=%d", who, Kpad, g->K);
  return IPSR_OK;
}

static int grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = 148ll * 32;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace ipsr

extern "C" int ipsr_patch_row_len(int C, int patch) { return C * patch * patch; }

extern "C" int ipsr_unfold_patches(const float* x, int B, int C, int H, int W, int patch, int stride, int Kpad,
                                   float* cols, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(x && cols, IPSR_ERR_INVALID_ARG, "ipsr_unfold_patches: null pointer");
  PatchGeom g;
  IPSR_FORWARD(make_geom("ipsr_unfold_patches", B, C, H, W, patch, stride, Kpad, &g));
  const long long total = (long long)B * g.Kpad * g.nH * g.nW;
  unfold_cols_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, g, total, cols);
  return check_launch("ipsr_unfold_patches");
}

extern "C" int ipsr_fold_patches(const float* cols, int B, int C, int H, int W, int patch, int stride, int Kpad,
                                 float* out, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(cols && out, IPSR_ERR_INVALID_ARG, "ipsr_fold_patches: null pointer");
  PatchGeom g;
  IPSR_FORWARD(make_geom("ipsr_fold_patches", B, C, H, W, patch, stride, Kpad, &g));
  IPSR_REQUIRE((g.nH - 1) * stride + patch == H && (g.nW - 1) * stride + patch == W, IPSR_ERR_UNSUPPORTED,
               "ipsr_fold_patches: patches of size %d / stride %d do not tile a %d x %d map", patch, stride, H, W);
  const long long total = (long long)B * C * H * W;
  fold_cols_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(cols, g, total, out);
  return check_launch("ipsr_fold_patches");
}

extern "C" int ipsr_patch_rows(const float* x, int B, int C, int H, int W, int patch, int stride,
                               float* rows, float* inv_norm, void* stream) {
  return ipsr_patch_rows_stats(x, B, C, H, W, patch, stride, rows, inv_norm, nullptr, nullptr, stream);
}

extern "C" int ipsr_patch_rows_stats(const float* x, int B, int C, int H, int W, int patch, int stride,
                                     float* rows, float* inv_norm, float* norm, float* maxabs, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(x && rows, IPSR_ERR_INVALID_ARG, "ipsr_patch_rows: null pointer");
  PatchGeom g;
  IPSR_FORWARD(make_geom("ipsr_patch_rows", B, C, H, W, patch, stride, 0, &g));
  IPSR_REQUIRE(B <= 65535, IPSR_ERR_UNSUPPORTED, "ipsr_patch_rows: B=%d > 65535", B);
  patch_rows_kernel<<<dim3(g.nH * g.nW, B), 256, 0, as_stream(stream)>>>(x, g, rows, inv_norm, norm, maxabs);
  return check_launch("ipsr_patch_rows");
}

extern "C" int ipsr_blend_wide(const float* rows, const float* inv_norm, const float* vmax, const int32_t* ind,
                               const int32_t* mask_idx, int B, int K, int P, int M,
                               float* y, float* wn, float* wo, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(rows && inv_norm && vmax && ind && mask_idx && y && wn && wo, IPSR_ERR_INVALID_ARG, "ipsr_blend_wide: null pointer");
  IPSR_REQUIRE(B > 0 && K > 0 && P > 0 && M > 0 && M <= P, IPSR_ERR_INVALID_ARG, "ipsr_blend_wide: bad dims");
  IPSR_REQUIRE(K <= 8 * kWideThreads, IPSR_ERR_UNSUPPORTED, "ipsr_blend_wide: K=%d > %d", K, 8 * kWideThreads);
  cudaStream_t st = as_stream(stream);
  const int E = (K + kWideThreads - 1) / kWideThreads;
  if (E <= 1) blend_wide_kernel<1><<<B, kWideThreads, 0, st>>>(rows, inv_norm, vmax, ind, mask_idx, K, P, M, y, wn, wo);
  else if (E <= 2) blend_wide_kernel<2><<<B, kWideThreads, 0, st>>>(rows, inv_norm, vmax, ind, mask_idx, K, P, M, y, wn, wo);
  else if (E <= 4) blend_wide_kernel<4><<<B, kWideThreads, 0, st>>>(rows, inv_norm, vmax, ind, mask_idx, K, P, M, y, wn, wo);
  else blend_wide_kernel<8><<<B, kWideThreads, 0, st>>>(rows, inv_norm, vmax, ind, mask_idx, K, P, M, y, wn, wo);
  return check_launch("ipsr_blend_wide");
}

extern "C" int ipsr_blend_wide_gram_floats(int B, int M) {
  if (B <= 0 || M <= 0) return 0;
  return B * ((M + ipsr::kWideT - 1) / ipsr::kWideT) * ipsr::kWideKS * ipsr::kWideT * ipsr::kWideT;
}

namespace ipsr {
template <int E>
static int launch_wide_scan(const float* rows, const float* inv_norm, const float* vmax, const int32_t* ind, const int32_t* mask_idx,
                            const float* gram, int B, int K, int P, int M, int kper, size_t smem, float* y, float* wn, float* wo,
                            cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)B * kWideCluster);
  cfg.blockDim = dim3(kWideScanThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kWideCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
  if (smem > 0) {
    e = cudaFuncSetAttribute(wide_scan_smem_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "ipsr_blend_wide_blocked smem attribute: %s", cudaGetErrorString(e));
    e = cudaLaunchKernelEx(&cfg, wide_scan_smem_kernel<E>, rows, inv_norm, vmax, ind, mask_idx, gram, K, P, M, kper, y, wn, wo);
  } else {
    e = cudaLaunchKernelEx(&cfg, wide_scan_kernel<E>, rows, inv_norm, vmax, ind, mask_idx, gram, K, P, M, kper, y, wn, wo);
  }
  IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "ipsr_blend_wide_blocked: launch failed: %s", cudaGetErrorString(e));
  return check_launch("ipsr_blend_wide_blocked");
}
}  // namespace ipsr

extern "C" int ipsr_blend_wide_blocked(const float* rows, const float* inv_norm, const float* vmax, const int32_t* ind,
                                       const int32_t* mask_idx, int B, int K, int P, int M, float* gram,
                                       float* y, float* wn, float* wo, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(rows && inv_norm && vmax && ind && mask_idx && gram && y && wn && wo, IPSR_ERR_INVALID_ARG,
               "ipsr_blend_wide_blocked: null pointer");
  IPSR_REQUIRE(B > 0 && B <= 8000 && K > 0 && P > 0 && M > 0 && M <= P, IPSR_ERR_INVALID_ARG, "ipsr_blend_wide_blocked: bad dims");
  // columns per CTA of the cluster (a multiple of 4: 16-byte cp.async pieces), E values per thread
  const int kper = ((K + kWideCluster - 1) / kWideCluster + 3) & ~3;
  const int E = (kper + kWideScanThreads - 1) / kWideScanThreads;
  // the prefetching kernel when its buffers (two blocks of 2 T row slices + two Gram matrices) fit
  size_t smem = (4 * (size_t)kWideT * kper + 2 * (size_t)kWideT * kWideT) * sizeof(float);
  static const bool no_prefetch = [] {
    const char* e = getenv("IPSR_WIDE_NO_PREFETCH");
    return e && atoi(e) != 0;
  }();
  if (smem > 200 * 1024 || no_prefetch) smem = 0;
  IPSR_REQUIRE(E <= 8, IPSR_ERR_UNSUPPORTED, "ipsr_blend_wide_blocked: K=%d > %d", K, 8 * kWideScanThreads * kWideCluster);
  cudaStream_t st = as_stream(stream);
  const int nblk = (M + kWideT - 1) / kWideT;
  wide_gram_kernel<<<dim3(nblk, kWideKS, B), 256, 0, st>>>(rows, inv_norm, ind, mask_idx, K, P, M, gram);
  IPSR_FORWARD(check_launch("ipsr_blend_wide_blocked (gram)"));
  if (E <= 1) return launch_wide_scan<1>(rows, inv_norm, vmax, ind, mask_idx, gram, B, K, P, M, kper, smem, y, wn, wo, st);
  if (E <= 2) return launch_wide_scan<2>(rows, inv_norm, vmax, ind, mask_idx, gram, B, K, P, M, kper, smem, y, wn, wo, st);
  if (E <= 4) return launch_wide_scan<4>(rows, inv_norm, vmax, ind, mask_idx, gram, B, K, P, M, kper, smem, y, wn, wo, st);
  return launch_wide_scan<8>(rows, inv_norm, vmax, ind, mask_idx, gram, B, K, P, M, kper, smem, y, wn, wo, st);
}

extern "C" int ipsr_fold_patch_rows(const float* rows, const float* y, const int32_t* ind, const int32_t* rank,
                                    int B, int C, int H, int W, int patch, int stride, int M, float* out, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(rows && ind && rank && out && (M == 0 || y), IPSR_ERR_INVALID_ARG, "ipsr_fold_patch_rows: null pointer");
  PatchGeom g;
  IPSR_FORWARD(make_geom("ipsr_fold_patch_rows", B, C, H, W, patch, stride, 0, &g));
  IPSR_REQUIRE((g.nH - 1) * stride + patch == H && (g.nW - 1) * stride + patch == W, IPSR_ERR_UNSUPPORTED,
               "ipsr_fold_patch_rows: patches of size %d / stride %d do not tile a %d x %d map", patch, stride, H, W);
  const long long total = (long long)B * C * H * W;
  fold_rows_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(rows, y, ind, rank, g, M, total, out);
  return check_launch("ipsr_fold_patch_rows");
}
