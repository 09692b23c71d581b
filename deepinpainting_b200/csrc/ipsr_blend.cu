// (d) the "coherent" blend over masked positions and the gather-paste.
//
// Replaces models/IPSRFunction.py:82-129 (an N-iteration python loop with ~15 launches, an
// nn.Conv2d construction and two .item() syncs per masked position) and :131 (a dense N x N x C
// conv_transpose against an almost one-hot attention tensor).
//
//   ipsr_blend_stage  parallel: one warp per masked position gathers the two C-vectors the
//                     recurrence needs and computes v_l = vmax[q_l] in exact fp32;
//   ipsr_blend_scan   sequential in l (a_l depends non-linearly on y_{l-1}): one warp per image,
//                     operands streamed into a shared-memory ring by bulk async copies, the
//                     C-vector state in registers, one shuffle reduction per step;
//   ipsr_paste        out[b,c,q] = x[b,c,ind[q]] (row staged in shared memory, gather by index)
//                     or y[b,rank[q],c] at masked positions; HBM traffic = read x + write out.
#include <stdlib.h>

#include "ipsr_bookkeeping.cuh"

namespace ipsr {

// Steps are processed in blocks of T consecutive masked positions (T = scan_block_steps(C)).  One staged
// block = everything the scan needs for T steps, contiguous, so that ONE bulk copy fetches it:
//   U  [T][C]  u_l = X[q_l] * inv_norm[q_l]                       (IPSRFunction.py:109)
//   K  [T][C]  X[p_l], the matched bank patch                     (:95)
//   Gt [T][T]  Gt[j][i] = <u_{l0+i}, X[p_{l0+j}>                  (in-block Gram matrix, transposed)
//   v  [T]     v_l = <R[q_l], Xn[p_l]>  exact fp32                (vmax at masked positions, :70)
__host__ __device__ inline int scan_block_steps(int C) { return C <= 256 ? 32 : (C <= 512 ? 16 : 8); }
__host__ __device__ inline int staged_block_floats(int C) {
  const int T = scan_block_steps(C);
  return 2 * T * C + T * T + T;
}
// rows of y per image: M rounded up to a multiple of 8
__host__ __device__ inline int padded_steps(int M) { return (M + 7) & ~7; }

// ---------------------------------------------------------------------------------------------
// stage: grid = (ceil(M / T), B), 256 threads.
//   phase 1  one warp per step: gather u_l, X[p_l] (to shared memory and to the staged block) and v_l;
//   phase 2  the T x T Gram matrix of the block from shared memory, 2 x 2 register tiles, the channel
//            range split over 256 / (T/2)^2 thread groups whose partial sums are added in fixed order.
// ---------------------------------------------------------------------------------------------
struct StageArgs {
  const float* xt; const float* r_masked; const float* inv_norm; const int* ind; const int* mask_idx;
  int B, C, N, M, nblocks;
  float* staged; float* vmask;
  // optional routes builders riding in the same launch (they depend on ind only): CTAs [0, n_routes)
  int n_routes; const int* flag; int* route_ptr; int* route_q;
  int ms; const int* mcount;        // per-image masks: flag / mask_idx rows ms = N apart, mcount[b] steps (ms = 0: shared)
};

template <int T>
__device__ __forceinline__ void
blend_stage_cta(int kblk, int b, float* stage_smem, const float* __restrict__ xt, const float* __restrict__ r_masked,
                const float* __restrict__ inv_norm, const int* __restrict__ ind, const int* __restrict__ mask_idx,
                int C, int N, int M, int nblocks, float* __restrict__ staged, float* __restrict__ vmask,
                int ms, const int* __restrict__ mcount) {
  mask_idx += (size_t)b * ms;
  const int Mc = mcount ? mcount[b] : M;                 // steps of THIS image; M stays the row stride of the batch
  constexpr int kTiles = (T / 2) * (T / 2);              // 2 x 2 output tiles
  constexpr int kSplit = 256 / kTiles;                   // channel groups (1, 4 or 16)
  const int ld = C + 4;                                  // padded row: conflict-free float4 reads across rows
  float* Us = stage_smem;                                // [T][ld]
  float* Ks = Us + (size_t)T * ld;                       // [T][ld]
  float* red = Ks + (size_t)T * ld;                      // [kSplit][T*T] (kSplit > 1 only)
  const int l0 = kblk * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* blk = staged + ((size_t)b * nblocks + kblk) * staged_block_floats(C);
  float* gU = blk;
  float* gK = blk + (size_t)T * C;
  float* gG = gK + (size_t)T * C;
  float* gV = gG + T * T;

#pragma unroll
  for (int rr_ = 0; rr_ < (T + 7) / 8; ++rr_) {           // unrolled: the index chains of the warp's rows overlap
    const int r = warp + 8 * rr_;
    if (r >= T) break;
    const int l = l0 + r;
    float* su = Us + (size_t)r * ld;
    float* sk = Ks + (size_t)r * ld;
    if (l >= Mc) {                                       // tail of the last block: inert steps
      for (int c = lane * 4; c < C; c += 128) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(su + c) = z4;
        *reinterpret_cast<float4*>(sk + c) = z4;
        *reinterpret_cast<float4*>(gU + (size_t)r * C + c) = z4;
        *reinterpret_cast<float4*>(gK + (size_t)r * C + c) = z4;
      }
      if (lane == 0) gV[r] = 1.f;
      continue;
    }
    const int q = mask_idx[l];
    const int p = ind[(size_t)b * N + q];
    const float inv_q = inv_norm[(size_t)b * N + q];
    const float inv_p = inv_norm[(size_t)b * N + p];
    const float* xq = xt + ((size_t)b * N + q) * C;
    const float* xp = xt + ((size_t)b * N + p) * C;
    const float* rq = r_masked + ((size_t)b * M + l) * C;
    float acc = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(xq + c));
      const float4 k = __ldg(reinterpret_cast<const float4*>(xp + c));
      const float4 rr = __ldg(reinterpret_cast<const float4*>(rq + c));
      // u = little_value * (1/(norm+1e-8))                          IPSRFunction.py:109
      const float4 u = make_float4(__fmul_rn(a.x, inv_q), __fmul_rn(a.y, inv_q), __fmul_rn(a.z, inv_q), __fmul_rn(a.w, inv_q));
      *reinterpret_cast<float4*>(su + c) = u;
      *reinterpret_cast<float4*>(sk + c) = k;
      *reinterpret_cast<float4*>(gU + (size_t)r * C + c) = u;
      *reinterpret_cast<float4*>(gK + (size_t)r * C + c) = k;
      // v = <R[q], Xn[p]> with Xn = fl(X * inv) as the encoder weights hold it (NPS:40)
      acc = fmaf(rr.x, __fmul_rn(k.x, inv_p), acc);
      acc = fmaf(rr.y, __fmul_rn(k.y, inv_p), acc);
      acc = fmaf(rr.z, __fmul_rn(k.z, inv_p), acc);
      acc = fmaf(rr.w, __fmul_rn(k.w, inv_p), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      gV[r] = acc;
      if (vmask) vmask[(size_t)b * M + l] = acc;
    }
  }
  __syncthreads();

  // Gram: tile (ti, tj) -> rows i = ti, ti + T/2 of U against rows j = tj, tj + T/2 of K
  const int tile = threadIdx.x % kTiles, ks = threadIdx.x / kTiles;
  const int ti = tile / (T / 2), tj = tile % (T / 2);
  const int cper = ((C / 4 + kSplit - 1) / kSplit) * 4;
  const int cbeg = ks * cper, cend = min(C, cbeg + cper);
  // rows (ti, ti + T/2) x (tj, tj + T/2): the 8 lanes of a quarter warp read 8 CONSECUTIVE rows of K, whose padded
  // stride of C + 4 floats puts their 16-byte vectors on 8 different bank groups (conflict-free)
  const float* k0 = Ks + (size_t)tj * ld;
  const float* k1 = k0 + (size_t)(T / 2) * ld;
  const int i0 = ti, i1 = ti + T / 2, j0 = tj, j1 = tj + T / 2;
  const float* u0 = Us + (size_t)ti * ld;
  const float* u1 = u0 + (size_t)(T / 2) * ld;
  float g00 = 0.f, g01 = 0.f, g10 = 0.f, g11 = 0.f;
#pragma unroll 4
  for (int c = cbeg; c < cend; c += 4) {
    const float4 a0 = *reinterpret_cast<const float4*>(u0 + c);
    const float4 a1 = *reinterpret_cast<const float4*>(u1 + c);
    const float4 b0 = *reinterpret_cast<const float4*>(k0 + c);
    const float4 b1 = *reinterpret_cast<const float4*>(k1 + c);
    g00 = fmaf(a0.x, b0.x, g00); g00 = fmaf(a0.y, b0.y, g00); g00 = fmaf(a0.z, b0.z, g00); g00 = fmaf(a0.w, b0.w, g00);
    g01 = fmaf(a0.x, b1.x, g01); g01 = fmaf(a0.y, b1.y, g01); g01 = fmaf(a0.z, b1.z, g01); g01 = fmaf(a0.w, b1.w, g01);
    g10 = fmaf(a1.x, b0.x, g10); g10 = fmaf(a1.y, b0.y, g10); g10 = fmaf(a1.z, b0.z, g10); g10 = fmaf(a1.w, b0.w, g10);
    g11 = fmaf(a1.x, b1.x, g11); g11 = fmaf(a1.y, b1.y, g11); g11 = fmaf(a1.z, b1.z, g11); g11 = fmaf(a1.w, b1.w, g11);
  }
  if (kSplit == 1) {
    gG[j0 * T + i0] = g00;
    gG[j1 * T + i0] = g01;
    gG[j0 * T + i1] = g10;
    gG[j1 * T + i1] = g11;
  } else {
    float* mine = red + (size_t)ks * T * T;
    mine[j0 * T + i0] = g00;
    mine[j1 * T + i0] = g01;
    mine[j0 * T + i1] = g10;
    mine[j1 * T + i1] = g11;
    __syncthreads();
    for (int e = threadIdx.x; e < T * T; e += 256) {
      float s = red[e];
      for (int k2 = 1; k2 < kSplit; ++k2) s += red[(size_t)k2 * T * T + e];
      gG[e] = s;
    }
  }
}

// grid = n_routes + ceil(M / T) * B CTAs; the latency-bound route builders are scheduled first
template <int T>
__global__ void __launch_bounds__(256) blend_stage_kernel(const StageArgs a) {
  extern __shared__ __align__(16) float stage_smem[];
  pdl_trigger();
  pdl_wait();
  int blk = blockIdx.x;
  if (blk < a.n_routes) {
    build_routes_cta(blk, reinterpret_cast<int*>(stage_smem), a.ind, a.flag, a.mask_idx, a.N, a.M, a.route_ptr, a.route_q, a.ms,
                     a.mcount);
    return;
  }
  blk -= a.n_routes;
  blend_stage_cta<T>(blk % a.nblocks, blk / a.nblocks, stage_smem, a.xt, a.r_masked, a.inv_norm, a.ind, a.mask_idx, a.C, a.N,
                     a.M, a.nblocks, a.staged, a.vmask, a.ms, a.mcount);
}

// ---------------------------------------------------------------------------------------------
// scan
//
// The recurrence of IPSRFunction.py:104-122,
//     a_l = <u_l, y_{l-1}>,  wn_l = a_l/(a_l+v_l),  wo_l = v_l/(a_l+v_l),  y_l = wn_l y_{l-1} + wo_l X[p_l],
// is sequential in l, but only through ONE scalar per step.  By linearity the scalars
//     z_i = <u_i, y_cur>   (i = the steps still ahead inside the block of T steps)
// follow y:  z_i <- wn_l z_i + wo_l <u_i, X[p_l]> = wn_l z_i + wo_l Gt[l][i],  so that a_l is simply z_l by the time step
// l is reached.  One CTA per image, per block of T steps:
//   B  warps 1..7: z_i = <u_i, y_prev> for the T rows of the block (re-anchors z on the real y every T steps, so
//                  rounding differences to the reference's dot product cannot accumulate); warp 0 meanwhile loads the
//                  chain's operands
//   C  warp 0    : the T dependent steps on scalars only; warps 1..7 meanwhile store the previous block's y rows
//   D  warps 0..7: y_l = wn_l y_{l-1} + wo_l X[p_l] channel-parallel, exactly the reference's two rounded products and
//                  one sum (:122), left as 16-byte vectors in a [C][T+4] tile that the store reads row-wise
//   warp 8       : the bulk async copies of the two-stage ring (issuing one costs its thread ~650 cycles:
//                  scripts/micro/bulk_bw.cu)
// The chain step: every lane carries the DIAGONAL element zd = z_j (the one step j divides by) itself -- lane j+1's state
// before step j is broadcast by a shuffle issued at the START of step j, and every lane applies step j's update to it with
// the formula lane j+1 uses for its own state (same bits) -- so the shuffle runs next to the reciprocal instead of in
// front of it: one add, one reciprocal, one product and one fma on the dependent path (~47 cycles; with the shuffle in
// the path it was ~165).  An alternative that takes B and D off the chain's path as well (z across a block boundary from a
// second Gram matrix <u_{next block}, X[p_l]>) was built and measured: 20.6 us against 21 us here at 32x32 (the workers
// become the limit), and the second Gram costs the stage kernel 60 us at 64x64x256 -- not kept.
// wn, wo use a * rcp(a+v): <= 2 ulp from the reference's IEEE divisions.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

constexpr int kScanMaxCpt = 4;                             // channels per thread: C <= 1024
constexpr int kScanCompute = 256;                          // warps 0..7
constexpr int kScanThreads = kScanCompute + 32;            // + the copy warp

__device__ __forceinline__ void scan_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kScanCompute) : "memory"); }

template <int T>
__global__ void __launch_bounds__(kScanThreads)
blend_scan_kernel(const float* __restrict__ staged, int C, int Mmax,
                  float* __restrict__ y, float* __restrict__ wn_out, float* __restrict__ wo_out,
                  const int* __restrict__ mcount) {
  extern __shared__ __align__(128) uint8_t scan_smem[];
  __shared__ __align__(8) unsigned long long bars[4];      // full[2], free[2]
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float zs[T], wn_s[T], wo_s[T];
  const int blk_floats = 2 * T * C + T * T + T;
  const uint32_t blk_bytes = (uint32_t)blk_floats * sizeof(float);
  const uint32_t stage_bytes = (blk_bytes + 127u) & ~127u;
  float* ysm = reinterpret_cast<float*>(scan_smem + 2 * (size_t)stage_bytes);     // [C] y at the end of the previous block
  float* ytile = ysm + C;                                                        // [C][T+4] the block's y rows
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // per-image masks: strides come from the batch maximum, the steps walked from this image's own count
  const int M = mcount ? mcount[b] : Mmax;
  const int nblocks = (M + T - 1) / T;
  const float* src = staged + (size_t)b * ((Mmax + T - 1) / T) * blk_floats;
  const int Mp = padded_steps(Mmax);
  float* yb = y + (size_t)b * C * Mp;                    // [C][Mp]: channel-major, so that the paste reads rows
  float* wnb = wn_out + (size_t)b * Mmax;
  float* wob = wo_out + (size_t)b * Mmax;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto free_bar = [&](int s) { return bar0 + 8u * (2 + s); };

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(bar0 + 8u * s, 1);
    mbar_fence_init();
  }
  for (int c = tid; c < C; c += kScanThreads) ysm[c] = 0.f;
  __syncthreads();
  if (nblocks == 0) return;

  if (warp == kScanCompute / 32) {
    // ------------------------------------------------------------------ the copy warp
    if (lane != 0) return;
    auto issue = [&](int k) {
      const int s = k & 1;
      mbar_expect_tx(full_bar(s), blk_bytes);
      bulk_g2s(smem_u32(scan_smem) + (uint32_t)s * stage_bytes, src + (size_t)k * blk_floats, blk_bytes, full_bar(s));
    };
    issue(0);
    if (nblocks > 1) issue(1);
    for (int k = 0; k + 2 < nblocks; ++k) {
      mbar_wait(free_bar(k & 1), (uint32_t)(k >> 1) & 1u);  // the compute warps are done with block k's stage
      issue(k + 2);
    }
    return;
  }

  float yreg[kScanMaxCpt];
#pragma unroll
  for (int m = 0; m < kScanMaxCpt; ++m) yreg[m] = 0.f;

  // y[c][l0 .. ) of block kb from ytile: T/4 lanes per channel row, 16 bytes each (rows of y are padded to a multiple of
  // 8 steps, so the last vector of the image's last block stays inside its row).  Runs OFF the critical path: warps 1..7
  // store block k-1 while warp 0 walks the scalar steps of block k.
  auto store_ytile = [&](int kb, int w0, int nw) {
    constexpr int LPC = T / 4, CPI = 32 / LPC;             // lanes per channel, channels per warp instruction
    const int lb = kb * T;
    const int nv = min(T, M - lb);
    const int lc = lane / LPC, l4 = lane % LPC;
    if (4 * l4 < nv) {
#pragma unroll 4
      for (int c = (warp - w0) * CPI + lc; c < C; c += nw * CPI)
        *reinterpret_cast<float4*>(yb + (size_t)c * Mp + lb + 4 * l4) =
            *reinterpret_cast<const float4*>(ytile + (size_t)c * (T + 4) + 4 * l4);
    }
  };

  for (int k = 0; k < nblocks; ++k) {
    const int s = k & 1;
    mbar_wait(full_bar(s), (uint32_t)(k >> 1) & 1u);
    const float* U = reinterpret_cast<const float*>(scan_smem + (size_t)s * stage_bytes);
    const float* K = U + (size_t)T * C;
    const float* Gt = K + (size_t)T * C;
    const float* V = Gt + T * T;
    const int l0 = k * T;

    // ---- B: z_i = <u_i, y_prev> by warps 1..7 (ceil(T/7) rows each); warp 0 meanwhile fetches what its chain needs and
    //         what does not depend on z: v of every step, v_j times the Gram column of its lane and times the element
    //         next to the diagonal ----
    const int li = lane < T ? lane : T - 1;
    float vg[T], vgd[T], vj[T];
    if (warp == 0) {
#pragma unroll
      for (int j = 0; j < T; ++j) {
        vj[j] = V[j];                                           // (uniform address: broadcast)
        vg[j] = __fmul_rn(vj[j], Gt[j * T + li]);
        vgd[j] = (j + 1 < T) ? __fmul_rn(vj[j], Gt[j * T + j + 1]) : 0.f;
      }
    } else {
      constexpr int NW = kScanCompute / 32 - 1;              // 7 warps
      constexpr int R = (T + NW - 1) / NW;                   // rows per warp
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll 2
      for (int c = lane * 4; c < C; c += 128) {
        const float4 yv = *reinterpret_cast<const float4*>(ysm + c);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = (warp - 1) + r * NW;
          if (i < T) {
            const float4 u = *reinterpret_cast<const float4*>(U + (size_t)i * C + c);
            acc[r] = fmaf(u.x, yv.x, acc[r]);
            acc[r] = fmaf(u.y, yv.y, acc[r]);
            acc[r] = fmaf(u.z, yv.z, acc[r]);
            acc[r] = fmaf(u.w, yv.w, acc[r]);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = (warp - 1) + r * NW;
          if (i < T) zs[i] = acc[r];
        }
      }
    }
    scan_barrier();

    // ---- C: T scalar steps ----
    if (warp == 0) {
      float z = zs[li];
      float zd = zs[0];
      float my_wn = 0.f, my_wo = 1.f;
#pragma unroll
      for (int j = 0; j < T; ++j) {
        const float p = (j + 1 < T) ? __shfl_sync(0xffffffffu, z, j + 1) : 0.f;
        const float r = rcp_approx(__fadd_rn(zd, vj[j]));       // no clamp: inf / nan propagate      :120
        float wn = __fmul_rn(zd, r);
        float wo = __fmul_rn(vj[j], r);                         //                                     :121
        float tz = __fmul_rn(r, vg[j]), td = __fmul_rn(r, vgd[j]);   // wo * <u_i, X[p_j]>
        if (k == 0 && j == 0) {                                 // first masked patch: plain copy      :98-101
          wn = 0.f;
          wo = 1.f;
          tz = Gt[li];
          td = Gt[1 < T ? 1 : 0];
        }
        z = fmaf(wn, z, tz);
        zd = fmaf(wn, p, td);
        if (lane == j) {
          my_wn = wn;
          my_wo = wo;
        }
      }
      if (lane < T) {
        wn_s[lane] = my_wn;
        wo_s[lane] = my_wo;
        if (l0 + lane < M) {
          wnb[l0 + lane] = my_wn;
          wob[l0 + lane] = my_wo;
        }
      }
    } else if (k > 0) {
      store_ytile(k - 1, 1, kScanCompute / 32 - 1);
    }
    scan_barrier();

    // ---- D: y rows of the block; a thread owns a channel (lane stride T+4 floats: conflict-free vector stores) ----
    const int nvalid = min(T, M - l0);
#pragma unroll
    for (int m = 0; m < kScanMaxCpt; ++m) {
      const int c = tid + m * kScanCompute;
      if (c < C) {
        float yy = yreg[m];
        float* yrow = ytile + (size_t)c * (T + 4);
        if (nvalid == T) {                                   // full block: operands first, then the 2-op chain
          float kv[T];
#pragma unroll
          for (int j = 0; j < T; ++j) kv[j] = K[(size_t)j * C + c];
#pragma unroll
          for (int j4 = 0; j4 < T / 4; ++j4) {
            const float4 a4 = *reinterpret_cast<const float4*>(&wn_s[4 * j4]);
            const float4 o4 = *reinterpret_cast<const float4*>(&wo_s[4 * j4]);
            float4 o;
            o.x = yy = __fadd_rn(__fmul_rn(a4.x, yy), __fmul_rn(o4.x, kv[4 * j4 + 0]));          // :122
            o.y = yy = __fadd_rn(__fmul_rn(a4.y, yy), __fmul_rn(o4.y, kv[4 * j4 + 1]));
            o.z = yy = __fadd_rn(__fmul_rn(a4.z, yy), __fmul_rn(o4.z, kv[4 * j4 + 2]));
            o.w = yy = __fadd_rn(__fmul_rn(a4.w, yy), __fmul_rn(o4.w, kv[4 * j4 + 3]));
            *reinterpret_cast<float4*>(yrow + 4 * j4) = o;
          }
        } else {                                             // the image's last block: zeros behind its last step
          for (int j = 0; j < T; ++j) {
            if (j < nvalid) yy = __fadd_rn(__fmul_rn(wn_s[j], yy), __fmul_rn(wo_s[j], K[(size_t)j * C + c]));   // :122
            yrow[j] = j < nvalid ? yy : 0.f;
          }
        }
        yreg[m] = yy;
        ysm[c] = yy;
      }
    }
    scan_barrier();                                        // ysm / ytile complete; stage s no longer read
    if (tid == 0) mbar_arrive(free_bar(s));
  }
  store_ytile(nblocks - 1, 0, kScanCompute / 32);          // the last block: every warp helps
}

// ---------------------------------------------------------------------------------------------
// paste
// ---------------------------------------------------------------------------------------------
// A CTA owns a contiguous range of channel tiles (CT rows each) of ONE image.  It stages rank[] and ind[b][]
// in shared memory once, then streams the tiles of x[b] (CT * N contiguous floats in NCHW) through a two-stage
// ring of bulk async copies: the copy of tile t+1 overlaps the gather of tile t, every index lookup is a
// shared-memory access, reads and writes of HBM are coalesced and the gather happens inside the SM.
__device__ __forceinline__ void
paste_cta(int part, int b, int tiles_per_cta, float* psm, const float* __restrict__ x, const float* __restrict__ y,
          const int* __restrict__ ind, const int* __restrict__ rank, int C, int N, int M, int CT,
          float* __restrict__ out, int ms = 0, const PasteLoss* cos = nullptr, int cta_index = 0, int cta_count = 0,
          long long loss_count = 0) {
  __shared__ __align__(8) unsigned long long paste_bars[2];
  __shared__ float loss_wsum[32];
  __shared__ bool loss_last;
  float loss_acc = 0.f;
  const float* tgt = cos ? cos->target + (size_t)b * C * N : nullptr;
  rank += (size_t)b * ms;                                      // per-image masks: rank is [B][ms]
  const int ntiles = (C + CT - 1) / CT;
  const int t0 = part * tiles_per_cta;
  const int t1 = min(ntiles, t0 + tiles_per_cta);
  if (t0 >= t1 && !cos) return;                                // (with the fused loss every CTA takes its ticket below)
  const int nthreads = blockDim.x;
  const int Mp = padded_steps(M);
  const int tile_x = CT * N, tile_y = CT * Mp, tile_elems = tile_x + tile_y;
  float* rows0 = psm;                                          // [2][CT*N + CT*Mp]: rows of x, then rows of y
  int* ind_s = reinterpret_cast<int*>(psm + 2 * (size_t)tile_elems);   // [N]
  int* rank_s = ind_s + N;                                     // [N]
  const float* ximg = x + (size_t)b * C * N;
  const float* yimg = y + (size_t)b * C * Mp;                  // [C][Mp] (ipsr_blend_scan)
  float* oimg = out + (size_t)b * C * N;
  const bool bulk = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(ximg) & 15) == 0) &&
                    (M == 0 || (reinterpret_cast<uintptr_t>(yimg) & 15) == 0);

  auto load_tile = [&](int t, int buf) {                       // thread 0 (bulk) or everybody (fallback)
    const int ct = min(CT, C - t * CT);
    const float* src = ximg + (size_t)t * tile_x;
    const float* srcy = yimg + (size_t)t * tile_y;
    float* dst = rows0 + (size_t)buf * tile_elems;
    if (bulk) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(&paste_bars[buf]), (uint32_t)(ct * (N + (M > 0 ? Mp : 0))) * 4u);
        bulk_g2s(smem_u32(dst), src, (uint32_t)(ct * N) * 4u, smem_u32(&paste_bars[buf]));
        if (M > 0) bulk_g2s(smem_u32(dst + tile_x), srcy, (uint32_t)(ct * Mp) * 4u, smem_u32(&paste_bars[buf]));
      }
    } else {
      for (int i = threadIdx.x; i < ct * N; i += nthreads) dst[i] = __ldg(src + i);
      if (M > 0)
        for (int i = threadIdx.x; i < ct * Mp; i += nthreads) dst[tile_x + i] = __ldg(srcy + i);
    }
  };
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&paste_bars[0]), 1);
    mbar_init(smem_u32(&paste_bars[1]), 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (t0 < t1) load_tile(t0, 0);
  if (t0 + 1 < t1) load_tile(t0 + 1, 1);
  const int* indb = ind + (size_t)b * N;
  for (int q = threadIdx.x; q < N; q += nthreads) {
    ind_s[q] = __ldg(indb + q);
    rank_s[q] = __ldg(rank + q);
  }
  __syncthreads();
  for (int t = t0; t < t1; ++t) {
    const int buf = (t - t0) & 1;
    const int ct = min(CT, C - t * CT);
    const float* rows = rows0 + (size_t)buf * tile_elems;
    const float* yrows = rows + tile_x;
    float* ob = oimg + (size_t)t * tile_x;
    if (bulk) mbar_wait(smem_u32(&paste_bars[buf]), (uint32_t)((t - t0) >> 1) & 1u);
    else __syncthreads();
    if (!cos) {
      for (int q = threadIdx.x; q < N; q += nthreads) {
        const int l = rank_s[q];
        const float* srow = (l < 0) ? rows + ind_s[q] : yrows + l;
        const int stride = (l < 0) ? N : Mp;
        for (int ch = 0; ch < ct; ++ch) ob[(size_t)ch * N + q] = srow[ch * stride];
      }
    } else {                                                   // the same, plus the side loss on the values just pasted
      const float* tg = tgt + (size_t)t * tile_x;
      for (int q = threadIdx.x; q < N; q += nthreads) {
        const int l = rank_s[q];
        const float* srow = (l < 0) ? rows + ind_s[q] : yrows + l;
        const int stride = (l < 0) ? N : Mp;
        const float mq = __ldg(cos->mask + q);
        for (int ch = 0; ch < ct; ++ch) {
          const float v = srow[ch * stride];
          ob[(size_t)ch * N + q] = v;
          // (in_data * mask) * strength - target, the reference's operation order     InnerCos.py:33-36
          const float d = __fmul_rn(__fmul_rn(v, mq), cos->strength) - __ldg(tg + (size_t)ch * N + q);
          loss_acc += cos->crit == 0 ? d * d : fabsf(d);
        }
      }
    }
    __syncthreads();                                           // everybody is done reading this buffer
    if (t + 2 < t1) load_tile(t + 2, buf);
  }
  if (cos) {
    // deterministic two-level reduction: one partial per CTA, the ticket-elected last CTA adds them in index order
    loss_acc = warp_sum(loss_acc);
    if ((threadIdx.x & 31) == 0) loss_wsum[threadIdx.x >> 5] = loss_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tsum = 0.f;
      for (int w = 0; w < (nthreads >> 5); ++w) tsum += loss_wsum[w];
      cos->partials[cta_index] = tsum;
      __threadfence();
      loss_last = (atomicAdd(cos->ticket, 1u) == (unsigned)cta_count - 1u);
    }
    __syncthreads();
    if (loss_last && threadIdx.x == 0) {
      __threadfence();
      double tsum = 0.0;
      for (int i = 0; i < cta_count; ++i) tsum += (double)((volatile float*)cos->partials)[i];
      *cos->loss = (float)(tsum / (double)loss_count);
      *cos->ticket = 0u;
    }
  }
}

// Tiles per CTA for a persistent tile loop: split every image's `ntiles` tiles over as many CTAs as fill whole
// waves of `slots` resident CTAs, charging the per-CTA setup (index staging) about one tile.
static int tiles_per_cta_for(int B, int ntiles, int slots) {
  int best_tpc = ntiles;
  double best_cost = 1e300;
  for (int pcand = 1; pcand <= ntiles; ++pcand) {
    const int tpc = (ntiles + pcand - 1) / pcand;
    const int np = (ntiles + tpc - 1) / tpc;
    const long long waves = ((long long)B * np + slots - 1) / slots;
    const double cost = (double)waves * (tpc + 1.0);
    if (cost < best_cost * 0.9999) {
      best_cost = cost;
      best_tpc = tpc;
    }
  }
  return best_tpc;
}

// grid = (parts, B)
__global__ void __launch_bounds__(512)
paste_kernel(const float* __restrict__ x, const float* __restrict__ y, const int* __restrict__ ind,
             const int* __restrict__ rank, int C, int N, int M, int CT, int tiles_per_cta, float* __restrict__ out, int ms) {
  extern __shared__ __align__(128) float rows[];         // [2][CT][N] + ind[N] + rank[N]
  pdl_trigger();
  pdl_wait();
  paste_cta(blockIdx.x, blockIdx.y, tiles_per_cta, rows, x, y, ind, rank, C, N, M, CT, out, ms);
}

// the paste with the InnerCos side loss fused in (the loss arguments by value: no pointer chasing on the device)
__global__ void __launch_bounds__(512)
paste_loss_kernel(const float* __restrict__ x, const float* __restrict__ y, const int* __restrict__ ind,
                  const int* __restrict__ rank, int C, int N, int M, int CT, int tiles_per_cta, float* __restrict__ out, int ms,
                  const PasteLoss cos, long long loss_count) {
  extern __shared__ __align__(128) float rows[];
  pdl_trigger();
  pdl_wait();
  paste_cta(blockIdx.x, blockIdx.y, tiles_per_cta, rows, x, y, ind, rank, C, N, M, CT, out, ms, &cos,
            (int)(blockIdx.y * gridDim.x + blockIdx.x), (int)(gridDim.x * gridDim.y), loss_count);
}

// The paste and the two bookkeeping builders of the backward are independent once the scan is done:
// one launch, block ranges = [routes: B CTAs][exceptions: B CTAs][paste: B * C/CT CTAs]  (exc_total is the
// [2B + 2] exception state of ipsr_build_exceptions, zero on entry).
// The latency-bound bookkeeping CTAs are scheduled first and overlap the bandwidth-bound paste.
struct FusedPasteArgs {
  const float* x; const float* y; const int* ind; const int* rank; float* out;
  int B, C, N, M, CT, parts, tiles_per_cta;
  const int* flag; const int* mask_idx; int* route_ptr; int* route_q;
  const float* wn; const float* wo; int* exc_start; int* exc_cnt; int* exc_l; float* exc_w; int* exc_total; int exc_cap;
  int n_routes, n_exc;
  int ms; const int* mcount;
};

__global__ void __launch_bounds__(512) paste_fused_kernel(const FusedPasteArgs a) {
  extern __shared__ __align__(128) float fsm[];
  int blk = blockIdx.x;
  if (blk < a.n_routes) {
    build_routes_cta(blk, reinterpret_cast<int*>(fsm), a.ind, a.flag, a.mask_idx, a.N, a.M, a.route_ptr, a.route_q, a.ms, a.mcount);
    return;
  }
  blk -= a.n_routes;
  if (blk < a.n_exc) {
    build_exceptions_cta(blk, a.B, fsm, a.ind, a.mask_idx, a.wn, a.wo, a.N, a.M, a.exc_start, a.exc_cnt, a.exc_l, a.exc_w,
                         a.exc_total, a.exc_cap, a.ms, a.mcount);
    return;
  }
  blk -= a.n_exc;
  paste_cta(blk % a.parts, blk / a.parts, a.tiles_per_cta, fsm, a.x, a.y, a.ind, a.rank, a.C, a.N, a.M, a.CT, a.out, a.ms);
}

static int paste_ct(int C, int N) {
  // channel rows per tile: ~32 KiB of x rows (N > 2048: 64 KiB), at least 1 row
  int CT = (int)(((N <= 2048 ? 32 : 64) * 1024) / ((size_t)N * sizeof(float)));
  if (CT < 1) CT = 1;
  if (CT > 16) CT = 16;
  if (CT > C) CT = C;
  return CT;
}
static size_t paste_smem(int CT, int N, int M) {
  return 2 * (size_t)CT * (N + padded_steps(M)) * sizeof(float) + 2 * (size_t)N * sizeof(int);
}
static int paste_threads(int N) { return N > 2048 ? 512 : 256; }
static int paste_slots(size_t smem, int threads) {
  int resident = (int)((227 * 1024) / (smem + 2 * 1024));
  if (resident < 1) resident = 1;
  if (resident > 2048 / threads) resident = 2048 / threads;
  return 148 * resident;
}

}  // namespace ipsr

extern "C" int ipsr_scan_block_steps(int C) { return ipsr::scan_block_steps(C); }
extern "C" int ipsr_staged_block_floats(int C) { return ipsr::staged_block_floats(C); }
extern "C" int ipsr_padded_steps(int M) { return ipsr::padded_steps(M); }

namespace ipsr {
template <int T>
static int launch_stage(StageArgs a, cudaStream_t st) {
  constexpr int kSplit = 256 / ((T / 2) * (T / 2));
  size_t smem = ((size_t)2 * T * (a.C + 4) + (kSplit > 1 ? (size_t)kSplit * T * T : 0)) * sizeof(float);
  if (a.n_routes > 0) {
    const size_t smem_routes = (size_t)(2 * a.N + 1) * sizeof(int);
    if (smem_routes > smem) smem = smem_routes;
  }
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_blend_stage: C=%d / N=%d too large", a.C, a.N);
  if (smem + 2048 > 48 * 1024) {   // dynamic + static shared memory above the default limit (set per call: the attribute is per device)
    cudaError_t e = cudaFuncSetAttribute(blend_stage_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "blend_stage smem attribute: %s", cudaGetErrorString(e));
  }
  a.nblocks = (a.M + T - 1) / T;
  const long long ctas = (long long)a.n_routes + (long long)a.nblocks * a.B;
  IPSR_REQUIRE(ctas <= 0x7FFFFFFFll, IPSR_ERR_UNSUPPORTED, "ipsr_blend_stage: grid too large");
  {
    cudaError_t le__ = launch_pdl(blend_stage_kernel<T>, dim3((unsigned)ctas), dim3(256), smem, st, a);
    IPSR_REQUIRE(le__ == cudaSuccess, IPSR_ERR_CUDA, "ipsr_blend_stage: launch failed: %s", cudaGetErrorString(le__));
  }
  return check_launch("ipsr_blend_stage");
}

static int dispatch_stage(const StageArgs& a, cudaStream_t st) {
  switch (scan_block_steps(a.C)) {
    case 32: return launch_stage<32>(a, st);
    case 16: return launch_stage<16>(a, st);
    default: return launch_stage<8>(a, st);
  }
}

template <int T>
static int launch_scan(const float* staged, int B, int C, int M, float* y, float* wn, float* wo, cudaStream_t st,
                       const int32_t* mcount) {
  const size_t blk_bytes = (size_t)staged_block_floats(C) * sizeof(float);
  const size_t smem = 2 * ((blk_bytes + 127) & ~(size_t)127) + ((size_t)C + (size_t)C * (T + 4)) * sizeof(float);
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_blend_scan: C=%d too large", C);
  if (smem + 2048 > 48 * 1024) {   // dynamic + static shared memory above the default limit (set per call: the attribute is per device)
    cudaError_t e = cudaFuncSetAttribute(blend_scan_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "blend_scan smem attribute: %s", cudaGetErrorString(e));
  }
  {
    cudaError_t le__ = launch_pdl(blend_scan_kernel<T>, dim3(B), dim3(kScanThreads), smem, st, staged, C, M, y, wn, wo, mcount);
    IPSR_REQUIRE(le__ == cudaSuccess, IPSR_ERR_CUDA, "ipsr_blend_scan: launch failed: %s", cudaGetErrorString(le__));
  }
  return check_launch("ipsr_blend_scan");
}
}  // namespace ipsr

extern "C" int ipsr_blend_stage(const float* xt, const float* r_masked, const float* inv_norm,
                                const int32_t* ind, const int32_t* mask_idx, int B, int C, int N, int M,
                                float* staged, float* vmask, void* stream) {
  return ipsr_blend_stage_with_routes(xt, r_masked, inv_norm, ind, mask_idx, nullptr, B, C, N, M, staged, vmask, nullptr,
                                      nullptr, stream);
}

extern "C" int ipsr_blend_stage_with_routes(const float* xt, const float* r_masked, const float* inv_norm,
                                            const int32_t* ind, const int32_t* mask_idx, const int32_t* flag,
                                            int B, int C, int N, int M, float* staged, float* vmask,
                                            int32_t* route_ptr, int32_t* route_q, void* stream) {
  return ipsr::blend_stage_with_routes_ex(xt, r_masked, inv_norm, ind, mask_idx, flag, B, C, N, M, staged, vmask, route_ptr, route_q,
                                          stream, 0, nullptr);
}

int ipsr::blend_stage_with_routes_ex(const float* xt, const float* r_masked, const float* inv_norm,
                                     const int32_t* ind, const int32_t* mask_idx, const int32_t* flag,
                                     int B, int C, int N, int M, float* staged, float* vmask,
                                     int32_t* route_ptr, int32_t* route_q, void* stream, int ms, const int32_t* mcount) {
  using namespace ipsr;
  const bool routes = route_ptr != nullptr;
  if (M == 0) {
    if (routes) return build_routes_ex(ind, flag, mask_idx, B, N, M, route_ptr, route_q, stream, ms, mcount);
    return IPSR_OK;
  }
  IPSR_REQUIRE(xt && r_masked && inv_norm && ind && mask_idx && staged, IPSR_ERR_INVALID_ARG,
               "ipsr_blend_stage: null pointer");
  IPSR_REQUIRE(!routes || (flag && route_q), IPSR_ERR_INVALID_ARG, "ipsr_blend_stage_with_routes: flag / route_q missing");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0 && M <= N, IPSR_ERR_INVALID_ARG, "ipsr_blend_stage: bad dims");
  IPSR_REQUIRE(C % 32 == 0 && C <= 1024, IPSR_ERR_UNSUPPORTED, "ipsr_blend_stage: C=%d must be a multiple of 32, <= 1024", C);
  IPSR_REQUIRE(!routes || N <= 16384, IPSR_ERR_UNSUPPORTED, "ipsr_blend_stage_with_routes: N=%d > 16384", N);
  StageArgs a;
  a.xt = xt; a.r_masked = r_masked; a.inv_norm = inv_norm; a.ind = ind; a.mask_idx = mask_idx;
  a.B = B; a.C = C; a.N = N; a.M = M; a.nblocks = 0;
  a.staged = staged; a.vmask = vmask;
  a.n_routes = routes ? B : 0; a.flag = flag; a.route_ptr = route_ptr; a.route_q = route_q;
  a.ms = ms; a.mcount = mcount;
  return dispatch_stage(a, as_stream(stream));
}

extern "C" int ipsr_blend_scan(const float* staged, int B, int C, int M,
                               float* y, float* wn, float* wo, void* stream) {
  return ipsr::blend_scan_ex(staged, B, C, M, y, wn, wo, stream, nullptr);
}

int ipsr::blend_scan_ex(const float* staged, int B, int C, int M, float* y, float* wn, float* wo, void* stream,
                        const int32_t* mcount) {
  using namespace ipsr;
  if (M == 0) return IPSR_OK;
  IPSR_REQUIRE(staged && y && wn && wo, IPSR_ERR_INVALID_ARG, "ipsr_blend_scan: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && M > 0, IPSR_ERR_INVALID_ARG, "ipsr_blend_scan: bad dims");
  IPSR_REQUIRE(C % 32 == 0 && C <= 1024, IPSR_ERR_UNSUPPORTED, "ipsr_blend_scan: C=%d must be a multiple of 32, <= 1024", C);
  cudaStream_t st = as_stream(stream);
  switch (scan_block_steps(C)) {
    case 32: return launch_scan<32>(staged, B, C, M, y, wn, wo, st, mcount);
    case 16: return launch_scan<16>(staged, B, C, M, y, wn, wo, st, mcount);
    default: return launch_scan<8>(staged, B, C, M, y, wn, wo, st, mcount);
  }
}

extern "C" int ipsr_paste(const float* x, const float* y, const int32_t* ind, const int32_t* rank,
                          int B, int C, int N, int M, float* out, void* stream) {
  return ipsr::paste_ex(x, y, ind, rank, B, C, N, M, out, stream, 0);
}

extern "C" int ipsr_paste_loss_partials(int B, int C, int N) {
  // one partial per paste CTA: at most one CTA per channel tile and image
  (void)N;
  return B * C;
}

int ipsr::paste_ex(const float* x, const float* y, const int32_t* ind, const int32_t* rank,
                   int B, int C, int N, int M, float* out, void* stream, int ms, const PasteLoss* loss) {
  using namespace ipsr;
  IPSR_REQUIRE(x && ind && rank && out && (M == 0 || y), IPSR_ERR_INVALID_ARG, "ipsr_paste: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0 && B <= 65535, IPSR_ERR_INVALID_ARG, "ipsr_paste: bad dims");
  const int CT = paste_ct(C, N);
  const size_t smem = paste_smem(CT, N, M);
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_paste: N=%d too large", N);
  if (smem + 2048 > 48 * 1024) {   // dynamic + static shared memory above the default limit (set per call: the attribute is per device)
    cudaError_t e = cudaFuncSetAttribute(paste_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "paste smem attribute: %s", cudaGetErrorString(e));
  }
  const int threads = paste_threads(N);
  const int ntiles = (C + CT - 1) / CT;
  const int tpc = tiles_per_cta_for(B, ntiles, paste_slots(smem, threads));
  if (loss) {
    IPSR_REQUIRE(loss->target && loss->mask && loss->partials && loss->ticket && loss->loss && (loss->crit == 0 || loss->crit == 1),
                 IPSR_ERR_INVALID_ARG, "ipsr_paste: fused InnerCos loss needs target, mask, partials, ticket, loss");
    if (smem + 2048 > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(paste_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "paste smem attribute: %s", cudaGetErrorString(e));
    }
    cudaError_t le = launch_pdl(paste_loss_kernel, dim3((ntiles + tpc - 1) / tpc, B), dim3(threads), smem, as_stream(stream), x, y, ind, rank,
                                C, N, M, CT, tpc, out, ms, *loss, (long long)B * C * N);
    IPSR_REQUIRE(le == cudaSuccess, IPSR_ERR_CUDA, "ipsr_paste: launch failed: %s", cudaGetErrorString(le));
    return check_launch("ipsr_paste");
  }
  cudaError_t le = launch_pdl(paste_kernel, dim3((ntiles + tpc - 1) / tpc, B), dim3(threads), smem, as_stream(stream), x, y, ind, rank, C, N,
                              M, CT, tpc, out, ms);
  IPSR_REQUIRE(le == cudaSuccess, IPSR_ERR_CUDA, "ipsr_paste: launch failed: %s", cudaGetErrorString(le));
  return check_launch("ipsr_paste");
}

extern "C" int ipsr_paste_with_bookkeeping(const float* x, const float* y, const int32_t* ind, const int32_t* rank,
                                           const int32_t* flag, const int32_t* mask_idx, const float* wn, const float* wo,
                                           int B, int C, int N, int M, float* out,
                                           int32_t* route_ptr, int32_t* route_q,
                                           int32_t* exc_start, int32_t* exc_cnt, int32_t* exc_l, float* exc_w,
                                           int32_t* exc_total, int exc_cap, void* stream) {
  return ipsr::paste_with_bookkeeping_ex(x, y, ind, rank, flag, mask_idx, wn, wo, B, C, N, M, out, route_ptr, route_q, exc_start,
                                         exc_cnt, exc_l, exc_w, exc_total, exc_cap, stream, 0, nullptr);
}

int ipsr::paste_with_bookkeeping_ex(const float* x, const float* y, const int32_t* ind, const int32_t* rank,
                                    const int32_t* flag, const int32_t* mask_idx, const float* wn, const float* wo,
                                    int B, int C, int N, int M, float* out, int32_t* route_ptr, int32_t* route_q,
                                    int32_t* exc_start, int32_t* exc_cnt, int32_t* exc_l, float* exc_w,
                                    int32_t* exc_total, int exc_cap, void* stream, int ms, const int32_t* mcount) {
  using namespace ipsr;
  const bool routes = route_ptr != nullptr;       // NULL: already built (ipsr_blend_stage_with_routes)
  IPSR_REQUIRE(x && ind && rank && out && (!routes || (flag && route_q)) && (M == 0 || (y && mask_idx)), IPSR_ERR_INVALID_ARG,
               "ipsr_paste_with_bookkeeping: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0, IPSR_ERR_INVALID_ARG, "ipsr_paste_with_bookkeeping: bad dims");
  IPSR_REQUIRE(N <= 16384, IPSR_ERR_UNSUPPORTED, "ipsr_paste_with_bookkeeping: N=%d > 16384", N);
  const bool exc = M > 1;
  if (exc)
    IPSR_REQUIRE(wn && wo && exc_start && exc_cnt && exc_l && exc_w && exc_total && exc_cap > 0, IPSR_ERR_INVALID_ARG,
                 "ipsr_paste_with_bookkeeping: exception buffers missing");
  FusedPasteArgs a;
  a.x = x; a.y = y; a.ind = ind; a.rank = rank; a.out = out;
  a.B = B; a.C = C; a.N = N; a.M = M; a.CT = paste_ct(C, N);
  const int threads = paste_threads(N);
  {
    const int ntiles = (C + a.CT - 1) / a.CT;
    a.tiles_per_cta = tiles_per_cta_for(B, ntiles, paste_slots(paste_smem(a.CT, N, M), threads));
    a.parts = (ntiles + a.tiles_per_cta - 1) / a.tiles_per_cta;
  }
  a.flag = flag; a.mask_idx = mask_idx; a.route_ptr = route_ptr; a.route_q = route_q;
  a.wn = wn; a.wo = wo; a.exc_start = exc_start; a.exc_cnt = exc_cnt; a.exc_l = exc_l; a.exc_w = exc_w;
  a.exc_total = exc_total; a.exc_cap = exc_cap;
  a.ms = ms; a.mcount = mcount;
  a.n_routes = routes ? B : 0;
  a.n_exc = exc ? B : 0;
  size_t smem = paste_smem(a.CT, N, M);
  const size_t smem_routes = (size_t)(2 * N + 1) * sizeof(int), smem_exc = exc_smem_words(N, M) * sizeof(int);
  if (routes && smem_routes > smem) smem = smem_routes;
  if (exc && smem_exc > smem) smem = smem_exc;
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_paste_with_bookkeeping: N=%d too large", N);
  if (smem + 2048 > 48 * 1024) {   // dynamic + static shared memory above the default limit (set per call: the attribute is per device)
    cudaError_t e = cudaFuncSetAttribute(paste_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "paste_fused smem attribute: %s", cudaGetErrorString(e));
  }
  const long long ctas = (long long)a.n_routes + a.n_exc + (long long)B * a.parts;
  IPSR_REQUIRE(ctas <= 0x7FFFFFFFll, IPSR_ERR_UNSUPPORTED, "ipsr_paste_with_bookkeeping: grid too large");
  paste_fused_kernel<<<(unsigned)ctas, threads, smem, as_stream(stream)>>>(a);
  return check_launch("ipsr_paste_with_bookkeeping");
}
