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
