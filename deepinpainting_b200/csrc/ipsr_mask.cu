// Mask helpers on device: feature mask (util/util.py:68-84) and flag / index vectors
// (util/util.py:88-147).  Both run once per new mask, replace host python loops (one of them
// O(N^2)), and keep the layer free of host synchronisation.
#include "ipsr_common.cuh"

namespace ipsr {

// One 4x4 / stride 2 / pad 1 box-filter layer on integer counts.  The reference convolves with
// weights 1/16 in fp32; every partial sum is a multiple of 16^-L below 2^24, hence exact, so the
// integer count / 16^L is the same number.
template <typename TIn>
__global__ void box4s2_kernel(const TIn* __restrict__ in, int Hi, int Wi, int* __restrict__ out, int Ho, int Wo) {
  in += (size_t)blockIdx.y * Hi * Wi;                      // blockIdx.y: image of the batch
  out += (size_t)blockIdx.y * Ho * Wo;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ho * Wo) return;
  const int oy = i / Wo, ox = i % Wo;
  int acc = 0;
#pragma unroll
  for (int dy = 0; dy < 4; ++dy) {
    const int y = oy * 2 - 1 + dy;
    if (y < 0 || y >= Hi) continue;
#pragma unroll
    for (int dx = 0; dx < 4; ++dx) {
      const int x = ox * 2 - 1 + dx;
      if (x < 0 || x >= Wi) continue;
      acc += (int)in[(size_t)y * Wi + x];
    }
  }
  out[i] = acc;
}

__global__ void threshold_kernel(const int* __restrict__ in, int n, float scale, float threshold,
                                 uint8_t* __restrict__ out) {
  in += (size_t)blockIdx.y * n;
  out += (size_t)blockIdx.y * n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (__fmul_rn((float)in[i], scale) > threshold) ? 1 : 0;   // `> threshold` util/util.py:82
}

__global__ void threshold_u8_kernel(const uint8_t* __restrict__ in, int n, float threshold, uint8_t* __restrict__ out) {
  in += (size_t)blockIdx.y * n;
  out += (size_t)blockIdx.y * n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = ((float)in[i] > threshold) ? 1 : 0;
}

// flag / mask_idx / rank in one CTA: window sums + stable compaction by ballot prefix sums.
__global__ void __launch_bounds__(1024)
build_flags_kernel(const uint8_t* __restrict__ feat, int H, int W, int k, int stride, int thred, int nH, int nW,
                   int* __restrict__ flag, int* __restrict__ mask_idx, int* __restrict__ rank, int* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int P = nH * nW;
  feat += (size_t)blockIdx.x * H * W;                      // blockIdx.x: image of the batch (rows of P entries each)
  flag += (size_t)blockIdx.x * P;
  mask_idx += (size_t)blockIdx.x * P;
  rank += (size_t)blockIdx.x * P;
  count += blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < P; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int f = 0;
    if (i < P) {
      const int h = i / nW, w = i % nW;
      int s = 0;
      for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) s += (int)feat[(size_t)(h * stride + dy) * W + (w * stride + dx)];
      f = (s >= thred) ? 1 : 0;                                    // util/util.py:131
      flag[i] = f;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = carry;
    for (int w2 = 0; w2 < warp; ++w2) off += warp_tot[w2];
    if (i < P) {
      if (f) {
        mask_idx[off + before] = i;
        rank[i] = off + before;
      } else {
        rank[i] = -1;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) t += warp_tot[w2];
      carry += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

}  // namespace ipsr

extern "C" int ipsr_feat_mask(const uint8_t* mask_u8, int S_h, int S_w, int conv_layers, float threshold,
                              uint8_t* feat_u8, int32_t* scratch_i32, void* stream) {
  return ipsr_feat_mask_batch(mask_u8, 1, S_h, S_w, conv_layers, threshold, feat_u8, scratch_i32, stream);
}

extern "C" int ipsr_feat_mask_batch(const uint8_t* mask_u8, int B, int S_h, int S_w, int conv_layers, float threshold,
                                    uint8_t* feat_u8, int32_t* scratch_i32, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(mask_u8 && feat_u8, IPSR_ERR_INVALID_ARG, "ipsr_feat_mask: null pointer");
  IPSR_REQUIRE(B > 0 && B <= 65535, IPSR_ERR_INVALID_ARG, "ipsr_feat_mask: bad batch %d", B);
  IPSR_REQUIRE(S_h > 0 && S_w > 0 && conv_layers >= 0 && conv_layers <= 5, IPSR_ERR_INVALID_ARG,
               "ipsr_feat_mask: bad arguments S=%dx%d layers=%d", S_h, S_w, conv_layers);
  IPSR_REQUIRE((S_h % (1 << conv_layers)) == 0 && (S_w % (1 << conv_layers)) == 0, IPSR_ERR_UNSUPPORTED,
               "ipsr_feat_mask: mask %dx%d must be a multiple of 2^%d", S_h, S_w, conv_layers);
  cudaStream_t st = as_stream(stream);
  if (conv_layers == 0) {
    const int n = S_h * S_w;
    threshold_u8_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(mask_u8, n, threshold, feat_u8);
    return check_launch("ipsr_feat_mask");
  }
  IPSR_REQUIRE(scratch_i32, IPSR_ERR_INVALID_ARG, "ipsr_feat_mask: scratch is null");
  int Hi = S_h, Wi = S_w;
  int* buf[2] = {scratch_i32, scratch_i32 + (size_t)B * (S_h / 2) * (S_w / 2)};
  const int* cur = nullptr;
  for (int l = 0; l < conv_layers; ++l) {
    const int Ho = Hi / 2, Wo = Wi / 2;
    int* dst = buf[l & 1];
    const int n = Ho * Wo;
    if (l == 0) box4s2_kernel<uint8_t><<<dim3((n + 255) / 256, B), 256, 0, st>>>(mask_u8, Hi, Wi, dst, Ho, Wo);
    else box4s2_kernel<int><<<dim3((n + 255) / 256, B), 256, 0, st>>>(cur, Hi, Wi, dst, Ho, Wo);
    cur = dst;
    Hi = Ho;
    Wi = Wo;
  }
  const int n = Hi * Wi;
  const float scale = ldexpf(1.0f, -4 * conv_layers);
  threshold_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(cur, n, scale, threshold, feat_u8);
  return check_launch("ipsr_feat_mask");
}

extern "C" int ipsr_build_flags(const uint8_t* feat_u8, int H, int W, int patch, int stride, int mask_thred,
                                int32_t* flag_i32, int32_t* mask_idx_i32, int32_t* rank_i32, int32_t* count_i32,
                                void* stream) {
  return ipsr_build_flags_batch(feat_u8, 1, H, W, patch, stride, mask_thred, flag_i32, mask_idx_i32, rank_i32, count_i32, stream);
}

extern "C" int ipsr_build_flags_batch(const uint8_t* feat_u8, int B, int H, int W, int patch, int stride, int mask_thred,
                                      int32_t* flag_i32, int32_t* mask_idx_i32, int32_t* rank_i32, int32_t* count_i32,
                                      void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(B > 0, IPSR_ERR_INVALID_ARG, "ipsr_build_flags: bad batch %d", B);
  IPSR_REQUIRE(feat_u8 && flag_i32 && mask_idx_i32 && rank_i32 && count_i32, IPSR_ERR_INVALID_ARG,
               "ipsr_build_flags: null pointer");
  IPSR_REQUIRE(H > 0 && W > 0 && patch > 0 && stride > 0 && patch <= H && patch <= W, IPSR_ERR_INVALID_ARG,
               "ipsr_build_flags: bad geometry H=%d W=%d k=%d s=%d", H, W, patch, stride);
  const int nH = (H - patch) / stride + 1, nW = (W - patch) / stride + 1;
  IPSR_REQUIRE((long long)nH * nW <= 65536, IPSR_ERR_UNSUPPORTED, "ipsr_build_flags: %d positions > 65536", nH * nW);
  build_flags_kernel<<<B, 1024, 0, as_stream(stream)>>>(feat_u8, H, W, patch, stride, mask_thred, nH, nW,
                                                         flag_i32, mask_idx_i32, rank_i32, count_i32);
  return check_launch("ipsr_build_flags");
}
