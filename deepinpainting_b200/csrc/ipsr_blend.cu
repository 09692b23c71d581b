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
#include "ipsr_common.cuh"

namespace ipsr {

// ---------------------------------------------------------------------------------------------
// stage
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
blend_stage_kernel(const float* __restrict__ xt, const float* __restrict__ r_masked, const float* __restrict__ inv_norm,
                   const int* __restrict__ ind, const int* __restrict__ mask_idx, int C, int N, int M,
                   float* __restrict__ staged, float* __restrict__ vmask) {
  const int b = blockIdx.y;
  const int l = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (l >= M) return;
  const int q = mask_idx[l];
  const int p = ind[(size_t)b * N + q];
  const float inv_q = inv_norm[(size_t)b * N + q];
  const float inv_p = inv_norm[(size_t)b * N + p];
  const float* xq = xt + ((size_t)b * N + q) * C;
  const float* xp = xt + ((size_t)b * N + p) * C;
  const float* rq = r_masked + ((size_t)b * M + l) * C;
  float* su = staged + (((size_t)b * M + l) * 2) * C;
  float* sp = su + C;
  float acc = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(xq + c));
    const float4 k = __ldg(reinterpret_cast<const float4*>(xp + c));
    const float4 r = __ldg(reinterpret_cast<const float4*>(rq + c));
    // u = little_value * (1/(norm+1e-8))                          IPSRFunction.py:109
    *reinterpret_cast<float4*>(su + c) =
        make_float4(__fmul_rn(a.x, inv_q), __fmul_rn(a.y, inv_q), __fmul_rn(a.z, inv_q), __fmul_rn(a.w, inv_q));
    *reinterpret_cast<float4*>(sp + c) = k;
    // v = <R[q], Xn[p]> with Xn = fl(X * inv) as the encoder weights hold it (NPS:40)
    acc = fmaf(r.x, __fmul_rn(k.x, inv_p), acc);
    acc = fmaf(r.y, __fmul_rn(k.y, inv_p), acc);
    acc = fmaf(r.z, __fmul_rn(k.z, inv_p), acc);
    acc = fmaf(r.w, __fmul_rn(k.w, inv_p), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) vmask[(size_t)b * M + l] = acc;
}

// ---------------------------------------------------------------------------------------------
// scan
// ---------------------------------------------------------------------------------------------
constexpr int kScanStages = 8;

// VPL = C / 32 values per lane; lane owns channels lane + 32*i (conflict-free smem reads,
// 128-byte coalesced y stores).  One stage = `steps_per_stage` consecutive steps = contiguous bytes
// of `staged`, fetched by one bulk copy.
template <int VPL>
__global__ void __launch_bounds__(32)
blend_scan_kernel(const float* __restrict__ staged, const float* __restrict__ vmask, int M, int steps_per_stage,
                  float* __restrict__ y, float* __restrict__ wn_out, float* __restrict__ wo_out) {
  extern __shared__ __align__(128) uint8_t scan_smem[];
  __shared__ __align__(8) unsigned long long bars[kScanStages];
  constexpr int C = VPL * 32;
  const int b = blockIdx.x;
  const int lane = threadIdx.x;
  const uint32_t ring = smem_u32(scan_smem);
  const uint32_t step_bytes = 2u * C * sizeof(float);
  const uint32_t stage_bytes = step_bytes * steps_per_stage;
  const float* src = staged + (size_t)b * M * 2 * C;
  const float* vm = vmask + (size_t)b * M;
  float* yb = y + (size_t)b * M * C;
  const int nchunks = (M + steps_per_stage - 1) / steps_per_stage;

  if (lane == 0) {
    for (int s = 0; s < kScanStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int chunk) {
    const int s = chunk % kScanStages;
    const int l0 = chunk * steps_per_stage;
    const uint32_t bytes = step_bytes * (uint32_t)min(steps_per_stage, M - l0);
    mbar_expect_tx(smem_u32(&bars[s]), bytes);
    bulk_g2s(ring + (uint32_t)s * stage_bytes, src + (size_t)l0 * 2 * C, bytes, smem_u32(&bars[s]));
  };
  if (lane == 0)
    for (int ch = 0; ch < min(nchunks, kScanStages); ++ch) issue(ch);

  float yv[VPL];
  for (int ch = 0; ch < nchunks; ++ch) {
    const int s = ch % kScanStages;
    mbar_wait(smem_u32(&bars[s]), (uint32_t)(ch / kScanStages) & 1u);
    const float* st = reinterpret_cast<const float*>(scan_smem + (size_t)s * stage_bytes);
    const int l0 = ch * steps_per_stage;
    const int nl = min(steps_per_stage, M - l0);
    for (int j = 0; j < nl; ++j) {
      const int l = l0 + j;
      const float* u = st + (size_t)j * 2 * C;
      const float* k = u + C;
      float uv[VPL], kv[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        uv[i] = u[lane + 32 * i];
        kv[i] = k[lane + 32 * i];
      }
      const float v = __ldg(vm + l);
      if (l == 0) {
        // first masked patch: plain copy of the matched patch          IPSRFunction.py:98-101
#pragma unroll
        for (int i = 0; i < VPL; ++i) yv[i] = kv[i];
        if (lane == 0) {
          wn_out[(size_t)b * M] = 0.f;
          wo_out[(size_t)b * M] = 1.f;
        }
      } else {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; i += 2) {
          a0 = fmaf(uv[i], yv[i], a0);
          if (i + 1 < VPL) a1 = fmaf(uv[i + 1], yv[i + 1], a1);
        }
        const float a = warp_sum(a0 + a1);                 // 1x1 conv == dot            :116
        const float den = __fadd_rn(a, v);
        const float wn = __fdiv_rn(a, den);                // no clamp, inf/nan propagate :120
        const float wo = __fdiv_rn(v, den);                //                              :121
#pragma unroll
        for (int i = 0; i < VPL; ++i)                       // two rounded products, one sum :122
          yv[i] = __fadd_rn(__fmul_rn(wn, yv[i]), __fmul_rn(wo, kv[i]));
        if (lane == 0) {
          wn_out[(size_t)b * M + l] = wn;
          wo_out[(size_t)b * M + l] = wo;
        }
      }
#pragma unroll
      for (int i = 0; i < VPL; ++i) yb[(size_t)l * C + lane + 32 * i] = yv[i];
    }
    __syncwarp();                                          // every lane is done reading stage s
    if (lane == 0 && ch + kScanStages < nchunks) issue(ch + kScanStages);
  }
}

// ---------------------------------------------------------------------------------------------
// paste
// ---------------------------------------------------------------------------------------------
// grid = (C / CT, B); a CTA stages CT channel rows of x[b] (N floats each) in shared memory and
// writes the CT output rows: coalesced reads, coalesced writes, gather inside the SM.
__global__ void __launch_bounds__(256)
paste_kernel(const float* __restrict__ x, const float* __restrict__ y, const int* __restrict__ ind,
             const int* __restrict__ rank, int C, int N, int M, int CT, float* __restrict__ out) {
  extern __shared__ __align__(16) float rows[];          // [CT][N]
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * CT;
  const int ct = min(CT, C - c0);
  const float* xb = x + ((size_t)b * C + c0) * N;
  float* ob = out + ((size_t)b * C + c0) * N;
  const int total = ct * N;
  if ((N & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(xb);
    float4* d4 = reinterpret_cast<float4*>(rows);
    for (int i = threadIdx.x; i < total / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) rows[i] = __ldg(xb + i);
  }
  __syncthreads();
  const int* indb = ind + (size_t)b * N;
  for (int q = threadIdx.x; q < N; q += blockDim.x) {
    const int l = rank[q];
    if (l < 0) {
      const int p = indb[q];
      for (int ch = 0; ch < ct; ++ch) ob[(size_t)ch * N + q] = rows[ch * N + p];
    } else {
      const float* yr = y + ((size_t)b * M + l) * C + c0;
      for (int ch = 0; ch < ct; ++ch) ob[(size_t)ch * N + q] = __ldg(yr + ch);
    }
  }
}

}  // namespace ipsr

extern "C" int ipsr_blend_stage(const float* xt, const float* r_masked, const float* inv_norm,
                                const int32_t* ind, const int32_t* mask_idx, int B, int C, int N, int M,
                                float* staged, float* vmask, void* stream) {
  using namespace ipsr;
  if (M == 0) return IPSR_OK;
  IPSR_REQUIRE(xt && r_masked && inv_norm && ind && mask_idx && staged && vmask, IPSR_ERR_INVALID_ARG,
               "ipsr_blend_stage: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0 && M <= N, IPSR_ERR_INVALID_ARG, "ipsr_blend_stage: bad dims");
  IPSR_REQUIRE(C % 4 == 0, IPSR_ERR_UNSUPPORTED, "ipsr_blend_stage: C=%d must be a multiple of 4", C);
  blend_stage_kernel<<<dim3((M + 7) / 8, B), 256, 0, as_stream(stream)>>>(xt, r_masked, inv_norm, ind, mask_idx, C, N, M,
                                                                          staged, vmask);
  return check_launch("ipsr_blend_stage");
}

namespace ipsr {
template <int VPL>
static int launch_scan(const float* staged, const float* vmask, int B, int M, float* y, float* wn, float* wo,
                       cudaStream_t st) {
  constexpr int C = VPL * 32;
  const size_t step_bytes = 2 * (size_t)C * sizeof(float);
  int sps = (int)(8192 / step_bytes);                     // ~8 KiB per stage
  if (sps < 1) sps = 1;
  const size_t smem = (size_t)kScanStages * sps * step_bytes;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(blend_scan_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "blend_scan smem attribute: %s", cudaGetErrorString(e));
  }
  blend_scan_kernel<VPL><<<B, 32, smem, st>>>(staged, vmask, M, sps, y, wn, wo);
  return check_launch("ipsr_blend_scan");
}
}  // namespace ipsr

extern "C" int ipsr_blend_scan(const float* staged, const float* vmask, int B, int C, int M,
                               float* y, float* wn, float* wo, void* stream) {
  using namespace ipsr;
  if (M == 0) return IPSR_OK;
  IPSR_REQUIRE(staged && vmask && y && wn && wo, IPSR_ERR_INVALID_ARG, "ipsr_blend_scan: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && M > 0, IPSR_ERR_INVALID_ARG, "ipsr_blend_scan: bad dims");
  cudaStream_t st = as_stream(stream);
  IPSR_REQUIRE(C % 32 == 0, IPSR_ERR_UNSUPPORTED, "ipsr_blend_scan: C=%d must be a multiple of 32", C);
  switch (C / 32) {
    case 1: return launch_scan<1>(staged, vmask, B, M, y, wn, wo, st);
    case 2: return launch_scan<2>(staged, vmask, B, M, y, wn, wo, st);
    case 3: return launch_scan<3>(staged, vmask, B, M, y, wn, wo, st);
    case 4: return launch_scan<4>(staged, vmask, B, M, y, wn, wo, st);
    case 6: return launch_scan<6>(staged, vmask, B, M, y, wn, wo, st);
    case 8: return launch_scan<8>(staged, vmask, B, M, y, wn, wo, st);
    case 12: return launch_scan<12>(staged, vmask, B, M, y, wn, wo, st);
    case 16: return launch_scan<16>(staged, vmask, B, M, y, wn, wo, st);
    case 24: return launch_scan<24>(staged, vmask, B, M, y, wn, wo, st);
    case 32: return launch_scan<32>(staged, vmask, B, M, y, wn, wo, st);
    default: break;
  }
  set_error("ipsr_blend_scan: C=%d not supported (C/32 must be one of 1,2,3,4,6,8,12,16,24,32)", C);
  return IPSR_ERR_UNSUPPORTED;
}

extern "C" int ipsr_paste(const float* x, const float* y, const int32_t* ind, const int32_t* rank,
                          int B, int C, int N, int M, float* out, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(x && ind && rank && out && (M == 0 || y), IPSR_ERR_INVALID_ARG, "ipsr_paste: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0 && B <= 65535, IPSR_ERR_INVALID_ARG, "ipsr_paste: bad dims");
  // channel rows per CTA: ~64 KiB of shared memory, at least 1 row
  int CT = (int)((64 * 1024) / ((size_t)N * sizeof(float)));
  if (CT < 1) CT = 1;
  if (CT > 16) CT = 16;
  if (CT > C) CT = C;
  const size_t smem = (size_t)CT * N * sizeof(float);
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_paste: N=%d too large", N);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(paste_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "paste smem attribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  paste_kernel<<<dim3((C + CT - 1) / CT, B), 256, smem, as_stream(stream)>>>(x, y, ind, rank, C, N, M, CT, out);
  return check_launch("ipsr_paste");
}
