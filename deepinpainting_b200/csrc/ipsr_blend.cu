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
#include "ipsr_bookkeeping.cuh"

namespace ipsr {

// one staged step: u_l [C], X[p_l] [C], then v_l, c_l = <u_l, X[p_{l-1}]> and 2 pad floats (16-byte sized)
__host__ __device__ inline int staged_stride(int C) { return 2 * C + 4; }
// rows of y per image: M rounded up to the unroll depth of the scan (tail steps store into the padding)
__host__ __device__ inline int padded_steps(int M) { return (M + 7) & ~7; }

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
  // matched patch of the PREVIOUS masked position: c_l = <u_l, X[p_{l-1}]> feeds the two-step
  // form of the recurrence used by the scan (a_l = wn_{l-1} <u_l, y_{l-2}> + wo_{l-1} c_l)
  const float* xk = xt + ((size_t)b * N + (l > 0 ? ind[(size_t)b * N + mask_idx[l - 1]] : p)) * C;
  float* su = staged + ((size_t)b * M + l) * staged_stride(C);
  float* sp = su + C;
  float acc = 0.f, cacc = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(xq + c));
    const float4 k = __ldg(reinterpret_cast<const float4*>(xp + c));
    const float4 r = __ldg(reinterpret_cast<const float4*>(rq + c));
    const float4 kp = __ldg(reinterpret_cast<const float4*>(xk + c));
    cacc = fmaf(__fmul_rn(a.x, inv_q), kp.x, cacc);
    cacc = fmaf(__fmul_rn(a.y, inv_q), kp.y, cacc);
    cacc = fmaf(__fmul_rn(a.z, inv_q), kp.z, cacc);
    cacc = fmaf(__fmul_rn(a.w, inv_q), kp.w, cacc);
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
  cacc = warp_sum(cacc);
  if (lane == 0) {
    *reinterpret_cast<float4*>(su + 2 * C) = make_float4(acc, cacc, 0.f, 0.f);
    if (vmask) vmask[(size_t)b * M + l] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// scan
// ---------------------------------------------------------------------------------------------
constexpr int kScanStages = 4;      // ring depth
constexpr int kScanSteps = 8;       // recurrence steps per ring stage (the inner loop is fully unrolled)

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// VPL = C / 32 values per lane; lane owns channels lane + 32*i (conflict-free smem reads, 128-byte
// coalesced y stores).  One warp per image; one ring stage = 8 consecutive steps = contiguous bytes of
// `staged`, fetched by one bulk copy.
//
// The warp is alone on its scheduler, so every dependent instruction costs its full latency: the loop
// body is branch-free straight-line code (8 steps unrolled) and uses the two-step form of the
// recurrence so that independent work fills the latency of the 32-lane reductions.
//   reference (IPSRFunction.py:116-122):  a_l = <u_l, y_{l-1}>,  y_l = wn_l y_{l-1} + wo_l X[p_l]
//   by linearity:  a_l = wn_{l-1} <u_l, y_{l-2}> + wo_{l-1} c_l,   c_l = <u_l, X[p_{l-1}]>  (staged)
// so the reduction for step l+1 (needs only y_{l-1}) overlaps the scalar chain of step l.  With
// y_{-1} = 0, wn_0 = 0, wo_0 = 1 steps 0 and 1 are regular.  Differences to the one-step form are
// rounding-level (and wn, wo use a * rcp(a+v): <= 2 ulp from the reference's IEEE divisions).
template <int VPL>
__global__ void __launch_bounds__(32)
blend_scan_kernel(const float* __restrict__ staged, int M,
                  float* __restrict__ y, float* __restrict__ wn_out, float* __restrict__ wo_out) {
  extern __shared__ __align__(128) uint8_t scan_smem[];
  __shared__ __align__(8) unsigned long long bars[kScanStages];
  constexpr int C = VPL * 32;
  constexpr int kStride = 2 * C + 4;
  constexpr uint32_t kStepBytes = (uint32_t)kStride * sizeof(float);
  constexpr uint32_t kStageBytes = kStepBytes * kScanSteps;
  const int b = blockIdx.x;
  const int lane = threadIdx.x;
  const uint32_t ring = smem_u32(scan_smem);
  const float* src = staged + (size_t)b * M * kStride;
  float* yb = y + (size_t)b * padded_steps(M) * C;
  float* wnb = wn_out + (size_t)b * M;
  float* wob = wo_out + (size_t)b * M;
  const int nchunks = (M + kScanSteps - 1) / kScanSteps;

  if (lane == 0) {
    for (int s = 0; s < kScanStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int chunk) {
    const int s = chunk % kScanStages;
    const int l0 = chunk * kScanSteps;
    const uint32_t bytes = kStepBytes * (uint32_t)min(kScanSteps, M - l0);
    mbar_expect_tx(smem_u32(&bars[s]), bytes);
    bulk_g2s(ring + (uint32_t)s * kStageBytes, src + (size_t)l0 * kStride, bytes, smem_u32(&bars[s]));
  };
  if (lane == 0)
    for (int ch = 0; ch < min(nchunks, kScanStages); ++ch) issue(ch);

  float y1[VPL];                       // y_{l-1}
#pragma unroll
  for (int i = 0; i < VPL; ++i) y1[i] = 0.f;
  float d = 0.f;                       // <u_l, y_{l-2}>
  float wn_prev = 0.f, wo_prev = 1.f;
  mbar_wait(smem_u32(&bars[0]), 0);
  for (int ch = 0; ch < nchunks; ++ch) {
    const int s = ch % kScanStages;
    const float* st = reinterpret_cast<const float*>(scan_smem + (size_t)s * kStageBytes);
    // the last step of this chunk looks one step ahead: the next stage must have landed too
    const float* st_next = st;
    if (ch + 1 < nchunks) {
      const int s2 = (ch + 1) % kScanStages;
      mbar_wait(smem_u32(&bars[s2]), (uint32_t)((ch + 1) / kScanStages) & 1u);
      st_next = reinterpret_cast<const float*>(scan_smem + (size_t)s2 * kStageBytes);
    }
    const int l0 = ch * kScanSteps;
#pragma unroll
    for (int j = 0; j < kScanSteps; ++j) {
      const int l = l0 + j;
      const bool live = l < M;                               // tail steps compute on stale smem into y's padding
      const float* u = st + j * kStride;
      const float* k = u + C;
      const float* un = (j + 1 < kScanSteps) ? (u + kStride) : st_next;
      // (1) d_{l+1} = <u_{l+1}, y_{l-1}>: independent of this step's weights
      float p0 = 0.f, p1 = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; i += 2) {
        p0 = fmaf(un[lane + 32 * i], y1[i], p0);
        if (i + 1 < VPL) p1 = fmaf(un[lane + 32 * (i + 1)], y1[i + 1], p1);
      }
      const float dn = warp_sum(p0 + p1);
      // (2) scalar chain of step l
      const float2 vc = *reinterpret_cast<const float2*>(u + 2 * C);     // v_l, c_l
      const float a = fmaf(wn_prev, d, __fmul_rn(wo_prev, vc.y));
      const float r = rcp_approx(__fadd_rn(a, vc.x));         // no clamp: inf / nan propagate      :120
      const float wn = (l == 0) ? 0.f : __fmul_rn(a, r);      // first masked patch: plain copy      :98-101
      const float wo = (l == 0) ? 1.f : __fmul_rn(vc.x, r);   //                                       :121
      // (3) y_l = wn y_{l-1} + wo X[p_l]: two rounded products, one sum                              :122
#pragma unroll
      for (int i = 0; i < VPL; ++i) y1[i] = __fadd_rn(__fmul_rn(wn, y1[i]), __fmul_rn(wo, k[lane + 32 * i]));
      if (live && lane == 0) {
        wnb[l] = wn;
        wob[l] = wo;
      }
#pragma unroll
      for (int i = 0; i < VPL; ++i) yb[(size_t)l * C + lane + 32 * i] = y1[i];    // y is padded to 8 steps
      d = dn;
      wn_prev = wn;
      wo_prev = wo;
    }
    __syncwarp();                                          // every lane is done reading stage s
    if (lane == 0 && ch + kScanStages < nchunks) issue(ch + kScanStages);
  }
}

// ---------------------------------------------------------------------------------------------
// paste
// ---------------------------------------------------------------------------------------------
// A CTA stages CT channel rows of x[b] (N floats each) in shared memory and writes the CT output
// rows: coalesced reads, coalesced writes, gather inside the SM.
__device__ __forceinline__ void
paste_cta(int cx, int b, float* rows, const float* __restrict__ x, const float* __restrict__ y,
          const int* __restrict__ ind, const int* __restrict__ rank, int C, int N, int M, int CT,
          float* __restrict__ out) {
  const int c0 = cx * CT;
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
      const float* yr = y + ((size_t)b * padded_steps(M) + l) * C + c0;
      for (int ch = 0; ch < ct; ++ch) ob[(size_t)ch * N + q] = __ldg(yr + ch);
    }
  }
}

// grid = (C / CT, B)
__global__ void __launch_bounds__(256)
paste_kernel(const float* __restrict__ x, const float* __restrict__ y, const int* __restrict__ ind,
             const int* __restrict__ rank, int C, int N, int M, int CT, float* __restrict__ out) {
  extern __shared__ __align__(16) float rows[];          // [CT][N]
  paste_cta(blockIdx.x, blockIdx.y, rows, x, y, ind, rank, C, N, M, CT, out);
}

// The paste and the two bookkeeping builders of the backward are independent once the scan is done:
// one launch, block ranges = [routes: B CTAs][exceptions: B CTAs][paste: B * C/CT CTAs].
// The latency-bound bookkeeping CTAs are scheduled first and overlap the bandwidth-bound paste.
struct FusedPasteArgs {
  const float* x; const float* y; const int* ind; const int* rank; float* out;
  int B, C, N, M, CT, ctiles;
  const int* flag; const int* mask_idx; int* route_ptr; int* route_q;
  const float* wn; const float* wo; int* exc_start; int* exc_cnt; int* exc_l; float* exc_w; int* exc_total; int exc_cap;
  int n_routes, n_exc, exc_per_img;
};

__global__ void __launch_bounds__(256) paste_fused_kernel(const FusedPasteArgs a) {
  extern __shared__ __align__(16) float fsm[];
  int blk = blockIdx.x;
  if (blk < a.n_routes) {
    build_routes_cta(blk, reinterpret_cast<int*>(fsm), a.ind, a.flag, a.mask_idx, a.N, a.M, a.route_ptr, a.route_q);
    return;
  }
  blk -= a.n_routes;
  if (blk < a.n_exc) {
    build_exceptions_cta(blk, fsm, a.ind, a.mask_idx, a.wn, a.wo, a.N, a.M, a.exc_start, a.exc_cnt, a.exc_l, a.exc_w,
                         a.exc_total, a.exc_cap);
    return;
  }
  blk -= a.n_exc;
  paste_cta(blk % a.ctiles, blk / a.ctiles, fsm, a.x, a.y, a.ind, a.rank, a.C, a.N, a.M, a.CT, a.out);
}

static int paste_ct(int C, int N) {
  // channel rows per CTA: ~64 KiB of shared memory, at least 1 row
  int CT = (int)((64 * 1024) / ((size_t)N * sizeof(float)));
  if (CT < 1) CT = 1;
  if (CT > 16) CT = 16;
  if (CT > C) CT = C;
  return CT;
}

}  // namespace ipsr

extern "C" int ipsr_staged_stride(int C) { return ipsr::staged_stride(C); }
extern "C" int ipsr_padded_steps(int M) { return ipsr::padded_steps(M); }

extern "C" int ipsr_blend_stage(const float* xt, const float* r_masked, const float* inv_norm,
                                const int32_t* ind, const int32_t* mask_idx, int B, int C, int N, int M,
                                float* staged, float* vmask, void* stream) {
  using namespace ipsr;
  if (M == 0) return IPSR_OK;
  IPSR_REQUIRE(xt && r_masked && inv_norm && ind && mask_idx && staged, IPSR_ERR_INVALID_ARG,
               "ipsr_blend_stage: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0 && M <= N, IPSR_ERR_INVALID_ARG, "ipsr_blend_stage: bad dims");
  IPSR_REQUIRE(C % 4 == 0, IPSR_ERR_UNSUPPORTED, "ipsr_blend_stage: C=%d must be a multiple of 4", C);
  blend_stage_kernel<<<dim3((M + 7) / 8, B), 256, 0, as_stream(stream)>>>(xt, r_masked, inv_norm, ind, mask_idx, C, N, M,
                                                                          staged, vmask);
  return check_launch("ipsr_blend_stage");
}

namespace ipsr {
template <int VPL>
static int launch_scan(const float* staged, int B, int M, float* y, float* wn, float* wo, cudaStream_t st) {
  constexpr int C = VPL * 32;
  const size_t smem = (size_t)kScanStages * kScanSteps * staged_stride(C) * sizeof(float);
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_blend_scan: C=%d too large", C);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(blend_scan_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "blend_scan smem attribute: %s", cudaGetErrorString(e));
  }
  blend_scan_kernel<VPL><<<B, 32, smem, st>>>(staged, M, y, wn, wo);
  return check_launch("ipsr_blend_scan");
}
}  // namespace ipsr

extern "C" int ipsr_blend_scan(const float* staged, int B, int C, int M,
                               float* y, float* wn, float* wo, void* stream) {
  using namespace ipsr;
  if (M == 0) return IPSR_OK;
  IPSR_REQUIRE(staged && y && wn && wo, IPSR_ERR_INVALID_ARG, "ipsr_blend_scan: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && M > 0, IPSR_ERR_INVALID_ARG, "ipsr_blend_scan: bad dims");
  cudaStream_t st = as_stream(stream);
  IPSR_REQUIRE(C % 32 == 0, IPSR_ERR_UNSUPPORTED, "ipsr_blend_scan: C=%d must be a multiple of 32", C);
  switch (C / 32) {
    case 1: return launch_scan<1>(staged, B, M, y, wn, wo, st);
    case 2: return launch_scan<2>(staged, B, M, y, wn, wo, st);
    case 3: return launch_scan<3>(staged, B, M, y, wn, wo, st);
    case 4: return launch_scan<4>(staged, B, M, y, wn, wo, st);
    case 6: return launch_scan<6>(staged, B, M, y, wn, wo, st);
    case 8: return launch_scan<8>(staged, B, M, y, wn, wo, st);
    case 12: return launch_scan<12>(staged, B, M, y, wn, wo, st);
    case 16: return launch_scan<16>(staged, B, M, y, wn, wo, st);
    case 24: return launch_scan<24>(staged, B, M, y, wn, wo, st);
    case 32: return launch_scan<32>(staged, B, M, y, wn, wo, st);
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
  const int CT = paste_ct(C, N);
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

extern "C" int ipsr_paste_with_bookkeeping(const float* x, const float* y, const int32_t* ind, const int32_t* rank,
                                           const int32_t* flag, const int32_t* mask_idx, const float* wn, const float* wo,
                                           int B, int C, int N, int M, float* out,
                                           int32_t* route_ptr, int32_t* route_q,
                                           int32_t* exc_start, int32_t* exc_cnt, int32_t* exc_l, float* exc_w,
                                           int32_t* exc_total, int exc_cap, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(x && ind && rank && out && flag && route_ptr && route_q && (M == 0 || (y && mask_idx)), IPSR_ERR_INVALID_ARG,
               "ipsr_paste_with_bookkeeping: null pointer");
  IPSR_REQUIRE(B > 0 && C > 0 && N > 0, IPSR_ERR_INVALID_ARG, "ipsr_paste_with_bookkeeping: bad dims");
  IPSR_REQUIRE(N <= 16384, IPSR_ERR_UNSUPPORTED, "ipsr_paste_with_bookkeeping: N=%d > 16384", N);
  const bool exc = M > 1;
  if (exc)
    IPSR_REQUIRE(wn && wo && exc_start && exc_cnt && exc_l && exc_w && exc_total && exc_cap > 0, IPSR_ERR_INVALID_ARG,
                 "ipsr_paste_with_bookkeeping: exception buffers missing");
  FusedPasteArgs a;
  a.x = x; a.y = y; a.ind = ind; a.rank = rank; a.out = out;
  a.B = B; a.C = C; a.N = N; a.M = M; a.CT = paste_ct(C, N); a.ctiles = (C + a.CT - 1) / a.CT;
  a.flag = flag; a.mask_idx = mask_idx; a.route_ptr = route_ptr; a.route_q = route_q;
  a.wn = wn; a.wo = wo; a.exc_start = exc_start; a.exc_cnt = exc_cnt; a.exc_l = exc_l; a.exc_w = exc_w;
  a.exc_total = exc_total; a.exc_cap = exc_cap;
  a.n_routes = B;
  a.exc_per_img = exc ? 1 : 0;
  a.n_exc = B * a.exc_per_img;
  size_t smem = (size_t)a.CT * N * sizeof(float);
  const size_t smem_routes = (size_t)(2 * N + 1) * sizeof(int), smem_exc = ((size_t)N + 3 * kExcChunk) * sizeof(int);
  if (smem_routes > smem) smem = smem_routes;
  if (smem_exc > smem) smem = smem_exc;
  IPSR_REQUIRE(smem <= 227 * 1024, IPSR_ERR_UNSUPPORTED, "ipsr_paste_with_bookkeeping: N=%d too large", N);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(paste_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "paste_fused smem attribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  const long long ctas = (long long)a.n_routes + a.n_exc + (long long)B * a.ctiles;
  IPSR_REQUIRE(ctas <= 0x7FFFFFFFll, IPSR_ERR_UNSUPPORTED, "ipsr_paste_with_bookkeeping: grid too large");
  paste_fused_kernel<<<(unsigned)ctas, 256, smem, as_stream(stream)>>>(a);
  return check_launch("ipsr_paste_with_bookkeeping");
}
