// InnerCos / InnerCos2 side loss (models/InnerCos.py:30-36, models/InnerCos2.py:34-41):
//   loss = mean( crit( x[:, :c_limit] * mask * strength - target ) ),  crit = square | abs.
// The reference runs mul, mul, MSELoss as three passes over B*C*N; here it is one fused, vectorised
// read of x and target with a deterministic two-level reduction (fixed grid, ticket-elected last CTA
// sums the partials in index order, in double).
#include "ipsr_common.cuh"

namespace ipsr {

constexpr int kIcThreads = 256;
constexpr int kIcMaxBlocks = 1024;

__global__ void __launch_bounds__(kIcThreads)
innercos_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ target,
                    int C_total, int c_limit, int N, float strength, int crit, long long total,
                    float* __restrict__ partials, unsigned int* __restrict__ ticket, float* __restrict__ loss) {
  __shared__ float wsum[kIcThreads / 32];
  __shared__ bool last;
  const long long per_img = (long long)c_limit * N;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per_img;
    const long long r = i - b * per_img;          // c*N + q
    const int q = (int)(r % N);
    const float xv = __ldg(x + b * (long long)C_total * N + r);
    const float d = __fmul_rn(__fmul_rn(xv, __ldg(mask + q)), strength) - __ldg(target + i);
    acc += crit == 0 ? d * d : fabsf(d);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kIcThreads / 32; ++w) t += wsum[w];
    partials[blockIdx.x] = t;
    __threadfence();
    const unsigned int done = atomicAdd(ticket, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned int i = 0; i < gridDim.x; ++i) t += (double)((volatile float*)partials)[i];
    *loss = (float)(t / (double)total);
    *ticket = 0u;
  }
}

__global__ void __launch_bounds__(kIcThreads)
innercos_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ target,
                    const float* __restrict__ grad_loss, int C_total, int c_limit, int N, float strength, int crit,
                    long long total_x, long long total, float* __restrict__ grad_x) {
  const float gl = __ldg(grad_loss);
  const float inv_total = 1.0f / (float)total;
  const long long per_img_x = (long long)C_total * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_x; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per_img_x;
    const long long r = i - b * per_img_x;
    const int c = (int)(r / N);
    const int q = (int)(r % N);
    float gx = 0.f;
    if (c < c_limit) {
      const float m = __ldg(mask + q);
      const float d = __fmul_rn(__fmul_rn(__ldg(x + i), m), strength) - __ldg(target + (b * c_limit + c) * (long long)N + q);
      const float dd = crit == 0 ? 2.f * d : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      gx = gl * dd * inv_total * m * strength;
    }
    grad_x[i] = gx;
  }
}

}  // namespace ipsr

extern "C" int innercos_loss_fwd(const float* x, const float* mask_f32, const float* target,
                                 int B, int C_total, int c_limit, int N, float strength, int crit,
                                 float* partials, uint32_t* ticket, float* loss, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(x && mask_f32 && target && partials && ticket && loss, IPSR_ERR_INVALID_ARG, "innercos_loss_fwd: null pointer");
  IPSR_REQUIRE(B > 0 && C_total > 0 && c_limit > 0 && c_limit <= C_total && N > 0 && (crit == 0 || crit == 1),
               IPSR_ERR_INVALID_ARG, "innercos_loss_fwd: bad arguments B=%d C=%d c_limit=%d N=%d crit=%d", B, C_total, c_limit, N, crit);
  const long long total = (long long)B * c_limit * N;
  long long blocks = (total + (long long)kIcThreads * 8 - 1) / ((long long)kIcThreads * 8);
  if (blocks > kIcMaxBlocks) blocks = kIcMaxBlocks;
  if (blocks < 1) blocks = 1;
  innercos_fwd_kernel<<<(unsigned)blocks, kIcThreads, 0, as_stream(stream)>>>(x, mask_f32, target, C_total, c_limit, N,
                                                                              strength, crit, total, partials, ticket, loss);
  return check_launch("innercos_loss_fwd");
}

extern "C" int innercos_loss_bwd(const float* x, const float* mask_f32, const float* target, const float* grad_loss,
                                 int B, int C_total, int c_limit, int N, float strength, int crit,
                                 float* grad_x, void* stream) {
  using namespace ipsr;
  IPSR_REQUIRE(x && mask_f32 && target && grad_loss && grad_x, IPSR_ERR_INVALID_ARG, "innercos_loss_bwd: null pointer");
  IPSR_REQUIRE(B > 0 && C_total > 0 && c_limit > 0 && c_limit <= C_total && N > 0 && (crit == 0 || crit == 1),
               IPSR_ERR_INVALID_ARG, "innercos_loss_bwd: bad arguments");
  const long long total_x = (long long)B * C_total * N;
  const long long total = (long long)B * c_limit * N;
  long long blocks = (total_x + (long long)kIcThreads * 4 - 1) / ((long long)kIcThreads * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  innercos_bwd_kernel<<<(unsigned)blocks, kIcThreads, 0, as_stream(stream)>>>(x, mask_f32, target, grad_loss, C_total,
                                                                              c_limit, N, strength, crit, total_x, total, grad_x);
  return check_launch("innercos_loss_bwd");
}
