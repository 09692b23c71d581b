// Issue rate of tcgen05.mma (kind::f16, cta_group::1) on fixed shared-memory tiles: cycles per instruction for
// 128 x N x 16 with both operands in shared memory (SS) and with A in tensor memory (TS), one CTA per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I deepinpainting_b200/csrc -o scripts/micro/umma_rate scripts/micro/umma_rate.cu
#include <cstdio>
#include "ipsr_common.cuh"
using namespace ipsr;

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + 16384u;     // A: one 128 x 64 tile, B: N/128 tiles
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3C003C00u;   // 1.0h
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    const uint64_t da = umma_desc_k_sw128(a_addr), db = umma_desc_k_sw128(b_addr);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS) umma_f16_ts(tmem + (uint32_t)((it & 1) * N) % 256u, tmem + 384u + 8u * k, db + 2 * k, idesc, 1u);
        else umma_f16(tmem + (uint32_t)((it & 1) * N) % 256u, da + 2 * k, db + 2 * k, idesc, 1u);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool TS>
static void run(const char* name, int grid) {
  long long* cyc; cudaMalloc(&cyc, 148 * sizeof(long long));
  auto kern = rate_kernel<N, TS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) kern<<<grid, 128, 60 * 1024>>>(iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[148]; cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = (double)mx / (iters * 4);
  printf("%-28s CTAs %3d: %7.1f cycles per instruction  (%5.1f %% of 8192 flop/cycle/SM)\n", name, grid, per, 100.0 * (128.0 * N * 16 * 2 / per) / 8192.0);
  cudaFree(cyc);
}
int main() {
  for (int g : {1, 148}) {
    run<128, false>("128x128x16 SS", g);
    run<256, false>("128x256x16 SS", g);
    run<128, true>("128x128x16 TS (A in TMEM)", g);
    run<256, true>("128x256x16 TS (A in TMEM)", g);
  }
  return 0;
}
