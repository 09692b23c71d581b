// Per-SM throughput of 1-D bulk async copies (cp.async.bulk global -> shared) from an L2-resident buffer.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bulk_bw scripts/micro/bulk_bw.cu && /tmp/bulk_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// one thread per CTA keeps `depth` copies of `bytes` in flight, `iters` copies in total; issuers > 1: that many warps do so independently
__global__ void bulk_kernel(const uint8_t* src, size_t per_cta, uint32_t bytes, int depth, int iters, int issuers, long long* cycles) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) unsigned long long bars[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (issuers < 0) {                                      // lanes 0 .. -issuers-1 of warp 0, converged
    if (warp != 0 || lane >= -issuers) return;
    const int nl = -issuers;
    const uint8_t* base = src + (size_t)blockIdx.x * per_cta;
    const size_t span = per_cta / bytes;
    const uint32_t sm0 = smem_u32(sm) + (uint32_t)lane * depth * bytes;
    const long long t0 = clock64();
    for (int i = 0; i < iters + depth; ++i) {
      const int s = i % depth;
      const uint32_t bar = smem_u32(&bars[lane * 8 + s]);
      if (i >= depth) {
        const uint32_t ph = (uint32_t)((i - depth) / depth) & 1u;
        while (!mbar_try_wait(bar, ph)) {}
      }
      __syncwarp((1u << nl) - 1u);
      if (i < iters) {
        mbar_expect_tx(bar, bytes);
        bulk_g2s(sm0 + (uint32_t)s * bytes, base + ((size_t)(i * nl + lane) % span) * bytes, bytes, bar);
      }
    }
    const long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    return;
  }
  if (lane != 0 || warp >= issuers) return;
  const uint8_t* base = src + (size_t)blockIdx.x * per_cta;
  const size_t span = per_cta / bytes;                     // copies that fit the CTA's region
  const uint32_t sm0 = smem_u32(sm) + (uint32_t)warp * depth * bytes;
  const long long t0 = clock64();
  for (int i = 0; i < iters + depth; ++i) {
    const int s = i % depth;
    const uint32_t bar = smem_u32(&bars[warp * 8 + s]);
    if (i >= depth) {
      const uint32_t ph = (uint32_t)((i - depth) / depth) & 1u;
      while (!mbar_try_wait(bar, ph)) {}
    }
    if (i < iters) {
      mbar_expect_tx(bar, bytes);
      bulk_g2s(sm0 + (uint32_t)s * bytes, base + ((size_t)(i * issuers + warp) % span) * bytes, bytes, bar);
    }
  }
  const long long t1 = clock64();
  if (warp == 0) cycles[blockIdx.x] = t1 - t0;
}
int main() {
  const size_t total = 64ull << 20;                        // 64 MiB: L2-resident
  uint8_t* src; cudaMalloc(&src, total); cudaMemset(src, 1, total);
  long long* cyc; cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int grids[] = {1, 16, 148};
  const uint32_t sizes[] = {2048, 16384};
  for (int issuers : {1, -4, 2}) for (int g : grids) for (uint32_t bytes : sizes) for (int depth : {1, 2, 4}) {
    const int ni = issuers < 0 ? -issuers : issuers;
    if ((size_t)ni * depth * bytes > 192 * 1024 || (g == 16)) continue;
    const int iters = 256;
    const size_t per_cta = total / 148 / 32768 * 32768;
    for (int rep = 0; rep < 2; ++rep) bulk_kernel<<<g, 32 * ni, (size_t)ni * depth * bytes, 0>>>(src, per_cta, bytes, depth, iters, issuers, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, cyc, g * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("issuers %d  CTAs %3d  copy %5u B  depth %d : %6.1f B/cycle/SM  (%lld cycles, %.0f cycles per copy)\n", issuers, g, bytes, depth,
           (double)iters * ni * bytes / mx, mx, (double)mx / iters);
  }
  return 0;
}
