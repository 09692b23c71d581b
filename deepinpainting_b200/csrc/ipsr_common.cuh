// Shared device/host helpers for libipsr_sm100.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ipsr_sm100.h"

namespace ipsr {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define IPSR_REQUIRE(cond, code, ...)                 \
  do {                                                \
    if (!(cond)) {                                    \
      ::ipsr::set_error(__VA_ARGS__);                 \
      return (code);                                  \
    }                                                 \
  } while (0)

#define IPSR_FORWARD(expr)                            \
  do {                                                \
    int rc__ = (expr);                                \
    if (rc__ != IPSR_OK) return rc__;                 \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (IPSR_PDL=1; OFF by default).  The kernels of one forward / backward are short
// (2 .. 35 us at 32x32) and follow each other in one stream, so the idea was: every kernel (a) lets its successor's CTAs
// become resident as soon as all of its own have started (pdl_trigger at the top) and (b) touches global memory only after
// its predecessor has completed and flushed (pdl_wait), overlapping launch, block scheduling and static prologue with the
// predecessor's last wave.  Measured on the B200 (CUDA-graph replay, same box, alternating runs): 143.9 us per step
// against 140.0 us without at 32x32 (batch 16), 1.565 ms against 1.520 ms at 64x64 (batch 64) -- the early residents take
// shared memory and issue slots from the wave that is still working.  Kept as a knob; without the launch attribute both
// instructions are no-ops.  A kernel launched with launch_pdl MUST call pdl_wait() before its first global access.
// ---------------------------------------------------------------------------------------------
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Launchers with PER-IMAGE masks (extension of the reference's one mask per batch): flag / mask_idx / rank are
// [B][ms] (ms = N) and mcount[b] is the number of masked positions of image b, M the batch maximum (the row stride of
// every [B][M] buffer).  ms = 0, mcount = NULL: the shared-mask behaviour of the C entry points of the same name.
int extract_normalize_ex(const float* x, const float* ref, int B, int C, int N, const int32_t* rank_i32, int M,
                         float* inv_norm, float* rnorm, float* xt, float* r_masked, void* x_tiles, void* r_tiles,
                         int32_t* nonfinite, float* rscale, float* rerr, float* xerr, float* xerr_max, void* stream, int ms);
int blend_stage_with_routes_ex(const float* xt, const float* r_masked, const float* inv_norm, const int32_t* ind,
                               const int32_t* mask_idx, const int32_t* flag, int B, int C, int N, int M, float* staged,
                               float* vmask, int32_t* route_ptr, int32_t* route_q, void* stream, int ms, const int32_t* mcount);
int blend_scan_ex(const float* staged, int B, int C, int M, float* y, float* wn, float* wo, void* stream,
                  const int32_t* mcount);
// ipsr_correlate_argmax_fp32 over the whole bank (sparse row lists) + ipsr_resolve_rows as ONE launch; done is [B], zero on entry
int recheck_resolve_ex(const float* x, const float* ref, const float* inv_norm, const float* xt, int B, int C, int N,
                       const int32_t* list, const int32_t* nlist, int64_t* packed, const int32_t* pair_list,
                       const int32_t* npair, const int32_t* cand2, int32_t* ind, int32_t* done, const int32_t* npass2,
                       int32_t* nrecheck_out, int32_t* npass2_out, void* stream);
// ipsr_resolve_rows that also hands the per-image counters out (nrecheck_out / npass2_out, optional)
int resolve_rows_ex(const int64_t* packed, const int32_t* recheck_list, const int32_t* nrecheck, const int32_t* pair_list,
                    const int32_t* npair, const int32_t* cand2, const float* xt, const float* ref, const float* inv_norm,
                    int B, int C, int N, int32_t* ind, float* vmax, const int32_t* npass2, int32_t* nrecheck_out,
                    int32_t* npass2_out, void* stream);
// Optional InnerCos side loss computed while the pasted tiles are still in shared memory (models/networks.py:347:
// `ipsr, innerCos, downnorm_3`; models/InnerCos.py:30-36): loss = mean(crit(out * mask * strength - target)).
struct PasteLoss {
  const float* target;     // [B,C,N]
  const float* mask;       // [N] float, 1 = hole
  float strength;
  int crit;                // 0 = squared error, 1 = absolute error
  float* partials;         // one float per paste CTA (ipsr_paste_loss_partials)
  unsigned int* ticket;    // zeroed u32, left zeroed
  float* loss;             // scalar out
};
int paste_ex(const float* x, const float* y, const int32_t* ind, const int32_t* rank, int B, int C, int N, int M,
             float* out, void* stream, int ms, const PasteLoss* loss = nullptr);
int paste_with_bookkeeping_ex(const float* x, const float* y, const int32_t* ind, const int32_t* rank, const int32_t* flag,
                              const int32_t* mask_idx, const float* wn, const float* wo, int B, int C, int N, int M,
                              float* out, int32_t* route_ptr, int32_t* route_q, int32_t* exc_start, int32_t* exc_cnt,
                              int32_t* exc_l, float* exc_w, int32_t* exc_total, int exc_cap, void* stream, int ms,
                              const int32_t* mcount);
// The decision of ipsr_finalize_argmax folded into the GEMM epilogue (launches without a column split: the epilogue
// thread already holds the row's best / runner-up / third-best): trusted rows get ind, rows with exactly two
// candidates inside the error band go to pair_list, the others to list; packed is reset.
struct TcFinalize {
  const float* rnorm; const float* rscale; const int32_t* nonfinite;
  const float* rerr; const float* xerr_max;   // non-NULL: the rigorous bound of the SINGLE pass (ipsr_finalize_argmax), else the split's
  float tol_rel, tol_abs;
  int32_t* ind; int32_t* list; int32_t* nlist; int64_t* packed;
  int32_t* cand2; int32_t* pair_list; int32_t* npair;
};
int correlate_argmax_tc_ex(const void* r_tiles, const void* x_tiles, int B, int C, int N, int col_begin, int col_end, int psplit,
                           int passes, int r_parts, const int32_t* row_limit, float* part_best, int32_t* part_idx,
                           float* part_second, int32_t* part_idx2, float* part_third, float* s_dump, int n_valid,
                           const TcFinalize* fin, void* stream);
bool tc_pass1_wide(int B, int C, int N, int col_begin, int col_end, int psplit);
int build_routes_ex(const int32_t* ind, const int32_t* flag, const int32_t* mask_idx, int B, int N, int M,
                    int32_t* route_ptr, int32_t* route_q, void* stream, int ms, const int32_t* mcount);
int build_exceptions_ex(const int32_t* ind, const int32_t* mask_idx, const float* wn, const float* wo, int B, int N, int M,
                        int32_t* exc_start, int32_t* exc_cnt, int32_t* exc_l, float* exc_w, int32_t* exc_total,
                        int exc_cap, void* stream, int ms, const int32_t* mcount);

// ---------------------------------------------------------------------------------------------
// tile-image geometry of the fp16 hi/lo operands (written by prep, read by the tcgen05 GEMM)
//   [B][KB = C/64][2 (hi, lo)][RB = N/128][128 rows x 128 bytes, 128B swizzle]
// ---------------------------------------------------------------------------------------------
constexpr int kTileRows = 128;
constexpr int kTileK = 64;                       // fp16 elements per row of a tile = 128 bytes
constexpr int kTileBytes = kTileRows * kTileK * 2;  // 16 KiB

__host__ __device__ inline size_t tile_offset_bytes(int b, int kb, int hl, int rb, int KB, int RB) {
  return ((((size_t)b * KB + kb) * 2 + hl) * RB + rb) * (size_t)kTileBytes;
}

// same with `nhalf` operand parts per 64-channel block (1: hi only, 2: hi, lo)
__host__ __device__ inline size_t tile_offset_bytes_n(int b, int kb, int hl, int rb, int KB, int RB, int nhalf) {
  return ((((size_t)b * KB + kb) * nhalf + hl) * RB + rb) * (size_t)kTileBytes;
}

// byte offset of the 16-byte chunk `chunk` (0..7, 8 fp16 each) of row `r` (0..127) inside a tile
// image: canonical UMMA K-major SWIZZLE_128B layout = Swizzle<3,4,3> on (row*128 + chunk*16).
__host__ __device__ inline uint32_t tile_chunk_offset(int r, int chunk) {
  return (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4));
}

// ---------------------------------------------------------------------------------------------
// order-preserving (score, index) key: signed 64-bit so that ncclMax/int64 (and gloo MAX) work.
//   high 32 bits: fp32 score as an orderable signed int (NaN canonicalised to +NaN > +inf, -0 -> +0)
//   low  32 bits: 0xFFFFFFFF - idx  (lowest index wins ties, util/MaxCoord.py:22 = torch.max)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline long long pack_maxidx(float v, int idx) {
  uint32_t bits;
#ifdef __CUDA_ARCH__
  if (v != v) bits = 0x7FC00000u;
  else bits = __float_as_uint(v + 0.0f);
#else
  if (v != v) bits = 0x7FC00000u;
  else { float t = v + 0.0f; memcpy(&bits, &t, 4); }
#endif
  int32_t key = (bits & 0x80000000u) ? (int32_t)(bits ^ 0x7FFFFFFFu) : (int32_t)bits;
  return (long long)(((unsigned long long)(uint32_t)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx));
}

__host__ __device__ inline void unpack_maxidx(long long packed, float* v, int* idx) {
  uint32_t hi = (uint32_t)((unsigned long long)packed >> 32);
  uint32_t lo = (uint32_t)((unsigned long long)packed & 0xFFFFFFFFull);
  uint32_t bits = (hi & 0x80000000u) ? (hi ^ 0x7FFFFFFFu) : hi;
#ifdef __CUDA_ARCH__
  *v = __uint_as_float(bits);
#else
  memcpy(v, &bits, 4);
#endif
  *idx = (int)(0xFFFFFFFFu - lo);
}

constexpr long long kPackedIdentity = (long long)0x8000000000000000ull;  // INT64_MIN

// ---------------------------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// float -> int64 store of the reference (models/IPSRFunction.py:36,134), returned as the float
// the backward multiplies with: truncation toward zero; NaN / inf / |v| >= 2^63 -> INT64_MIN.
__device__ __forceinline__ float trunc_as_reference(float e) {
  if (!(fabsf(e) < 9.2233720368547758e18f)) return -9.2233720368547758e18f;
  return truncf(e);
}

// ---------------------------------------------------------------------------------------------
// mbarrier / bulk-copy / tcgen05 PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a protocol bug must end in a trap (reported as a CUDA error), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();   // ~2 s
  }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 1-D bulk async copy global -> the SAME shared-memory offset of every CTA of the cluster selected by cta_mask;
// each destination CTA's mbarrier (same offset) receives the complete_tx of the bytes written into it.
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 x fp16 -> fp32 (operand types come from idesc), issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]^T: the A operand from TENSOR memory (lane = row, two fp16 per 32-bit column, K
// contiguous: 8 columns per K = 16 step), so that shared memory feeds B only.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const uint32_t zero = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(zero)
      : "memory");
}
// mbarrier arrive when every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// same, arriving on the mbarrier at this offset in every CTA of the cluster selected by cta_mask
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask)
               : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 consecutive 32-bit columns of this thread's TMEM lane, from registers.
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor of a K-major, 128B-swizzled operand whose 8-row groups are
// 1024 bytes apart (dense tile image): start address, LBO = 16 B (ignored for swizzled K-major),
// SBO = 1024 B, descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__host__ __device__ inline uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// tcgen05 instruction descriptor, kind::f16: D = fp32 (bits 4-5 = 1), A = B = fp16 (format fields bits 7-9 and
// 10-12 = 0; 1 would be bf16), both K-major, M x N tile.
__host__ __device__ inline uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace ipsr
