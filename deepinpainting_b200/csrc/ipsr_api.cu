// Error plumbing, workspace carving and the fused forward of libipsr_sm100.so.
//
// ipsr_shift_forward is the C-ABI counterpart of models/IPSRFunction.py:13-140 for
// shift_sz = stride = 1 (the only configuration the reference can execute): it
// enqueues (a) extract+normalise, (b,c) correlation + arg-max (+ exact recheck), (d) blend scan and
// paste, and the bookkeeping the backward needs -- on one stream, without host synchronisation or
// allocation, so the whole layer is CUDA-graph capturable.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>

#include "ipsr_common.cuh"

namespace ipsr {

static thread_local char g_err[512] = "no error";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("IPSR_PDL");
    return e && atoi(e) == 1;                              // measured slower (see ipsr_common.cuh): off unless asked for
  }();
  return on;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return IPSR_ERR_CUDA;
  }
  return IPSR_OK;
}

constexpr int kMaxPsplit = 8;
// Error allowances of the tcgen05 passes, relative to ||R[q]|| (measured fp32-accumulation error of a C <= 512
// contraction: 4e-7, scripts/tc_accum_error.py; worst case of truncating adds ~ 6e-6).
constexpr float kAccumAllowance = 8e-6f;                 // pass 1: added to the exact operand-rounding bound
constexpr float kDefaultTolRel = 2.0f * (kAccumAllowance + 1.2e-6f);   // pass 2: accumulation + dropped lo*lo + lo rounding
constexpr float kDefaultTolAbs = 1e-12f;

struct Workspace {
  size_t total = 0;
  size_t inv_norm, rnorm, counters /* nonfinite[B] + nrecheck[B] + npass2[B] + xerr_max[B] + npair[B] + done[B] */, xt, r_masked, staged, vmask, y,
      packed, list;
  size_t x_tiles, r_tiles, c_tiles, rscale, rerr, list2, part_best, part_idx, part_second, part_idx2, part_third, cand2, pair_list;
  bool tensor = false;
};

static size_t take(size_t& cur, size_t bytes) {
  const size_t at = cur;
  cur += (bytes + 255) & ~(size_t)255;
  return at;
}

static int resolve_mode(int mode, int C, int N) {
  if (mode == IPSR_MODE_AUTO) return ipsr_tensor_path_supported(C, N) ? IPSR_MODE_TENSOR : IPSR_MODE_EXACT;
  return mode;
}

static Workspace carve(int B, int C, int N, int M, int mode) {
  Workspace w;
  size_t cur = 0;
  const size_t BN = (size_t)B * N, BM = (size_t)B * (M > 0 ? M : 1);
  w.inv_norm = take(cur, BN * 4);
  w.rnorm = take(cur, BN * 4);
  w.counters = take(cur, (size_t)6 * B * 4);
  w.xt = take(cur, BN * C * 4);
  w.r_masked = take(cur, BM * C * 4);
  w.staged = take(cur, (size_t)B * (size_t)((M + ipsr_scan_block_steps(C) - 1) / ipsr_scan_block_steps(C) + 1) *
                            (size_t)ipsr_staged_block_floats(C) * 4);
  w.vmask = take(cur, BM * 4);
  w.y = take(cur, (size_t)B * (size_t)((M + 7) & ~7) * C * 4 + 256);
  w.packed = take(cur, BN * 8);
  w.list = take(cur, BN * 4);
  w.tensor = (resolve_mode(mode, C, N) == IPSR_MODE_TENSOR);
  if (w.tensor) {
    w.x_tiles = take(cur, BN * C * 4);                 // fp16 hi + lo of Xn
    w.r_tiles = take(cur, BN * C * 4);                 // fp16 hi + lo of R
    w.c_tiles = take(cur, BN * C * 4);                 // fp16 hi + lo of the compacted ambiguous rows of R
    w.rscale = take(cur, BN * 4);
    w.rerr = take(cur, BN * 4);
    w.list2 = take(cur, BN * 4);
    w.part_best = take(cur, (size_t)kMaxPsplit * BN * 4);
    w.part_idx = take(cur, (size_t)kMaxPsplit * BN * 4);
    w.part_second = take(cur, (size_t)kMaxPsplit * BN * 4);
    w.part_idx2 = take(cur, (size_t)kMaxPsplit * BN * 4);
    w.part_third = take(cur, (size_t)kMaxPsplit * BN * 4);
    w.cand2 = take(cur, BN * 4);
    w.pair_list = take(cur, BN * 4);
  } else {
    w.x_tiles = w.r_tiles = w.c_tiles = w.rscale = w.rerr = w.list2 = w.part_best = w.part_idx = w.part_second = 0;
    w.part_idx2 = w.part_third = w.cand2 = w.pair_list = 0;
  }
  w.total = cur;
  return w;
}

// A side stream per device for the latency-bound bookkeeping of the backward (exception lists): it forks from
// the caller's stream after the scan, runs next to the bandwidth-bound paste and joins before the forward
// returns -- legal under CUDA-graph stream capture (the fork / join events pull the side stream into the capture).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  std::mutex use;          // held by a caller from its fork record to its join wait: the event pair is shared by every
                           // host thread using this device, and a record / wait pair must not be interleaved with another's
};
static SideStream* side_stream() {
  static SideStream table[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  SideStream& s = table[dev];
  if (!s.stream) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
      s.stream = nullptr;
      return nullptr;
    }
  }
  return &s;
}

template <typename T>
static T* at(const ipsr_fwd_args* a, size_t off) {
  return reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(a->workspace) + off);
}

// col_end < 0, or col_begin == col_end == 0 (a zero-initialised struct): the whole bank
static int shard_end(const ipsr_fwd_args* a, int N) {
  return (a->col_end < 0 || (a->col_end == 0 && a->col_begin == 0)) ? N : a->col_end;
}

static int validate(const ipsr_fwd_args* a, Workspace* w, int* mode_out) {
  IPSR_REQUIRE(a != nullptr, IPSR_ERR_INVALID_ARG, "ipsr_shift_forward: args is null");
  IPSR_REQUIRE(a->x && a->ref && a->flag && a->rank && a->out && a->ind, IPSR_ERR_INVALID_ARG,
               "ipsr_shift_forward: null tensor pointer");
  IPSR_REQUIRE(a->B > 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->M >= 0 && a->M <= a->H * a->W, IPSR_ERR_INVALID_ARG,
               "ipsr_shift_forward: bad dims B=%d C=%d H=%d W=%d M=%d", a->B, a->C, a->H, a->W, a->M);
  IPSR_REQUIRE(a->M == 0 || (a->mask_idx && a->wn && a->wo), IPSR_ERR_INVALID_ARG,
               "ipsr_shift_forward: mask_idx / wn / wo required when M > 0");
  const int N = a->H * a->W;
  IPSR_REQUIRE(a->C % 32 == 0, IPSR_ERR_UNSUPPORTED, "ipsr_shift_forward: C=%d must be a multiple of 32", a->C);
  const int mode = resolve_mode(a->mode, a->C, N);
  IPSR_REQUIRE(mode == IPSR_MODE_TENSOR || mode == IPSR_MODE_EXACT, IPSR_ERR_INVALID_ARG, "ipsr_shift_forward: bad mode %d", a->mode);
  IPSR_REQUIRE(mode != IPSR_MODE_TENSOR || ipsr_tensor_path_supported(a->C, N), IPSR_ERR_UNSUPPORTED,
               "ipsr_shift_forward: tensor mode needs C %% 64 == 0 and N %% 128 == 0 (C=%d N=%d)", a->C, N);
  const int cb = a->col_begin, ce = shard_end(a, N);
  // an EMPTY shard (cb == ce > 0: more ranks than 128-column tiles) is legal in the bank-sharded mode only: the rank
  // contributes identity keys to the exchange and runs everything after it like the others
  IPSR_REQUIRE(cb >= 0 && cb <= ce && ce <= N && (cb < ce || a->stop_after_corr), IPSR_ERR_INVALID_ARG,
               "ipsr_shift_forward: bad column shard [%d,%d)", cb, ce);
  IPSR_REQUIRE(mode != IPSR_MODE_TENSOR || (cb % 128 == 0 && ce % 128 == 0), IPSR_ERR_UNSUPPORTED,
               "ipsr_shift_forward: tensor mode needs 128-aligned column shards, got [%d,%d)", cb, ce);
  if (a->need_grad) {
    IPSR_REQUIRE(a->route_ptr && a->route_q, IPSR_ERR_INVALID_ARG, "ipsr_shift_forward: need_grad without route buffers");
    IPSR_REQUIRE(a->M <= 1 || (a->exc_start && a->exc_cnt && a->exc_l && a->exc_w && a->exc_total && a->exc_cap > 0),
                 IPSR_ERR_INVALID_ARG, "ipsr_shift_forward: need_grad without exception buffers");
  }
  IPSR_REQUIRE(a->mask_stride == 0 || (a->mask_stride == N && a->m_count), IPSR_ERR_INVALID_ARG,
               "ipsr_shift_forward: per-image masks need mask_stride == H*W and m_count (got stride %d)", a->mask_stride);
  *w = carve(a->B, a->C, N, a->M, mode);
  IPSR_REQUIRE(a->workspace && a->workspace_bytes >= w->total, IPSR_ERR_WORKSPACE,
               "ipsr_shift_forward: workspace %zu B < required %zu B", a->workspace_bytes, w->total);
  IPSR_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, IPSR_ERR_WORKSPACE,
               "ipsr_shift_forward: workspace must be 256-byte aligned");
  *mode_out = mode;
  return IPSR_OK;
}

// (d) and the backward bookkeeping; ind[b,q] must be final.
static int run_blend_and_paste(const ipsr_fwd_args* a, const Workspace& w, void* stream) {
  const int B = a->B, C = a->C, N = a->H * a->W, M = a->M;
  const bool grad = a->need_grad != 0;
  // the route builders depend on ind only: they ride in the stage launch
  const int ms = a->mask_stride;                           // 0: one mask for the batch; N: per-image flag / mask_idx / rank rows
  const int32_t* mcount = ms ? a->m_count : nullptr;
  PasteLoss cos_args;
  const PasteLoss* cos = nullptr;
  if (a->cos_target) {
    cos_args.target = a->cos_target; cos_args.mask = a->cos_mask; cos_args.strength = a->cos_strength; cos_args.crit = a->cos_crit;
    cos_args.partials = a->cos_partials; cos_args.ticket = a->cos_ticket; cos_args.loss = a->cos_loss;
    cos = &cos_args;
  }
  IPSR_FORWARD(blend_stage_with_routes_ex(at<float>(a, w.xt), at<float>(a, w.r_masked), at<float>(a, w.inv_norm), a->ind,
                                          a->mask_idx, a->flag, B, C, N, M, at<float>(a, w.staged), at<float>(a, w.vmask),
                                          grad ? a->route_ptr : nullptr, grad ? a->route_q : nullptr, stream, ms, mcount));
  if (M > 0) IPSR_FORWARD(blend_scan_ex(at<float>(a, w.staged), B, C, M, at<float>(a, w.y), a->wn, a->wo, stream, mcount));
  if (grad && M > 1) {
    // the exception lists depend on wn / wo only: they are built on the side stream while the paste streams x -> out
    auto bookkeeping = [&](void* s) -> int {
      return build_exceptions_ex(a->ind, a->mask_idx, a->wn, a->wo, B, N, M, a->exc_start, a->exc_cnt, a->exc_l, a->exc_w,
                                 a->exc_total, a->exc_cap, s, ms, mcount);
    };
    SideStream* ss = side_stream();
    cudaStream_t st = as_stream(stream);
    std::unique_lock<std::mutex> in_use;
    if (ss) in_use = std::unique_lock<std::mutex>(ss->use);
    if (ss && cudaEventRecord(ss->fork, st) == cudaSuccess && cudaStreamWaitEvent(ss->stream, ss->fork, 0) == cudaSuccess) {
      int rc = bookkeeping(ss->stream);
      const int rc2 = paste_ex(a->x, at<float>(a, w.y), a->ind, a->rank, B, C, N, M, a->out, stream, ms, cos);
      // always join, even after an error, so that a capture in progress is not left forked
      const bool joined = cudaEventRecord(ss->join, ss->stream) == cudaSuccess && cudaStreamWaitEvent(st, ss->join, 0) == cudaSuccess;
      if (rc == IPSR_OK) rc = rc2;
      IPSR_REQUIRE(joined, IPSR_ERR_CUDA, "ipsr_shift_forward: side stream join failed");
      return rc;
    }
    (void)cudaGetLastError();
    if (in_use.owns_lock()) in_use.unlock();
    IPSR_FORWARD(bookkeeping(stream));
  }
  return paste_ex(a->x, at<float>(a, w.y), a->ind, a->rank, B, C, N, M, a->out, stream, ms, cos);
}

}  // namespace ipsr

extern "C" const char* ipsr_last_error_string(void) { return ipsr::g_err; }
extern "C" int ipsr_version(void) { return 200; }
extern "C" int ipsr_abi_fwd_args_bytes(void) { return (int)sizeof(ipsr_fwd_args); }

extern "C" int ipsr_tensor_cascade(int B, int C, int N) {
  if (!ipsr_tensor_path_supported(C, N)) return 0;
  // tile MMAs of one single pass; tuning knob IPSR_CASCADE_MIN_WORK (read once) moves the switch-over
  static const long long min_work = [] {
    const char* e = getenv("IPSR_CASCADE_MIN_WORK");
    return e ? atoll(e) : 12000ll;
  }();
  const long long rb = N / ipsr::kTileRows;
  return ((long long)B * rb * rb * (C / ipsr::kTileK) >= min_work) ? 1 : 0;
}

// tensor passes over EVERY row that ipsr_shift_forward issues for a whole-bank call of this size: 1 = one hi*hi pass (the
// cascade, or its "lite" form on small problems without a column split), 3 = the three-pass split over every row
extern "C" int ipsr_tensor_full_passes(int B, int C, int N) {
  if (!ipsr_tensor_path_supported(C, N)) return 0;
  if (ipsr_tensor_cascade(B, C, N)) return 1;
  const char* e = getenv("IPSR_DIRECT_PASSES");
  if (!(e && atoi(e) == 1)) return 3;
  long long tiles = (long long)B * (N / ipsr::kTileRows);
  int ps = (int)(148 / (tiles > 0 ? tiles : 1));
  if (ps > ipsr::kMaxPsplit) ps = ipsr::kMaxPsplit;
  if (ps > N / 128) ps = N / 128;
  return ps <= 1 ? 1 : 3;
}

extern "C" size_t ipsr_workspace_bytes(int B, int C, int H, int W, int M, int mode) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || M < 0) return 0;
  return ipsr::carve(B, C, H * W, M, mode).total;
}

extern "C" int64_t* ipsr_workspace_packed(const ipsr_fwd_args* a) {
  if (!a || !a->workspace) return nullptr;
  const ipsr::Workspace w = ipsr::carve(a->B, a->C, a->H * a->W, a->M, a->mode);
  return ipsr::at<int64_t>(a, w.packed);
}

extern "C" int ipsr_shift_forward(const ipsr_fwd_args* a, void* stream) {
  using namespace ipsr;
  Workspace w;
  int mode = 0;
  IPSR_FORWARD(validate(a, &w, &mode));
  const int B = a->B, C = a->C, N = a->H * a->W, M = a->M;
  const int cb = a->col_begin, ce = shard_end(a, N);
  cudaStream_t st = as_stream(stream);
  int32_t* nonfinite = at<int32_t>(a, w.counters);
  int32_t* nrecheck = nonfinite + B;
  int32_t* npass2 = nrecheck + B;
  float* xerr_max = reinterpret_cast<float*>(npass2 + B);
  int32_t* npair = npass2 + 2 * B;
  int32_t* list = at<int32_t>(a, w.list);
  int64_t* packed = at<int64_t>(a, w.packed);

  cudaError_t e = cudaMemsetAsync(nonfinite, 0, (size_t)6 * B * sizeof(int32_t), st);
  IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "ipsr_shift_forward: memset: %s", cudaGetErrorString(e));
  if (a->need_grad && M > 1) {
    e = cudaMemsetAsync(a->exc_total, 0, ((size_t)2 * B + 2) * sizeof(int32_t), st);
    IPSR_REQUIRE(e == cudaSuccess, IPSR_ERR_CUDA, "ipsr_shift_forward: memset: %s", cudaGetErrorString(e));
  }


  const bool tensor = (mode == IPSR_MODE_TENSOR);
  auto record = [&](void* ev) -> int {
    if (!ev) return IPSR_OK;
    cudaError_t ee = cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev), st);
    IPSR_REQUIRE(ee == cudaSuccess, IPSR_ERR_CUDA, "ipsr_shift_forward: event record: %s", cudaGetErrorString(ee));
    return IPSR_OK;
  };
  IPSR_FORWARD(extract_normalize_ex(a->x, a->ref, B, C, N, a->rank, M, at<float>(a, w.inv_norm), at<float>(a, w.rnorm),
                                    at<float>(a, w.xt), at<float>(a, w.r_masked),
                                    tensor ? at<void>(a, w.x_tiles) : nullptr, tensor ? at<void>(a, w.r_tiles) : nullptr,
                                    nonfinite, tensor ? at<float>(a, w.rscale) : nullptr,
                                    tensor ? at<float>(a, w.rerr) : nullptr, nullptr, tensor ? xerr_max : nullptr, stream,
                                    a->mask_stride));
  if (cb == ce) {
    // empty bank shard: identity keys for the exchange (the operands prepared above serve the steps after it)
    return ipsr_select_all_rows(B, N, list, nrecheck, packed, stream);
  }
  if (tensor) {
    const int RB = N / kTileRows;
    auto auto_split = [&](long long tiles) {
      // one CTA per SM is resident (the ring + row tile fill shared memory): split the bank columns only
      // while that still adds whole CTAs to a single wave
      int ps = a->psplit > 0 ? a->psplit : (int)(148 / (tiles > 0 ? tiles : 1));
      if (ps > kMaxPsplit) ps = kMaxPsplit;
      const int max_split = (ce - cb) / 128;
      if (ps > max_split) ps = max_split;
      return ps < 1 ? 1 : ps;
    };
    const float tol_abs = a->tol_abs >= 0.f ? a->tol_abs : kDefaultTolAbs;
    const float tol_rel = a->tol_rel >= 0.f ? a->tol_rel : kDefaultTolRel;
    int32_t* list2 = at<int32_t>(a, w.list2);
    // The cascade pays ~25 us of extra launches; below ~12 us of single-pass tensor time (tile MMAs per SM) the
    // three-pass split over every row is the shorter path.
    const bool direct = ipsr_tensor_cascade(B, C, N) == 0;
    if (direct) {
      const int ps = auto_split((long long)B * RB);
      IPSR_FORWARD(record(a->ev_corr_begin));
      if (ps == 1 && cb == 0 && ce == N) {
        // No column split: the epilogue decides per row itself (one launch less).  IPSR_DIRECT_PASSES=1 (A/B runs) replaces
        // the three-pass split by ONE hi*hi pass ("cascade lite": rows inside the rigorous single-pass band are settled by
        // two exact dot products or the exact fp32 correlation, no second tensor launch) -- measured SLOWER at configs[1]
        // (166 us per step against 144 us): tracking the third-best score doubles the work of the epilogue, which already
        // limits this kernel, and 3 rows per image reach the fp32 correlation.
        static const int direct_passes = [] {
          const char* e = getenv("IPSR_DIRECT_PASSES");
          return (e && atoi(e) == 1) ? 1 : 3;
        }();
        TcFinalize fin;
        fin.rnorm = at<float>(a, w.rnorm); fin.rscale = at<float>(a, w.rscale); fin.nonfinite = nonfinite;
        fin.rerr = direct_passes == 1 ? at<float>(a, w.rerr) : nullptr;
        fin.xerr_max = direct_passes == 1 ? xerr_max : nullptr;
        fin.tol_rel = direct_passes == 1 ? kAccumAllowance : tol_rel; fin.tol_abs = tol_abs;
        fin.ind = a->ind; fin.list = list; fin.nlist = nrecheck; fin.packed = packed;
        fin.cand2 = at<int32_t>(a, w.cand2); fin.pair_list = at<int32_t>(a, w.pair_list); fin.npair = npair;
        IPSR_FORWARD(correlate_argmax_tc_ex(at<void>(a, w.r_tiles), at<void>(a, w.x_tiles), B, C, N, cb, ce, 1, direct_passes, 2, nullptr,
                                            at<float>(a, w.part_best), at<int32_t>(a, w.part_idx), at<float>(a, w.part_second),
                                            at<int32_t>(a, w.part_idx2), at<float>(a, w.part_third), nullptr, N, &fin, stream));
        IPSR_FORWARD(record(a->ev_corr_end));
      } else {
        IPSR_FORWARD(ipsr_correlate_argmax_tc(at<void>(a, w.r_tiles), at<void>(a, w.x_tiles), B, C, N, cb, ce, ps, 3, 2, nullptr,
                                              at<float>(a, w.part_best), at<int32_t>(a, w.part_idx),
                                              at<float>(a, w.part_second), at<int32_t>(a, w.part_idx2),
                                              at<float>(a, w.part_third), nullptr, stream));
        IPSR_FORWARD(record(a->ev_corr_end));
        IPSR_FORWARD(ipsr_finalize_argmax(at<float>(a, w.part_best), at<int32_t>(a, w.part_idx), at<float>(a, w.part_second),
                                          ps, at<float>(a, w.rnorm), at<float>(a, w.rscale), nullptr, nullptr, nonfinite,
                                          nullptr, nullptr, B, N, tol_rel, tol_abs, a->ind, list, nrecheck, packed, nullptr,
                                          nullptr, C, at<int32_t>(a, w.part_idx2), at<float>(a, w.part_third),
                                          at<int32_t>(a, w.cand2), at<int32_t>(a, w.pair_list), npair, stream));
      }
    } else {
      // pass 1: hi * hi over every row.  Two row tiles per CTA halve the L2 -> SM traffic of the streamed bank tiles:
      // prefer them (with a column split that brings the CTA count back up) whenever that still occupies most SMs
      int ps1 = auto_split((long long)B * RB);
      // (with 128 x 256 x 16 instructions a CTA owns ONE row tile: the split above is already the right one)
      if (RB % 2 == 0 && a->psplit <= 0 && !tc_pass1_wide(B, C, N, cb, ce, ps1)) {
        const int ps_two = auto_split((long long)B * (RB / 2));
        if ((long long)B * (RB / 2) * ps_two >= 100) ps1 = ps_two;
      }
      IPSR_FORWARD(record(a->ev_corr_begin));
      IPSR_FORWARD(ipsr_correlate_argmax_tc(at<void>(a, w.r_tiles), at<void>(a, w.x_tiles), B, C, N, cb, ce, ps1, 1, 2, nullptr,
                                            at<float>(a, w.part_best), at<int32_t>(a, w.part_idx),
                                            at<float>(a, w.part_second), nullptr, nullptr, nullptr, stream));
      IPSR_FORWARD(record(a->ev_corr_end));
      IPSR_FORWARD(ipsr_finalize_argmax(at<float>(a, w.part_best), at<int32_t>(a, w.part_idx), at<float>(a, w.part_second),
                                        ps1, at<float>(a, w.rnorm), at<float>(a, w.rscale), at<float>(a, w.rerr), xerr_max,
                                        nonfinite, nullptr, nullptr, B, N, kAccumAllowance, tol_abs, a->ind, list2, npass2,
                                        packed, at<void>(a, w.r_tiles), at<void>(a, w.c_tiles), C, nullptr, nullptr, nullptr,
                                        nullptr, nullptr, stream));
      // pass 2: the three-pass split over the ambiguous rows finalize just compacted (typically a few % of the rows)
      const int ps2 = auto_split((long long)B * (RB >= 8 ? RB / 8 : 1));
      IPSR_FORWARD(ipsr_correlate_argmax_tc(at<void>(a, w.c_tiles), at<void>(a, w.x_tiles), B, C, N, cb, ce, ps2, 3, 2, npass2,
                                            at<float>(a, w.part_best), at<int32_t>(a, w.part_idx),
                                            at<float>(a, w.part_second), at<int32_t>(a, w.part_idx2),
                                            at<float>(a, w.part_third), nullptr, stream));
      IPSR_FORWARD(ipsr_finalize_argmax(at<float>(a, w.part_best), at<int32_t>(a, w.part_idx), at<float>(a, w.part_second),
                                        ps2, at<float>(a, w.rnorm), at<float>(a, w.rscale), nullptr, nullptr, nonfinite, list2,
                                        npass2, B, N, tol_rel, tol_abs, a->ind, list, nrecheck, nullptr, nullptr, nullptr, C,
                                        at<int32_t>(a, w.part_idx2), at<float>(a, w.part_third), at<int32_t>(a, w.cand2),
                                        at<int32_t>(a, w.pair_list), npair, stream));
    }
  } else {
    IPSR_FORWARD(ipsr_select_all_rows(B, N, list, nrecheck, packed, stream));
  }
  if (tensor && cb == 0 && ce == N && !a->stop_after_corr) {
    // whole bank: the exact recheck of the listed rows and the pair / key resolution in ONE launch
    IPSR_FORWARD(recheck_resolve_ex(a->x, a->ref, at<float>(a, w.inv_norm), at<float>(a, w.xt), B, C, N, list, nrecheck, packed,
                                    at<int32_t>(a, w.pair_list), npair, at<int32_t>(a, w.cand2), a->ind, npair + B, npass2,
                                    a->nrecheck_out, a->npass2_out, stream));
    return run_blend_and_paste(a, w, stream);
  }
  if (!tensor) IPSR_FORWARD(record(a->ev_corr_begin));
  IPSR_FORWARD(ipsr_correlate_argmax_fp32(a->x, a->ref, at<float>(a, w.inv_norm), B, C, N, cb, ce, list, nrecheck,
                                          tensor ? 0 : (N + 63) / 64, packed, stream));
  if (!tensor) IPSR_FORWARD(record(a->ev_corr_end));
  // rows recomputed in exact fp32 take their key; rows with exactly two candidates are settled by two exact dot products
  // (the kernel also copies the per-image counters out: memcpy nodes here would cut the programmatic launch chain)
  IPSR_FORWARD(resolve_rows_ex(packed, list, nrecheck, tensor ? at<int32_t>(a, w.pair_list) : nullptr, npair,
                               tensor ? at<int32_t>(a, w.cand2) : nullptr, at<float>(a, w.xt), a->ref, at<float>(a, w.inv_norm), B, C, N,
                               a->ind, nullptr, npass2, a->nrecheck_out, a->npass2_out, stream));
  if (a->stop_after_corr) {
    // bank-sharded mode: every row leaves with an exact (score, idx) key of its LOCAL winner
    if (tensor)
      IPSR_FORWARD(ipsr_pack_winner_scores(at<float>(a, w.xt), a->ref, at<float>(a, w.inv_norm), a->ind, B, C, N, packed, stream));
    return IPSR_OK;
  }
  return run_blend_and_paste(a, w, stream);
}

extern "C" int ipsr_shift_forward_finish(const ipsr_fwd_args* a, void* stream) {
  using namespace ipsr;
  Workspace w;
  int mode = 0;
  IPSR_FORWARD(validate(a, &w, &mode));
  const int B = a->B, N = a->H * a->W;
  IPSR_FORWARD(ipsr_unpack_maxidx(at<int64_t>(a, w.packed), (int64_t)B * N, nullptr, a->ind, stream));
  return run_blend_and_paste(a, w, stream);
}
