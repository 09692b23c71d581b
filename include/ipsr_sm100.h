/*
 * ipsr_sm100.h -- C ABI of libipsr_sm100.so: the B200 (sm_100a) implementation of the
 * IPSR / CSA patch-shift attention layer of DeepInPainting.
 *
 * The reference has no native code and no FFI: every function below replaces a PyTorch
 * call sequence of the reference (cited as file:line relative to the reference tree), and is
 * bound from Python with ctypes by deepinpainting_b200/_lib.py (see INTEGRATION.md for the
 * binding a maintainer of the reference would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - tensors are contiguous; feature maps are NCHW fp32 exactly as the reference holds them;
 *   - N = H*W positions in raster order (q = i*W + j), M = number of masked positions;
 *   - every launcher enqueues on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     never allocates: scratch memory comes from the caller (ipsr_workspace_bytes);
 *   - return value 0 = success, negative = error (ipsr_last_error_string() describes it).
 */
#ifndef IPSR_SM100_H_
#define IPSR_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPSR_OK                 0
#define IPSR_ERR_INVALID_ARG   -1
#define IPSR_ERR_UNSUPPORTED   -2
#define IPSR_ERR_CUDA          -3
#define IPSR_ERR_WORKSPACE     -4

/* correlation precision modes (north_star (b)) */
#define IPSR_MODE_AUTO    0   /* tensor path when the shape allows it, else exact */
#define IPSR_MODE_TENSOR  1   /* tcgen05 fp16 precision cascade (1 pass, 3-pass split on ambiguous rows) + fp32 recheck */
#define IPSR_MODE_EXACT   2   /* fp32 FFMA correlation for every row */

const char* ipsr_last_error_string(void);
int ipsr_version(void);
/* sizeof(ipsr_fwd_args) as this library was compiled: a binding checks its own mirror of the struct against it. */
int ipsr_abi_fwd_args_bytes(void);
/* 1 when (C, N) can run on the tcgen05 path (C % 64 == 0, N % 128 == 0, N <= 65536). */
int ipsr_tensor_path_supported(int C, int N);
/* 1 when the tensor path of ipsr_shift_forward runs the precision cascade (single pass + three-pass split on the
 * ambiguous rows) for this problem size, 0 when it runs the three-pass split over every row (small problems, where
 * the cascade's extra launches cost more than the tensor time they save). */
int ipsr_tensor_cascade(int B, int C, int N);
/* Tensor passes that ipsr_shift_forward issues over EVERY row for a whole-bank call of this size: 1 (one hi*hi pass: the
 * cascade) or 3 (the three-pass split over every row: small problems); 0: no tensor path. */
int ipsr_tensor_full_passes(int B, int C, int N);

/* ---------------------------------------------------------------------------------------------
 * Mask helpers
 * ------------------------------------------------------------------------------------------- */

/* util/util.py:68-84 cal_feat_mask: `conv_layers` chained 4x4/stride-2/pad-1 box filters
 * (weights 1/16) then `> threshold`.  mask_u8 [S_h,S_w] (0/1) -> feat_u8 [S_h>>L, S_w>>L].
 * scratch_i32 must hold 2 * (S_h/2)*(S_w/2) ints.  Exact integer arithmetic
 * (S_h, S_w multiples of 2^L, L <= 5). */
int ipsr_feat_mask(const uint8_t* mask_u8, int S_h, int S_w, int conv_layers, float threshold,
                   uint8_t* feat_u8, int32_t* scratch_i32, void* stream);

/* The same for a batch of B masks [B][S_h][S_w] -> [B][S_h>>L][S_w>>L] (one free-form mask per sample, BASELINE.json
 * configs[4]); scratch_i32 must hold 2 * B * (S_h/2)*(S_w/2) ints. */
int ipsr_feat_mask_batch(const uint8_t* mask_u8, int B, int S_h, int S_w, int conv_layers, float threshold,
                         uint8_t* feat_u8, int32_t* scratch_i32, void* stream);

/* util/util.py:88-147 cal_mask_given_mask_thred: flag[P] = (sum of mask over the k x k window
 * >= mask_thred); mask_idx = ascending positions with flag==1; rank[q] = position of q inside
 * mask_idx or -1; *count = M.  P = nH*nW patch positions.  One CTA; P <= 65536. */
int ipsr_build_flags(const uint8_t* feat_u8, int H, int W, int patch, int stride, int mask_thred,
                     int32_t* flag_i32, int32_t* mask_idx_i32, int32_t* rank_i32, int32_t* count_i32,
                     void* stream);

/* The same for B feature masks [B][H][W]: flag / mask_idx / rank rows of P entries per image (mask_idx row b holds
 * count[b] valid entries) -- exactly the per-image layout of ipsr_fwd_args.mask_stride = P. */
int ipsr_build_flags_batch(const uint8_t* feat_u8, int B, int H, int W, int patch, int stride, int mask_thred,
                           int32_t* flag_i32, int32_t* mask_idx_i32, int32_t* rank_i32, int32_t* count_i32,
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * (a) patch extraction + L2 normalisation   (util/NonparametricShift.py:36-40,59-73)
 *
 * For X = x[b] viewed [N,C]:  inv_norm[b,p] = 1/(||X[p]||_2 + 1e-8);
 *   xt        fp32 [B,N,C]   position-major copy of x (raw patches = decoder weights, NPS:54);
 *   x_tiles   fp16 hi/lo split of Xn * 2^11, Xn = X*inv_norm, in the UMMA tile image layout (may be NULL);
 * For R = ref[b] viewed [N,C]:  rnorm[b,q] = ||R[q]||_2;
 *   r_masked  fp32 [B,M,C]   rows of R at masked positions (rank_i32 from ipsr_build_flags);
 *   r_tiles   fp16 hi/lo split of R[q] * 2^s_q (max |.| in [2^13, 2^14)), same layout (may be NULL);
 *   rscale    [B,N]  2^-(s_q + 11): true score = tensor score * rscale[b,q];
 *   rerr      [B,N]  ||R[q] - hi part||_2 (true units);  xerr [B,N] (may be NULL) the same for Xn[p];
 *   xerr_max  [B]    max_p xerr[b,p] (must be zero on entry).
 *   => |single-pass score - exact score| <= rerr[q] + ||R~[q]|| * xerr_max[b] (+ fp32 accumulation).
 * nonfinite [B] (may be NULL): set to 1 when x[b] or ref[b] holds a NaN/inf (the tensor path
 * then defers every row of that image to the exact path).  Must be zero on entry.
 * Tile image layout: [B][C/64][2 (hi, lo)][N/128][128 rows x 64 fp16, 128B-swizzled K-major].
 * ------------------------------------------------------------------------------------------- */
int ipsr_extract_normalize(const float* x, const float* ref, int B, int C, int N,
                           const int32_t* rank_i32, int M,
                           float* inv_norm, float* rnorm, float* xt, float* r_masked,
                           void* x_tiles, void* r_tiles, int32_t* nonfinite,
                           float* rscale, float* rerr, float* xerr, float* xerr_max, void* stream);

/* Rows list[b][0 .. nlist[b]) of ref[b] -> compact fp16 hi AND lo tile images (2 parts) for the three-pass
 * correlation of the ambiguous rows: compact row r = position list[b][r], scaled by the same 2^s_q
 * (rscale from ipsr_extract_normalize); rows up to the next multiple of 128 are zero-filled. */
int ipsr_compact_rows(const float* ref, int B, int C, int N, const int32_t* list, const int32_t* nlist,
                      const float* rscale, void* c_tiles, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (b)+(c) correlation + fused row arg-max   (models/IPSRFunction.py:59, util/MaxCoord.py:22)
 *
 * S[q,p] = <R[q], Xn[p]> is never written to memory.  Precision cascade: one tcgen05 pass (hi*hi, fp16)
 * over every row; rows whose top-2 gap is inside twice its rigorous error bound are compacted and redone
 * with the three-pass split (hi*lo + lo*hi + hi*hi, ~fp32 accurate); rows still ambiguous after that are
 * recomputed in exact fp32.
 * ------------------------------------------------------------------------------------------- */

/* tcgen05/TMEM GEMM over bank columns [col_begin, col_end) (multiples of 128), split into `psplit` column
 * ranges per row tile.  passes = 1: hi * hi only (128 x 256 x 16 instructions, one row tile per CTA, when the column
 * range is a multiple of 256, the splits stay whole 256-column blocks and C <= 512; two row tiles per CTA with
 * 128 x 128 x 16 instructions otherwise); passes = 3: hi*lo + lo*hi + hi*hi.  r_parts: parts per
 * 64-channel block in the r_tiles image (2; 1 is accepted for a hi-only image with passes = 1).
 * row_limit (optional, [B]): only the first row_limit[b] rows of r_tiles are populated (compacted operand).
 * Writes per row and per split the best SCALED score, its (global) column index and the runner-up score:
 * part_* are [psplit][B][N].  part_idx2 / part_third (optional, passes = 3 only): column of the runner-up and the
 * third-best score.  s_dump (optional, tests only): fp32 [B][N][N] receives the full scaled score
 * matrix. */
int ipsr_correlate_argmax_tc(const void* r_tiles, const void* x_tiles, int B, int C, int N,
                             int col_begin, int col_end, int psplit, int passes, int r_parts,
                             const int32_t* row_limit, float* part_best, int32_t* part_idx, float* part_second,
                             int32_t* part_idx2, float* part_third, float* s_dump, void* stream);

/* ipsr_correlate_argmax_tc with padding: bank columns >= n_valid (<= N) never win (patch maps whose position count is
 * not a multiple of 128 are padded with zero rows up to N). */
int ipsr_correlate_argmax_tc_valid(const void* r_tiles, const void* x_tiles, int B, int C, int N,
                                   int col_begin, int col_end, int psplit, int passes, int r_parts,
                                   const int32_t* row_limit, float* part_best, int32_t* part_idx, float* part_second,
                                   int32_t* part_idx2, float* part_third, float* s_dump, int n_valid, void* stream);

/* Merge the `psplit` partial (best, idx, second) triples per row, then decide per row:
 *   gap = (best - second) * rscale[b,q] >= tol[q]  -> ind[b,q] = idx (trusted);
 *   otherwise (or when nonfinite[b] != 0) the row is appended to list_out[b] (ind[b,q] = idx provisionally).
 * Rows: list_in == NULL: row r is position q = r; otherwise row r < nlist_in[b] is position list_in[b][r].
 * Tolerance: rerr != NULL (after the single pass): tol[q] = 2 (rerr[q] + rnorm[q] (1.001 xerr_max[b] + tol_rel)) + tol_abs;
 *   rerr == NULL (after the three-pass split): tol[q] = tol_rel rnorm[q] + tol_abs.
 * packed (optional): packed[b,q] is reset for every row seen.  c_tiles (optional, list_in == NULL only): COMPACTS --
 *   the hi and lo tile rows of every appended position are copied from r_tiles to row `pos` of the compact image
 *   c_tiles (same layout) that the three-pass split then reads with row_limit = nlist_out.
 * part_idx2 / part_third (optional): an untrusted row whose THIRD-best score is outside the band has exactly two
 *   candidates; it goes to pair_list[b] (cand2[b,q] = the runner-up's column) instead of list_out[b].
 * nlist_out[b] (and npair[b]) must be zero on entry. */
int ipsr_finalize_argmax(const float* part_best, const int32_t* part_idx, const float* part_second,
                         int psplit, const float* rnorm, const float* rscale, const float* rerr,
                         const float* xerr_max, const int32_t* nonfinite,
                         const int32_t* list_in, const int32_t* nlist_in,
                         int B, int N, float tol_rel, float tol_abs,
                         int32_t* ind, int32_t* list_out, int32_t* nlist_out, int64_t* packed,
                         const void* r_tiles, void* c_tiles, int C,
                         const int32_t* part_idx2, const float* part_third,
                         int32_t* cand2, int32_t* pair_list, int32_t* npair, void* stream);

/* ipsr_finalize_argmax over the first n_valid (<= N) rows only (the others are padding of the tile image). */
int ipsr_finalize_argmax_valid(const float* part_best, const int32_t* part_idx, const float* part_second,
                               int psplit, const float* rnorm, const float* rscale, const float* rerr,
                               const float* xerr_max, const int32_t* nonfinite,
                               const int32_t* list_in, const int32_t* nlist_in,
                               int B, int N, float tol_rel, float tol_abs,
                               int32_t* ind, int32_t* list_out, int32_t* nlist_out, int64_t* packed,
                               const void* r_tiles, void* c_tiles, int C,
                               const int32_t* part_idx2, const float* part_third,
                               int32_t* cand2, int32_t* pair_list, int32_t* npair, int n_valid, void* stream);

/* Select every row for the exact path (IPSR_MODE_EXACT): recheck_list[b] = 0..N-1,
 * nrecheck[b] = N, packed = identity. */
int ipsr_select_all_rows(int B, int N, int32_t* recheck_list, int32_t* nrecheck, int64_t* packed,
                         void* stream);

/* Exact fp32 correlation (sequential-in-channel FFMA on Xn = fl(X*inv_norm), the reference's
 * fp32 arithmetic) for the rows in recheck_list over bank columns [col_begin, col_end):
 * packed[b,q] = max(packed[b,q], pack(score, col)) with the order-preserving packing of
 * ipsr_pack_maxidx (largest score wins, lowest index on ties, NaN wins -- torch.max).
 * row_ctas >= 1: dense kernel (64 x 64 register tiles), that many CTAs cooperate over the row list of
 * one image; row_ctas <= 0: sparse kernel for short lists (8 rows at a time against 64 columns). */
int ipsr_correlate_argmax_fp32(const float* x, const float* ref, const float* inv_norm,
                               int B, int C, int N, int col_begin, int col_end,
                               const int32_t* recheck_list, const int32_t* nrecheck, int row_ctas,
                               int64_t* packed, void* stream);

/* Settle what the tensor passes left open, in one launch: (1) packed[b,q] -> ind[b,q] (and optionally vmax[b,q])
 * for the rows in recheck_list; (2) for the rows in pair_list (may be NULL) -- exactly two candidates, ind[b,q]
 * and cand2[b,q], inside the error band of the three-pass split -- one warp computes both exact fp32 scores
 * and keeps the larger (the lower column on a tie). */
int ipsr_resolve_rows(const int64_t* packed, const int32_t* recheck_list, const int32_t* nrecheck,
                      const int32_t* pair_list, const int32_t* npair, const int32_t* cand2,
                      const float* xt, const float* ref, const float* inv_norm,
                      int B, int C, int N, int32_t* ind, float* vmax, void* stream);
/* packed[b,q] -> ind[b,q] (and optionally vmax[b,q]) for the rows in recheck_list. */
int ipsr_apply_recheck(const int64_t* packed, const int32_t* recheck_list, const int32_t* nrecheck,
                       int B, int N, int32_t* ind, float* vmax, void* stream);

/* Bank-sharded mode: exact fp32 score of the (trusted) local winner of EVERY row, packed for the
 * exchange step: packed[b,q] = pack(<R[q], Xn[ind[b,q]]>, ind[b,q]).  Rows already holding an
 * exact key from ipsr_correlate_argmax_fp32 (packed != identity) are left alone. */
int ipsr_pack_winner_scores(const float* xt, const float* ref, const float* inv_norm, const int32_t* ind,
                            int B, int C, int N, int64_t* packed, void* stream);

/* (max, idx) <-> int64 keys for ncclAllReduce(ncclInt64, ncclMax) / gloo MAX. */
int ipsr_pack_maxidx(const float* v, const int32_t* idx, int64_t n, int64_t* packed, void* stream);
int ipsr_unpack_maxidx(const int64_t* packed, int64_t n, float* v, int32_t* idx, void* stream);

/* util/MaxCoord.py:16-28 on a MATERIALISED score tensor s [P,L] (P bank patches, L locations):
 * ind[l] = argmax_p s[p,l] (first index on ties, NaN wins), vmax[l] = the maximum. */
int ipsr_maxcoord(const float* s, int P, int L, int64_t* ind_i64, float* vmax, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (d) coherent blend over masked positions + gather-paste   (models/IPSRFunction.py:70-133)
 * ------------------------------------------------------------------------------------------- */

/* Stage the operands of the sequential blend.  Steps l = 0..M-1 (q_l = mask_idx[l], p_l = ind[b,q_l]) are
 * grouped in blocks of T = ipsr_scan_block_steps(C) consecutive steps; one staged block is contiguous:
 *   U  [T][C]  u_l = X[q_l] * inv_norm[q_l]                      (IPSRFunction.py:109)
 *   K  [T][C]  X[p_l]                                            (known_region, :95)
 *   Gt [T][T]  Gt[j][i] = <u_{l0+i}, X[p_{l0+j}]>                (in-block Gram matrix)
 *   v  [T]     v_l = <R[q_l], Xn[p_l]>  exact fp32               (vmax at masked positions, :70)
 * staged is [B][ceil(M/T)][ipsr_staged_block_floats(C)]; vmask [B,M] (optional copy of v_l, may be NULL).
 * C % 32 == 0, C <= 1024. */
int ipsr_blend_stage(const float* xt, const float* r_masked, const float* inv_norm,
                     const int32_t* ind, const int32_t* mask_idx, int B, int C, int N, int M,
                     float* staged, float* vmask, void* stream);
/* ipsr_blend_stage + ipsr_build_routes as ONE launch (the route builders depend on ind only and are
 * latency-bound; they overlap the gather / Gram work).  route_ptr == NULL: plain ipsr_blend_stage.
 * M == 0: only the routes are built. */
int ipsr_blend_stage_with_routes(const float* xt, const float* r_masked, const float* inv_norm,
                                 const int32_t* ind, const int32_t* mask_idx, const int32_t* flag,
                                 int B, int C, int N, int M, float* staged, float* vmask,
                                 int32_t* route_ptr, int32_t* route_q, void* stream);
/* steps per staged block (32 for C <= 256, 16 for C <= 512, else 8) and floats per staged block */
int ipsr_scan_block_steps(int C);
int ipsr_staged_block_floats(int C);
/* rows of y per image: M rounded up to a multiple of 8 */
int ipsr_padded_steps(int M);

/* The recurrence itself, one CTA per image.  l=0: y_0 = X[p_0] (:98-101);  l>0: a = <u_l, y_{l-1}>;
 * wn = a/(a+v); wo = v/(a+v); y_l = wn*y_{l-1} + wo*X[p_l] (:104-122, no clamping).  Inside a block the
 * scalars a_l are tracked by linearity (a_i <- wn_l a_i + wo_l Gt[l][i]) so that the dependent chain
 * is one scalar step per masked position; y is re-anchored at every block boundary.
 * Writes y [B][C][ipsr_padded_steps(M)] (channel-major: the paste reads rows), wn/wo [B,M] (wn[b,0] = 0,
 * wo[b,0] = 1). */
int ipsr_blend_scan(const float* staged, int B, int C, int M,
                    float* y, float* wn, float* wo, void* stream);

/* floats of the partial-sum buffer of the fused InnerCos loss (ipsr_fwd_args.cos_partials) */
int ipsr_paste_loss_partials(int B, int C, int N);

/* out[b,:,q] = y[b,:,rank[q]] (y laid out [B][C][ipsr_padded_steps(M)]) for masked q,
 * x[b,:,ind[b,q]] otherwise (replaces the dense
 * conv_transpose of IPSRFunction.py:131). */
int ipsr_paste(const float* x, const float* y, const int32_t* ind, const int32_t* rank,
               int B, int C, int N, int M, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (e) backward   (models/IPSRFunction.py:144-178)
 *
 * The reference stores the attention A in a LongTensor (:36,134), so the backward sees
 * trunc(A): weight 1 for every unmasked row q and for the first masked row q_0 (the "unit
 * routes" q -> ind[q]), and for masked rows l >= 1 only the entries with |A| >= 1
 * ("exceptions", absent for well-conditioned inputs).
 * ------------------------------------------------------------------------------------------- */

/* Unit routes grouped by bank column p = ind[b,q], ascending q inside a group (deterministic):
 * route_ptr [B][N+1] (CSR), route_q [B][N].  N <= 16384. */
int ipsr_build_routes(const int32_t* ind, const int32_t* flag, const int32_t* mask_idx,
                      int B, int N, int M, int32_t* route_ptr, int32_t* route_q, void* stream);

/* Replay the attention rows (:123-125: row_l = row_{l-1}*wn_l; row_l[p_l] += wo_l) per bank
 * column and emit every entry of rows l >= 1 that survives the float -> int64 store.  The entries of the WHOLE BATCH
 * share one pool exc_l / exc_w [exc_cap] (exc_l holds the POSITION q_l = mask_idx[l] of the row the entry belongs to,
 * exc_w its truncated weight): image b reserves the contiguous range [base_b, base_b + count_b) with one atomic on
 * the pool cursor, so a chaotic image borrows the room the others do not need.  exc_start / exc_cnt [B][N]: range of
 * column p inside its image's range.  exc_total is the exception STATE, int32 [2B + 2], zero on entry:
 *   [b] = count_b (>= 0x3FFFFFFF: the lists of image b are unusable -- pool exhausted or non-finite weights -- and
 *   the backward replays the recurrence for that image), [B + b] = base_b, [2B] = pool cursor. */
int ipsr_build_exceptions(const int32_t* ind, const int32_t* mask_idx, const float* wn, const float* wo,
                          int B, int N, int M, int32_t* exc_start, int32_t* exc_cnt,
                          int32_t* exc_l, float* exc_w, int32_t* exc_total, int exc_cap, void* stream);

/* gin[b,:,p] = g[b,:,p] + triple_w * ( sum_{q in routes(p)} g[b,:,q]
 *                                      + sum_{e in exc(p)} exc_w[e] * g[b,:,exc_l[e]] ) */
int ipsr_shift_bwd(const float* g, int B, int C, int N, int M,
                   const int32_t* route_ptr, const int32_t* route_q,
                   const int32_t* exc_start, const int32_t* exc_cnt, const int32_t* exc_l, const float* exc_w,
                   const int32_t* exc_total, int exc_cap,
                   const int32_t* ind, const int32_t* mask_idx, const float* wn, const float* wo,
                   float triple_w, float* gin, void* stream);

/* ipsr_shift_bwd with per-image masks (see ipsr_fwd_args.mask_stride): mask_idx is [B][mask_stride], m_count [B];
 * mask_stride = 0, m_count = NULL is ipsr_shift_bwd. */
int ipsr_shift_bwd_masks(const float* g, int B, int C, int N, int M,
                         const int32_t* route_ptr, const int32_t* route_q,
                         const int32_t* exc_start, const int32_t* exc_cnt, const int32_t* exc_l, const float* exc_w,
                         const int32_t* exc_total, int exc_cap,
                         const int32_t* ind, const int32_t* mask_idx, const float* wn, const float* wo,
                         float triple_w, float* gin, int mask_stride, const int32_t* m_count, void* stream);

/* ipsr_paste (+ ipsr_build_routes) + ipsr_build_exceptions as ONE launch (independent once the scan has
 * finished; the latency-bound builders overlap the bandwidth-bound paste).  route_ptr == NULL: the routes
 * were already built (ipsr_blend_stage_with_routes).  exc_total ([2B + 2] state) must be zero on entry. */
int ipsr_paste_with_bookkeeping(const float* x, const float* y, const int32_t* ind, const int32_t* rank,
                                const int32_t* flag, const int32_t* mask_idx, const float* wn, const float* wo,
                                int B, int C, int N, int M, float* out,
                                int32_t* route_ptr, int32_t* route_q,
                                int32_t* exc_start, int32_t* exc_cnt, int32_t* exc_l, float* exc_w,
                                int32_t* exc_total, int exc_cap, void* stream);

/* ---------------------------------------------------------------------------------------------
 * shift_sz = k > 1 / stride = s > 1 ("patch mode"), FORWARD ONLY   (models/IPSRFunction.py:46-133,
 * util/NonparametricShift.py:59-73).  The reference computes the whole output for these settings and then fails
 * storing the attention for its backward (:134); its backward is undefined.
 *
 * A patch is a row of K = C*k*k values in unfold order kk = (c*k + dy)*k + dx; P = nH*nW patch positions,
 * nH = (H-k)/s + 1.  K <= 1024: unfold x and ref into patch maps, run ipsr_shift_forward on them as
 * [B, Kpad, nH, nW] feature maps, fold the result.  K > 1024: ipsr_patch_rows, the exact correlation
 * (ipsr_correlate_argmax_fp32 on the patch maps), ipsr_blend_wide, ipsr_fold_patch_rows.
 * ------------------------------------------------------------------------------------------- */
int ipsr_patch_row_len(int C, int patch);
/* cols [B][Kpad][P]: cols[b][kk][i*nW+j] = x[b][c][i*s+dy][j*s+dx]; rows K <= kk < Kpad are zero-filled
 * (Kpad <= 0: Kpad = K). */
int ipsr_unfold_patches(const float* x, int B, int C, int H, int W, int patch, int stride, int Kpad,
                        float* cols, void* stream);
/* ConvTranspose2d(P, C, k, s) with the pasted patches as input (IPSRFunction.py:131): out[b][c][Y][X] = sum of
 * cols[b][(c,dy,dx)][(i,j)] over the patches with i*s+dy = Y, j*s+dx = X.  Needs (nH-1)*s + k == H (the reference
 * fails at :133 otherwise). */
int ipsr_fold_patches(const float* cols, int B, int C, int H, int W, int patch, int stride, int Kpad,
                      float* out, void* stream);
/* rows [B][P][K] position-major raw patches and (optional) inv_norm [B][P] = 1/(||patch||_2 + 1e-8). */
int ipsr_patch_rows(const float* x, int B, int C, int H, int W, int patch, int stride,
                    float* rows, float* inv_norm, void* stream);
/* the same, plus (optional) norm [B][P] = ||patch||_2 and maxabs [B][P] = max |patch value|. */
int ipsr_patch_rows_stats(const float* x, int B, int C, int H, int W, int patch, int stride,
                          float* rows, float* inv_norm, float* norm, float* maxabs, void* stream);

/* Long patch rows (K > 1024) on the tensor cores (BASELINE configs[3]; models/IPSRFunction.py:54-59 with shift_sz = 3).
 * Kpad = K rounded up to 64, Ppad = P rounded up to 128; every [B][Ppad] array below has row stride Ppad.
 * ipsr_patch_tiles: rows [B][P][K] -> fp16 hi/lo operand images [B][Kpad/64][2][Ppad/128][128 x 64] (zero padding) for
 *   ipsr_correlate_argmax_tc_valid (passes = 3, n_valid = P).  is_ref = 0: values fl(fl(x * inv_norm) * 2^11);
 *   is_ref = 1: values r * 2^s_q (max |.| in [2^13, 2^14)), rscale [B][Ppad] = 2^-(s_q + 11) and rnorm_pad [B][Ppad] =
 *   norm (padding rows: 1 and 0) for ipsr_finalize_argmax_valid. */
int ipsr_patch_tiles(const float* rows, const float* inv_norm, const float* maxabs, const float* norm, int is_ref,
                     int B, int K, int P, void* tiles, float* rscale, float* rnorm_pad, void* stream);
/* Rows list[b][0 .. nlist[b]) (row stride Ppad) recomputed in fp32 against bank columns [col_begin, col_end):
 * packed[b][q] (stride Ppad) = max(packed, pack(score, column)) as ipsr_correlate_argmax_fp32 does. */
int ipsr_patch_recheck(const float* rows_x, const float* rows_r, const float* inv_norm, int B, int K, int P,
                       int col_begin, int col_end, const int32_t* list, const int32_t* nlist, int64_t* packed, void* stream);
/* Rows of pair_list (exactly two candidates, ind_pad[b][q] and cand2[b][q], inside the error band: the pair output of
 * ipsr_finalize_argmax_valid; strides Ppad): two exact fp32 dot products per row; the winner goes to ind_pad and, as an
 * exact (score, column) key, to packed. */
int ipsr_patch_resolve_pairs(const float* rows_x, const float* rows_r, const float* inv_norm, int B, int K, int P,
                             const int32_t* pair_list, const int32_t* npair, const int32_t* cand2, int32_t* ind_pad,
                             int64_t* packed, void* stream);
/* Per row q < P: exact fp32 score of its winner ind_pad[b][q] (or the key the recheck left in packed_pad, both with
 * stride Ppad; packed_pad may be NULL) -> ind [B][P], vmax [B][P] (IPSRFunction.py:70) and keys [B][P] for the
 * bank-sharded exchange (each optional). */
int ipsr_patch_winner_scores(const float* rows_x, const float* rows_r, const float* inv_norm, const int32_t* ind_pad,
                             const int64_t* packed_pad, int B, int K, int P, int32_t* ind, float* vmax, int64_t* keys,
                             void* stream);
/* The blend recurrence (IPSRFunction.py:82-126) on rows of K <= 8192 values, one CTA per image:
 * y [B][M][K], wn / wo [B][M] (wn[b,0] = 0, wo[b,0] = 1).  vmax [B][P] = row maxima of the correlation. */
int ipsr_blend_wide(const float* rows, const float* inv_norm, const float* vmax, const int32_t* ind,
                    const int32_t* mask_idx, int B, int K, int P, int M,
                    float* y, float* wn, float* wo, void* stream);
/* The same recurrence in blocks of 32 steps: by linearity only one scalar per step is sequential (the scheme of
 * ipsr_blend_scan for rows too long for its shared-memory tiles); the in-block Gram matrices <u_i, X[p_j]> are computed
 * for all blocks in parallel first (partial matrices per K range, added in fixed order).  The recurrence runs on a
 * CLUSTER of 8 CTAs per image, each owning 1/8 of the columns of every row (a single SM cannot stream the rows fast
 * enough); the per-block partial dot products meet in distributed shared memory.
 * gram: ipsr_blend_wide_gram_floats(B, M) floats of scratch. */
int ipsr_blend_wide_gram_floats(int B, int M);
int ipsr_blend_wide_blocked(const float* rows, const float* inv_norm, const float* vmax, const int32_t* ind,
                            const int32_t* mask_idx, int B, int K, int P, int M, float* gram,
                            float* y, float* wn, float* wo, void* stream);
/* Gather + fold in one pass: the patch pasted at position q is y[b][rank[q]] when rank[q] >= 0, else
 * rows[b][ind[b][q]]. */
int ipsr_fold_patch_rows(const float* rows, const float* y, const int32_t* ind, const int32_t* rank,
                         int B, int C, int H, int W, int patch, int stride, int M, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * InnerCos / InnerCos2   (models/InnerCos.py:30-36, models/InnerCos2.py:34-41)
 *   loss = mean_{b, c < c_limit, q} crit(x[b,c,q]*mask[q]*strength - target[b,c,q]);
 *   crit: 0 = squared error (MSELoss), 1 = absolute error (L1Loss).
 *   x is [B,C_total,N]; target is [B,c_limit,N].  partials: 1024 floats; ticket: one zeroed u32
 *   (left zeroed on return).
 * ------------------------------------------------------------------------------------------- */
int innercos_loss_fwd(const float* x, const float* mask_f32, const float* target,
                      int B, int C_total, int c_limit, int N, float strength, int crit,
                      float* partials, uint32_t* ticket, float* loss, void* stream);
/* grad_x[b,c,q] (c < c_limit; channels beyond are zeroed) = dloss/dx * grad_loss[0]. */
int innercos_loss_bwd(const float* x, const float* mask_f32, const float* target, const float* grad_loss,
                      int B, int C_total, int c_limit, int N, float strength, int crit,
                      float* grad_x, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused forward: (a) -> (b,c) -> recheck -> (d) [-> routes/exceptions], one call, one stream,
 * no host sync, CUDA-graph capturable.
 * ------------------------------------------------------------------------------------------- */
typedef struct ipsr_fwd_args {
  const float* x;          /* [B,C,H,W] layer input                                      */
  const float* ref;        /* [B,C,H,W] ref.relu4_3                                      */
  const int32_t* flag;     /* [N]                                                        */
  const int32_t* mask_idx; /* [M]                                                        */
  const int32_t* rank;     /* [N]                                                        */
  int32_t B, C, H, W, M;
  int32_t mode;            /* IPSR_MODE_*                                                */
  int32_t need_grad;       /* build backward routes / exceptions                         */
  int32_t col_begin, col_end; /* bank column shard [begin, end); end < 0 (or 0, 0): the whole bank; begin == end > 0: an
                               * EMPTY shard (stop_after_corr only): identity keys for the exchange */
  int32_t stop_after_corr; /* bank-sharded mode: stop after (b,c) with packed keys ready */
  int32_t psplit;          /* column splits per row tile on the tensor path (<=0: auto)  */
  int32_t exc_cap;
  float tol_rel, tol_abs;  /* recheck threshold after the three-pass split (<0: default) */
  float* out;              /* [B,C,H,W]                                                  */
  /* saved for backward (caller-owned) */
  int32_t* ind;            /* [B,N]                                                      */
  float* wn; float* wo;    /* [B,M]                                                      */
  int32_t* route_ptr;      /* [B,N+1]                                                    */
  int32_t* route_q;        /* [B,N]                                                      */
  int32_t* exc_start; int32_t* exc_cnt;   /* [B,N]                                       */
  int32_t* exc_l; float* exc_w;           /* [exc_cap] pool shared by the batch          */
  int32_t* exc_total;      /* [2B+2] exception state (ipsr_build_exceptions)             */
  int32_t* nrecheck_out;   /* [B] optional: rows that took the exact path (diagnostics)  */
  int32_t* npass2_out;     /* [B] optional: rows redone by the three-pass split          */
  void* ev_corr_begin;     /* optional cudaEvent_t pair recorded on `stream` around the      */
  void* ev_corr_end;       /*   correlation kernel ((b,c)), for live roofline measurement    */
  void* workspace; size_t workspace_bytes;
  /* Per-image masks -- an EXTENSION of the reference, whose 2-D mask is shared by the batch (IPSRFunction.py:32):
   * mask_stride = 0: flag [N], mask_idx [M], rank [N] serve every image (reference behaviour);
   * mask_stride = N: flag, mask_idx and rank are [B][N] (row b for image b; mask_idx rows hold m_count[b] valid
   *   entries), m_count [B] is the number of masked positions per image and M = max_b m_count[b] (the row stride of
   *   wn, wo and of the workspace's per-step buffers). */
  int32_t mask_stride;
  const int32_t* m_count;
  /* Optional: the InnerCos side loss that follows the layer in the generator (models/networks.py:347 `ipsr, innerCos,
   * downnorm_3`; models/InnerCos.py:30-36) computed in the paste while the pasted tiles are still in shared memory:
   * *cos_loss = mean(crit(out * cos_mask * cos_strength - cos_target)).  cos_target NULL: not computed. */
  const float* cos_target;   /* [B,C,H,W]                                                */
  const float* cos_mask;     /* [H*W] float, 1 = hole                                    */
  float cos_strength;
  int32_t cos_crit;          /* 0 = squared error (MSELoss), 1 = absolute error (L1Loss) */
  float* cos_partials;       /* ipsr_paste_loss_partials(B, C, N) floats                 */
  uint32_t* cos_ticket;      /* one zeroed u32 (left zeroed)                             */
  float* cos_loss;           /* scalar                                                   */
} ipsr_fwd_args;

size_t ipsr_workspace_bytes(int B, int C, int H, int W, int M, int mode);
int ipsr_shift_forward(const ipsr_fwd_args* args, void* stream);
/* Second half of the fused forward for the bank-sharded mode: after the caller all-reduced
 * (MAX) the packed keys returned by ipsr_workspace_packed, unpack them and run (d). */
int ipsr_shift_forward_finish(const ipsr_fwd_args* args, void* stream);
/* Address inside the workspace of the [B,N] int64 keys the bank-sharded exchange reduces. */
int64_t* ipsr_workspace_packed(const ipsr_fwd_args* args);

#ifdef __cplusplus
}
#endif
#endif /* IPSR_SM100_H_ */
