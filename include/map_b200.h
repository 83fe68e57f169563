/* map_b200.h — C ABI of libmap_b200.so: the B200 (sm_100a) kernels behind MAP's pretraining / finetuning step.
 *
 * The reference (CHIANGEL/MAP-CODE) is pure Python on torch eager; it has no FFI layer.  The boundary that replaces its
 * ATen / cuBLAS dispatches is this header.  Each entry point cites the reference call site it replaces
 * (paths relative to the reference root, `code/...:line`).
 *
 * Conventions
 *   - every function returns 0 on success, a negative MAP_E* code on failure; map_last_error() gives the text
 *     (thread-local).  Nothing here falls back to the CPU: a launch failure is an error, not a slow path.
 *   - the CALLER owns every buffer.  Device functions never allocate, never synchronise, never touch host memory
 *     (except the `const map_gemm_args*` struct, read before the launch) and launch only on `stream`:
 *     they are CUDA-graph capturable and safe to call concurrently on different streams.
 *   - `stream` is a cudaStream_t passed as void*.
 *   - ids are int64 (the reference's dtype), floating point is IEEE fp32.  Tensor-core GEMMs multiply in TF32
 *     (10-bit mantissa) and accumulate in fp32; see DESIGN.md for the tolerance.
 *   - RNG: Philox4x32-10, key = seed (lo,hi), counter = (elem_lo, elem_hi, offset_lo, offset_hi): the result for
 *     element `elem` does not depend on grid shape or on how many GPUs the batch is split over.
 *     bounded(u64 r, n) = mulhi64(r, n), r = w0 | w1<<32 ; bounded32(w0, n) = (w0*n)>>32 ; u01 = (w2>>8)*2^-24.
 */
#ifndef MAP_B200_H
#define MAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAP_B200_ABI_VERSION 1

#define MAP_OK 0
#define MAP_EINVAL (-1)   /* bad argument (shape, alignment, null pointer, unsupported size) */
#define MAP_ECUDA (-2)    /* CUDA runtime / driver error (text in map_last_error) */
#define MAP_EWORKSPACE (-3) /* workspace too small */
#define MAP_EUNSUPPORTED (-4)

typedef void* map_stream_t;

int map_abi_version(void);
const char* map_last_error(void);
/* number of SMs of the current device (grid sizing on the host side) */
int map_sm_count(int* out);
/* *slot = %globaltimer (ns) when the stream reaches this point: a one-thread marker kernel.  Used by bench.py --timeline to
 * reconstruct the timeline of the captured multi-stream step (there is no nsys in the image). */
int map_timestamp_ns(unsigned long long* slot, map_stream_t stream);

/* ------------------------------------------------------------------ K1  embedding gather
 * replaces nn.Embedding forward: code/layers.py:98 (Embeddings.forward), code/models.py:139 (LR.embed_w),
 * code/nce/index_linear.py:99-100 (index_select of emb / bias rows).
 * out[i, :] = table[ids[i], :]  (bit-exact copy).  D % 4 == 0 uses 16-byte lanes; any D >= 1 is accepted.
 * ids outside [0, V) write a zero row and set *oob_flag = 1 if oob_flag != NULL (the reference raises IndexError). */
int map_emb_gather_f32(const float* table, int64_t V, int D, const int64_t* ids, int64_t n_ids, float* out,
                       int32_t* oob_flag, map_stream_t stream);

/* ------------------------------------------------------------------ K2  deduplicated embedding backward
 * replaces aten::embedding_dense_backward (autograd of code/layers.py:98) and the index_add of the NCE tables
 * (autograd of code/nce/index_linear.py:99-100).  Pipeline: radix-sort (id, occurrence) -> unique rows + segment
 * starts -> segmented row sum into a COMPACT gradient [U, D] -> row-wise AdamW on the U touched rows (sparse) or an
 * exact dense sweep over all V rows (dense_exact = what transformers.AdamW does in the reference, trainer.py:75,140). */
size_t map_dedup_workspace_bytes(int64_t n_ids);
/* Sorts ids; writes uniq_ids[U] (ascending), seg_start[U+1] (positions in sorted order), occ_sorted[n] (original
 * occurrence index of each sorted position) and *n_unique (device int32).  key_bits = bits needed for V-1. */
int map_dedup_ids(const int64_t* ids, int64_t n_ids, int key_bits, int64_t* uniq_ids, int32_t* seg_start,
                  int32_t* occ_sorted, int32_t* n_unique, void* workspace, size_t workspace_bytes, map_stream_t stream);
/* grad_compact[u, :] = sum over occurrences o of segment u of  scale[o] * rows[(o / group) , :]
 * rows: [n_rows, D] with row stride ld_rows; scale == NULL means 1; group >= 1 (NCE: group = K+1, rows = input,
 * scale = dz; embedding: group = 1, rows = dY).  If scalar_out != NULL also scalar_out[u] = sum scale[o] (bias grad). */
int map_segment_reduce_rows(const float* rows, int64_t ld_rows, int D, const float* scale, int group,
                            const int32_t* occ_sorted, const int32_t* seg_start, const int32_t* n_unique,
                            int64_t n_ids, float* grad_compact, float* scalar_out, map_stream_t stream);
/* Extended forms used by the owner-side merge of the row-sharded tables (map_owned_compact below):
 *   n_dev      (may be NULL) device int32: only the first min(n_ids, *n_dev) keys exist (grids are sized for n_ids);
 *   seg_shift  keys that agree above bit `seg_shift` form ONE segment and uniq_ids = key >> seg_shift (the low bits only
 *              fix the order of the occurrences, e.g. by source rank);
 *   occ_map    (may be NULL) occurrence o of the sorted list refers to row code occ_map[o] instead of o;
 *   row_ptrs   (may be NULL) host array of n_peers device pointers: row code c lives at row_ptrs[c / rows_per_peer] + (c % rows_per_peer) * ld_rows
 *              (peer memory mapped with map_p2p_open; the loads travel over NVLink);
 *   pos_seg    (may be NULL) int32 [n_ids]: map_dedup_ids_ex writes the segment index of every sorted position; given to
 *              map_segment_reduce_rows_ex it replaces the per-tile binary search over seg_start.
 * map_dedup_ids(_ex) run the whole pipeline as ONE launch of persistent CTAs with grid barriers (one barrier per radix pass). */
int map_dedup_ids_ex(const int64_t* ids, int64_t n_ids, const int32_t* n_dev, int key_bits, int seg_shift, int64_t* uniq_ids,
                     int32_t* seg_start, int32_t* occ_sorted, int32_t* n_unique, int32_t* pos_seg, void* workspace,
                     size_t workspace_bytes, map_stream_t stream);
size_t map_dedup_debug_offset(int64_t n_ids); /* tuning aid: offset of the single-launch kernel's phase timestamps in the workspace */
int map_segment_reduce_rows_ex(const float* rows, int64_t ld_rows, int D, const float* scale, int group,
                               const int32_t* occ_sorted, const int32_t* seg_start, const int32_t* n_unique,
                               const int32_t* pos_seg, int64_t n_ids, const int32_t* n_dev, const int32_t* occ_map,
                               const float* const* row_ptrs, int n_peers, int64_t rows_per_peer, float* grad_compact,
                               float* scalar_out, map_stream_t stream);
/* dense[uniq_ids[u], :] = grad_compact[u, :]  (dense must be zero-filled by the caller): the `.grad` the reference sees */
int map_scatter_rows(const float* grad_compact, const int64_t* uniq_ids, const int32_t* n_unique, int64_t max_unique,
                     int D, float* dense, map_stream_t stream);

/* ------------------------------------------------------------------ K12 / K2  AdamW (transformers==4.26.1 semantics)
 * replaces transformers.AdamW.step (code/trainer.py:75-76,140):
 *   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= step_size * m / (sqrt(v)+eps) ; p -= lr*wd*p
 *   step_size = lr * sqrt(1-b2^t)/(1-b1^t).
 * hyper: device float[8] = {lr, step_size, beta1, beta2, eps, t, 1-beta1, 1-beta2}, produced by map_adamw_hyper_step so that a
 * captured CUDA graph picks up the new learning rate at every replay without host involvement.  Scalars are doubles
 * because the reference evaluates the bias corrections with Python floats (beta2 = 0.999 is not a float32). */
#define MAP_SCHED_CONST 0   /* transformers.get_constant_schedule_with_warmup */
#define MAP_SCHED_COSINE 1  /* transformers.get_cosine_schedule_with_warmup (num_cycles=0.5) */
/* step_counter (device int64) is incremented; lr = base_lr * lambda(step_counter_before) (trainer.py:141) */
int map_adamw_hyper_step(float* hyper, int64_t* step_counter, double base_lr, double beta1, double beta2, double eps,
                         int sched, int64_t warmup_steps, int64_t total_steps, map_stream_t stream);
/* explicit (host-driven) variant: step = 1-based step count for bias correction */
int map_adamw_hyper_set(float* hyper, double lr, double beta1, double beta2, double eps, int64_t step, map_stream_t stream);

typedef struct {
    float* p;        /* parameter */
    const float* g;  /* gradient  */
    float* m;        /* exp_avg   */
    float* v;        /* exp_avg_sq */
    float* p_t;      /* optional transposed copy of p ([cols, rows]) refreshed by the update, or NULL */
    int64_t n;       /* element count */
    int32_t rows, cols; /* only used when p_t != NULL (n == rows*cols) */
    float weight_decay;
    int32_t n_planes;        /* 0, or 1..3: the update also rewrites the parameter's bf16 planes (operand format of map_gemm_bf16s_group) */
    uint16_t* planes;        /* plane i at planes + i * plane_stride, same flat layout as p (n % 4 == 0 required) */
    int64_t plane_stride;
} map_adamw_tensor;
/* one launch over a device-resident list of tensors (dense parameters) */
int map_adamw_multi_tensor(const map_adamw_tensor* tensors_dev, int n_tensors, int64_t max_elems_per_tensor,
                           const float* hyper, map_stream_t stream);
/* sparse: only the U touched rows (north_star subsystem 1: "fused sparse optimizer update") */
int map_adamw_sparse_rows(float* table, float* m, float* v, int D, const int64_t* uniq_ids, const float* grad_compact,
                          const int32_t* n_unique, int64_t max_unique, const float* hyper, float weight_decay,
                          map_stream_t stream);
/* dense_exact: all V rows, gradient of untouched rows = 0 (bit-comparable with the reference's dense AdamW) */
int map_adamw_dense_rows_sparse_grad(float* table, float* m, float* v, int64_t V, int D, const int64_t* uniq_ids,
                                     const float* grad_compact, const int32_t* n_unique, const float* hyper,
                                     float weight_decay, map_stream_t stream);

/* ------------------------------------------------------------------ K8 / K9  dynamic_mask
 * replaces Trainer.dynamic_mask (code/trainer.py:217-266).  One warp per sample row; duplicate masked fields resolve
 * last-writer-wins like the reference's CPU scatter.  row0 = global index of the first row (row-sharded runs). */
#define MAP_SAMPLING_RANDINT 0 /* trainer.py:224-225 */
#define MAP_SAMPLING_NORMAL 1  /* trainer.py:222-223 randperm(F)[:L] */
/* step_dev (nullable, device int64): offset_eff = offset + 8 * (*step_dev) — lets a captured CUDA graph draw a new
 * Philox subsequence at every replay (8 = streams per step; the counter is advanced by map_adamw_hyper_step). */
int map_mask_index_philox(int64_t* masked_index, int64_t B, int L, int F, int sampling_method, uint64_t seed,
                          uint64_t offset, int64_t row0, const int64_t* step_dev, map_stream_t stream);
/* MFP (trainer.py:229-233): labels[b,l] = ids[b, mi[b,l]] ; ids_out = ids with masked fields set to mask_id (3) */
int map_mfp_mask_apply(const int64_t* ids, const int64_t* masked_index, int64_t B, int F, int L, int64_t mask_id,
                       int64_t* ids_out, int64_t* labels, map_stream_t stream);
#define MAP_RFD_UNIGRAM 0        /* trainer.py:235-240 */
#define MAP_RFD_UNIFORM 1        /* trainer.py:241-246 */
#define MAP_RFD_WHOLE_UNIFORM 2  /* trainer.py:247-252 */
#define MAP_RFD_WHOLE_UNIGRAM 3  /* trainer.py:253-260 */
/* RFD: draws the replacement per (b,l), scatters it, labels[b,f] = (ids[b,f] != ids_out[b,f]) as float */
int map_rfd_replace_philox(const int64_t* ids, const int64_t* masked_index, int64_t B, int F, int L, int mode,
                           const int64_t* x_train, int64_t n_train, const int64_t* idx_low, const int64_t* idx_high,
                           int64_t input_size, uint64_t seed, uint64_t offset_replace, uint64_t offset_field2,
                           int64_t row0, const int64_t* step_dev /* nullable */, int64_t* ids_out, float* labels,
                           int64_t* replace_feat_out /* nullable */, map_stream_t stream);

/* ------------------------------------------------------------------ K5  alias sampler
 * map_alias_build replaces the O(V) Python loop of AliasMultinomial.__init__ (code/nce/alias_multinomial.py:40-73):
 * HOST function, float32 arithmetic, same scan order and LIFO stacks => bit-identical tables. */
int map_alias_build(const float* probs, int64_t V, float* out_prob, int64_t* out_alias);
/* replaces AliasMultinomial.draw (code/nce/alias_multinomial.py:81-97): out[e] for e in [0,n), Philox element elem0+e */
int map_alias_draw_philox(const float* prob, const int64_t* alias, int64_t V, uint64_t seed, uint64_t offset,
                          int64_t elem0, int64_t n, const int64_t* step_dev /* nullable */, int64_t* out, map_stream_t stream);

/* ------------------------------------------------------------------ K6 / K7  fused NCE head
 * replaces NCELoss.forward + IndexLinear._compute_sampled_logit + nce_loss / sampled_softmax_loss
 * (code/nce/nce_loss.py:79-144,201-244, code/nce/index_linear.py:68-106).  Position n reads input[n*P .. n*P+P).
 * Outputs (ids_out, d_input, acc_count nullable):
 *   logits[N,K+1]   = <input, emb[idx]> + bias[idx] - norm_term              (nce_loss.py:171-172)
 *   ids_out[N,K+1]  = [target | noise]                                        (nce_loss.py:138)
 *   loss_pos[N]     per-position loss (nce: sum_j BCEWithLogits ; sampled: CE with label 0)
 *   dz[N,K+1]       grad_scale * d(loss_pos[n])/d(logit[n,j])   (grad_scale = 1/N_global for the mean reduction)
 *   d_input[N,P]    grad_scale * d(loss_pos[n])/d(input[n,:])
 *   acc_count       += #positions with argmax_j logits == 0                   (models.py:77)
 * P in {4,8,16,32,64,128}. */
#define MAP_NCE_LOSS_NCE 0
#define MAP_NCE_LOSS_SAMPLED 1
int map_nce_fwd(const float* input, int64_t N, int P, int K, const int64_t* target, const int64_t* noise,
                const float* emb, const float* bias, const float* logprob_noise, int64_t V, float norm_term,
                int loss_type, float grad_scale, float* logits, int64_t* ids_out, float* loss_pos, float* dz,
                float* d_input, int32_t* acc_count, map_stream_t stream);
/* The backward of the slice gather as a DENSE, deterministic pass: d_enc[b, f*P + p] = sum_{l: masked_index[b,l] == f} d_input[b,l,p]
 * (zero elsewhere) for every element of d_enc [B, F*P] — no pre-zeroing, no atomics (map_scatter_add_slices needs both);
 * optionally also the bf16 planes of d_enc (operand format of map_gemm_bf16s_group).  P % 4 == 0. */
int map_expand_slices(const float* d_input, const int64_t* masked_index, int64_t B, int L, int F, int P, float* d_enc, int64_t ld_enc,
                      uint16_t* planes, int64_t ld_p, int64_t plane_stride, int n_planes, map_stream_t stream);
/* Full-softmax cross entropy over the whole vocabulary: replaces IndexLinear.ce_loss (code/nce/index_linear.py:145-151, the
 * `loss_type != nce/sampled` fallback of NCELoss.forward, nce_loss.py:133-135): loss_pos[n] = logsumexp_v(<input[n], emb[v]> +
 * bias[v]) - (<input[n], emb[target[n]]> + bias[target[n]]).  Forward only (evaluation); the [N, V] score matrix of the
 * reference is never materialised.  P <= 64. */
size_t map_nce_full_ce_workspace_bytes(int64_t N, int64_t V);
int map_nce_full_ce(const float* input, int64_t N, int P, const float* emb, const float* bias, int64_t V, const int64_t* target,
                    float* loss_pos, void* workspace, size_t workspace_bytes, map_stream_t stream);
/* sel[n,:] = enc[(b*F + masked_index[n])*P + :], n = b*L + l   (torch.gather of the masked slices, models.py:75) */
int map_gather_slices(const float* enc, const int64_t* masked_index, int64_t N, int L, int F, int P, float* sel,
                      map_stream_t stream);
/* d_enc[(b*F + masked_index[n])*P + :] += d_input[n, :]   (autograd of torch.gather, models.py:75; d_enc pre-zeroed) */
int map_scatter_add_slices(const float* d_input, const int64_t* masked_index, int64_t N, int L, int F, int P,
                           float* d_enc, map_stream_t stream);
/* ------------------------------------------------------------------ K16  MFP feature encoder evaluated BY FIELD
 * replaces `enc = feat_encoder(final).view(B, F, P)` + `torch.gather(enc, 1, masked_index)` (code/models.py:73-78) and their
 * autograd: only the L masked P-wide slices of the [B, F*P] encoder output are ever used, so forward / dgrad / wgrad are computed
 * for those slices alone (F/L times fewer flops, no [B, F*P] tensor).  n = b*L + l indexes the masked positions,
 * f(n) = masked_index[n], W_f = rows [f*P, (f+1)*P) of feat_encoder.weight [F*P, K].  Exact fp32 FMAs, deterministic order.
 * F <= 256, P in {8, 16, 32, 64} (map_field_enc_supported), K and all strides multiples of 4 floats.
 *
 * map_field_bucket: stable counting sort of the positions by field: perm[s] = n of the s-th position in (field, n) order,
 *   fstart[f] .. fstart[f+1] = the slots of field f (fstart has F+1 entries).
 * map_field_enc_fwd:   sel[n, :] = X[n / L, 0:K] . W_f(n)^T + bias[f(n)*P : (f(n)+1)*P]              (X = `final` [B, ldx])
 * map_field_enc_dgrad: dxpos[n, 0:K] = d_sel[n, :] . W_f(n)     (one row per POSITION; map_head_bwd_fold sums the L rows of a sample)
 * map_field_enc_wgrad: dW[f*P + j, 0:K] = sum_{n: f(n) = f} d_sel[n, j] * X[n / L, 0:K];  dbias[f*P + j] = sum_{n: f(n) = f} d_sel[n, j]
 *   (every row of dW / dbias is written: fields without masked positions get zeros; dbias may be NULL) */
int map_field_enc_supported(int F, int P);
int map_field_bucket(const int64_t* masked_index, int64_t N, int F, int32_t* perm, int32_t* fstart, map_stream_t stream);
int map_field_enc_fwd(const float* X, int64_t ldx, int K, const float* W, int64_t ldw, const float* bias, const int32_t* perm,
                      const int32_t* fstart, int64_t N, int L, int F, int P, float* sel, map_stream_t stream);
int map_field_enc_dgrad(const float* d_sel, const float* W, int64_t ldw, int K, const int32_t* perm, const int32_t* fstart,
                        int64_t N, int F, int P, float* dxpos, int64_t ld_dx, map_stream_t stream);
int map_field_enc_wgrad(const float* d_sel, const float* X, int64_t ldx, int K, const int32_t* perm, const int32_t* fstart, int64_t N,
                        int L, int F, int P, float* dW, int64_t ldw, float* dbias, map_stream_t stream);
/* Fold of the per-position gradients + the first stage of the towers' backward (what the epilogues of the dense head dgrad GEMMs
 * do: MAP_EPI_CROSS_BWD with no layer above, MAP_EPI_MUL_RELUMASK, fused bias column sums, bf16 operand planes), one pass:
 *   g[b, c] = sum_l dxpos[b*L + l, c]                                            (fixed order l = 0 .. L-1)
 *   CrossNet columns [cross_col0, +cross_w): g_out = g ; du_out = g * x0 ; dx0_out = g * u ; cross_bias_grad += colsum(du_out)
 *   ReLU columns     [relu_col0,  +relu_w):  dz_out = g * (y > 0) ;                         relu_bias_grad  += colsum(dz_out)
 *   scalar_col (>= 0): scalar_out[b * ld_scalar] = g[b, scalar_col]               (DeepFM: d(lr_fm), code/models.py:224)
 * Region pointers are indexed from the region's first column; the *_bias_grad accumulate with fp32 reductions (pre-zeroed by the
 * caller, may be NULL); planes are bf16 [n_planes][B][pl_ld] (operand format of map_gemm_bf16s_group), may be NULL. */
typedef struct {
    const float* dxpos; int64_t ld_dx;
    int64_t B; int32_t L; int32_t ncols;
    int32_t cross_col0, cross_w;
    const float* x0; int64_t ld_x0;
    const float* u; int64_t ld_u;
    float* g_out; int64_t ld_g;
    float* du_out; int64_t ld_du;
    float* dx0_out; int64_t ld_dx0;
    uint16_t* du_planes; int64_t du_pl_ld; int64_t du_pl_stride; int32_t du_nplanes; int32_t reserved0_;
    float* cross_bias_grad;
    int32_t relu_col0, relu_w;
    const float* y; int64_t ld_y;
    float* dz_out; int64_t ld_dz;
    uint16_t* dz_planes; int64_t dz_pl_ld; int64_t dz_pl_stride; int32_t dz_nplanes; int32_t scalar_col;
    float* relu_bias_grad;
    float* scalar_out; int64_t ld_scalar;
} MapHeadBwdArgs;
int map_head_bwd_fold(const MapHeadBwdArgs* args, map_stream_t stream);
/* ------------------------------------------------------------------ K17  Compressed Interaction Network (xDeepFM)
 * replaces CIN.forward (code/layers.py:709-721) and its autograd.  One layer of the reference is
 *   hadamard[b, h*M + m, d] = X0[b,h,d] * Xi[b,m,d];  X_next = Conv1d(F*M -> O, kernel 1)(hadamard);  pooled = X_next.sum(-1).
 * The 1x1 convolution is a GEMM over the pairs p = b*D + d, so CIN activations are kept PAIR-MAJOR ([B*D, channels]) and the
 * contraction runs on map_gemm_bf16s_group; these entry points are the glue around it (exact fp32, deterministic):
 * map_cin_relayout:     to_pairs=1: src [B, C, D] -> dst [(b*D + d), c] (leading dimension ld_pairs);
 *                       to_pairs=0: src pairs -> dst [B, C, D] (accumulate=1 adds into dst)
 * map_cin_hadamard_fwd: z[p, h*M + m] = x0[p, h] * xi[p, m]; columns [F*M, ldz) are written as zeros (K padding of the GEMM)
 * map_cin_hadamard_bwd: dx0[p, h] (+)= sum_m dz[p, h*M + m] * xi[p, m];  dxi[p, m] = sum_h dz[p, h*M + m] * x0[p, h]
 * map_cin_pool_fwd:     pooled[b, o] = sum_d y[b*D + d, o]
 * map_cin_pool_bwd:     dy[b*D + d, o] = d_pooled[b, o] (+ d_next[b*D + d, o] if d_next != NULL); columns [O, ld_dy) zeroed */
int map_cin_relayout(const float* src, float* dst, int64_t B, int C, int D, int64_t ld_pairs, int to_pairs, int accumulate,
                     map_stream_t stream);
int map_cin_hadamard_fwd(const float* x0, int64_t ld_x0, const float* xi, int64_t ld_xi, int64_t P, int F, int M, float* z,
                         int64_t ldz, map_stream_t stream);
int map_cin_hadamard_bwd(const float* dz, int64_t ldz, const float* x0, int64_t ld_x0, const float* xi, int64_t ld_xi, int64_t P,
                         int F, int M, float* dx0, int64_t ld_dx0, int accumulate_dx0, float* dxi, int64_t ld_dxi,
                         map_stream_t stream);
int map_cin_pool_fwd(const float* y, int64_t ldy, int64_t B, int D, int O, float* pooled, int64_t ld_pooled, map_stream_t stream);
int map_cin_pool_bwd(const float* d_pooled, int64_t ld_pooled, const float* d_next, int64_t ld_next, int64_t B, int D, int O,
                     float* dy, int64_t ld_dy, map_stream_t stream);
/* deterministic mean/sum: out[0] = scale * sum(x[0..n)) */
int map_reduce_sum_f32(const float* x, int64_t n, float scale, float* out, void* workspace, size_t workspace_bytes,
                       map_stream_t stream);
size_t map_reduce_workspace_bytes(int64_t n);

/* ------------------------------------------------------------------ K10  BCE-with-logits heads
 * replaces BCEWithLogitsLoss + the accuracy / positive-ratio reductions of the RFD head (code/models.py:80-84) and the
 * CTR loss (code/models.py:91-92).  stats_out[4] = {mean loss, #correct ((sigmoid>0.5)==label), sum(labels), n};
 * dlogits[i] = (sigmoid(z_i) - y_i)/n. */
int map_bce_logits_fwd(const float* logits, const float* labels, int64_t n, float* stats_out, float* dlogits,
                       void* workspace, size_t workspace_bytes, map_stream_t stream);

/* On-device ROC AUC for Trainer.eval (replaces sklearn.metrics.roc_auc_score on host copies, code/trainer.py:183-197):
 * keys[i] = order-preserving 32-bit image of scores[i] (as int64, the id type of map_dedup_ids; key_bits = 32); after
 * map_dedup_ids(keys) and map_segment_reduce_rows(labels as [n,1] rows) -> pos_count[u] positives in tie group u,
 * map_auc_rank_sum leaves out2 = {sum of the average 1-based ranks of the positives, number of positives P} (fp64), so that
 * AUC = (out2[0] - P(P+1)/2) / (P (n - P)). */
int map_float_sort_keys(const float* scores, int64_t n, int64_t* keys, map_stream_t stream);
int map_auc_rank_sum(const int32_t* seg_start, const float* pos_count, const int32_t* n_unique, int64_t max_unique, double* out2,
                     map_stream_t stream);

/* ------------------------------------------------------------------ K11  FM second-order + LR term (DeepFM)
 * replaces LR.forward (code/models.py:137-143) + InnerProductLayer product_sum (code/layers.py:125-131):
 * out[b] = sum_f w[ids[b,f]] + lr_bias + 0.5 * sum_d ((sum_f e[b,f,d])^2 - sum_f e[b,f,d]^2)
 * ids == NULL: lr_w holds one weight per occurrence [B*F] (already gathered, e.g. from a row-sharded table). */
int map_fm_lr_fwd(const float* feat_embed, const int64_t* ids, const float* lr_w, const float* lr_bias, int64_t B,
                  int F, int D, float* out, int64_t ld_out, map_stream_t stream);
/* d_embed[b,f,d] (+)= g[b] * (sum_f' e[b,f',d] - e[b,f,d]) ; d_w_occ[b*F+f] = g[b] (per-occurrence LR grad, fed to the
 * dedup pipeline) ; g has stride ld_g */
int map_fm_lr_bwd(const float* feat_embed, const float* g, int64_t ld_g, int64_t B, int F, int D, int accumulate,
                  float* d_embed, float* d_w_occ, map_stream_t stream);

/* ------------------------------------------------------------------ K3 / K4  GEMMs with fused epilogues
 * replaces every nn.Linear on the path: CrossNetV2 (code/layers.py:195-200), MLPBlock (code/layers.py:179-188),
 * feat_encoder / pred_rfd / fc_out (code/models.py:116-123,304) forward and backward.
 *   acc[m,n] = sum_k A[m,k] * B[n,k]                      (C = A . B^T; "transX" describe the storage of A / B)
 * A: logical [M,K]; trans_a=0 -> stored row-major [M,K] (lda = row stride); trans_a=1 -> stored [K,M].
 * B: logical [N,K]; trans_b=0 -> stored row-major [N,K] (an nn.Linear weight); trans_b=1 -> stored [K,N]. */
#define MAP_EPI_NONE 0          /* C = acc */
#define MAP_EPI_BIAS 1          /* C = acc + bias[n] */
#define MAP_EPI_BIAS_RELU 2     /* C = relu(acc + bias[n])                       (MLPBlock) */
#define MAP_EPI_CROSS 3         /* U = acc + bias[n]; aux_out = U; C = aux0 + aux1 * U   (aux0 = Xi, aux1 = X0) */
#define MAP_EPI_MUL_RELUMASK 4  /* C = acc * (aux0 > 0)                          (dX through a ReLU, aux0 = fwd output) */
#define MAP_EPI_ADD 5           /* C = acc + aux0                                 (residual gradient) */
#define MAP_EPI_ADD_MUL 6       /* C = (acc + aux0) * aux1 ; aux_out = (acc + aux0)  (next cross layer's dU from G) */
/* One whole CrossNetV2 backward stage in the dgrad GEMM's epilogue (autograd of code/layers.py:200, replaces the separate
 * elementwise pass):  G = acc + aux0 (aux0 = G of the layer above, NULL for the first stage) ; aux_out = G ;
 * C = G * aux1 (dU of the layer below, aux1 = X0) ; acc_out (+)= G * aux2 (dX0 accumulator, aux2 = U of the layer below;
 * acc_accumulate = 0 stores, 1 adds with fp32 reductions). */
#define MAP_EPI_CROSS_BWD 7
#define MAP_EPI_ADD3 8          /* C = acc + aux0 + aux1 (+ aux2 if not NULL)     (dE = G0 + dX0_cross + dX0_mlp) */
typedef struct {
    int32_t M, N, K;
    int32_t trans_a, trans_b;
    int32_t epilogue;
    const float* A; int64_t lda;
    const float* B; int64_t ldb;
    float* C; int64_t ldc;
    const float* bias;
    const float* aux0; int64_t ld_aux0;
    const float* aux1; int64_t ld_aux1;
    float* aux_out; int64_t ld_aux_out;
    const float* aux2; int64_t ld_aux2;
    float* acc_out; int64_t ld_acc_out;   /* MAP_EPI_CROSS_BWD only */
    int32_t acc_accumulate;
    int32_t reserved_;
    /* optional, any epilogue: colsum_out[n] += sum_m C[m,n] with fp32 reductions (bias gradient of the layer whose
     * pre-activation gradient this GEMM writes); the caller zeroes it beforehand.  Not with split-K. */
    float* colsum_out;
} map_gemm_args;
/* bias / aux0 / aux1 are read through the read-only (non-coherent) path: they must not alias C or aux_out. */
/* exact fp32 on CUDA cores: skinny shapes (N < 64: pred_rfd.2, fc_out) and the fp32 cross-check of the TF32 path */
int map_gemm_f32_simt(const map_gemm_args* args, map_stream_t stream);
/* tcgen05 (kind::tf32) + TMA + TMEM, fp32 accumulate.  Requires 16-byte aligned base pointers and lda/ldb % 4 == 0. */
int map_gemm_tf32_tcgen05(const map_gemm_args* args, map_stream_t stream);
/* Up to 4 INDEPENDENT problems (none reads what another writes) in ONE persistent launch: one CTA per SM walks the tiles of
 * all problems, two TMEM accumulators overlap each tile's epilogue with the next tile's main loop.  Same per-problem
 * contract as map_gemm_tf32_tcgen05 (every problem must satisfy map_gemm_tf32_supported).  This is how the step issues the
 * CrossNet / MLP layers of one depth (code/layers.py:187-201 run them back to back) and the dgrad + wgrad GEMMs that consume
 * the same upstream gradient. */
int map_gemm_tf32_group(const map_gemm_args* args, int count, map_stream_t stream);
/* 1 if map_gemm_tf32_tcgen05 accepts these arguments (shape / alignment), else 0 */
int map_gemm_tf32_supported(const map_gemm_args* args);
/* ---- split-bf16 tensor-core GEMM (the default backend of the step; gemm_bf16s.cu).
 * Every fp32 operand x is carried as bf16 planes hi = bf16(x), lo = bf16(x - hi), lo2 = bf16(x - hi - lo) in the SAME storage
 * layout as x (plane i at base + i * plane_stride elements, row stride `ld` elements; 16-byte aligned, strides % 8 == 0).
 * terms = 3: hi*hi + hi*lo + lo*hi into one fp32 accumulator (~2^-17 per product; needs 2 planes of A and B);
 * terms = 6: + lo*lo + hi*lo2 + lo2*hi (~2^-24, fp32-level; needs 3 planes) — used for the forward GEMMs whose output feeds a
 * ReLU: d(loss)/d(weights) is discontinuous in those pre-activations.  g.A / g.B are ignored; g carries the
 * shapes, trans flags, epilogue and fp32 outputs exactly as for map_gemm_tf32_tcgen05.  Optional c_planes: the epilogue also
 * writes C as c_nplanes bf16 planes (the operand format of the GEMMs that consume C next).
 * One call = up to 4 INDEPENDENT problems in one persistent launch of CTA pairs (cta_group::2) with two TMEM accumulators. */
typedef struct {
    map_gemm_args g;
    const uint16_t* a_planes; int64_t a_ld; int64_t a_plane_stride; int32_t a_nplanes; int32_t terms;
    const uint16_t* b_planes; int64_t b_ld; int64_t b_plane_stride; int32_t b_nplanes; int32_t reserved0_;
    uint16_t* c_planes; int64_t c_ld; int64_t c_plane_stride; int32_t c_nplanes; int32_t reserved1_;
} map_gemm_split_args;
int map_gemm_bf16s_group(const map_gemm_split_args* args, int count, map_stream_t stream);
/* Tuning aid (scripts/trace_gemm_bf16s.py): while dev_buf != NULL every CTA of the following map_gemm_bf16s_group launches writes a
 * 64 x uint64 record (entry / exit times, and per tile: first TMA issue, accumulator granted, first stage landed, last MMA issued,
 * accumulator complete, epilogue end) to dev_buf[64 * cta ...] if capacity_u64 covers the grid.  NULL switches it off. */
int map_gemm_bf16s_set_trace(unsigned long long* dev_buf, int64_t capacity_u64);
/* fp32 [rows, cols] (row stride ld) -> n_planes (1..3) bf16 planes [rows, ld_p]; for operands no kernel of ours produces
 * (parameters after load_state_dict, test inputs).  cols, ld, ld_p, plane_stride multiples of 4 elements. */
int map_split_bf16(const float* src, int64_t ld, int64_t rows, int64_t cols, uint16_t* planes, int64_t ld_p, int64_t plane_stride,
                   int n_planes, map_stream_t stream);
/* Tuning aid (scripts/trace_gemm.py): while dev_buf != NULL every CTA of the following tcgen05 GEMM launches (grids of at
 * most capacity_records CTAs) writes 8 uint64 phase timestamps to dev_buf[8 * linear_cta_id ...].  NULL switches it off. */
int map_gemm_set_trace(unsigned long long* dev_buf, int64_t capacity_records);

/* ------------------------------------------------------------------ K13  row-sharded tables (owner-side kernels)
 * The reference initialises NCCL (code/arguments.py:74) but never issues a collective; this is the multi-GPU path the
 * north_star asks for: table row `id` lives on rank id % R at local row id / R.  The exchange steps between these
 * kernels are fixed-size NCCL collectives issued by the host (map_code_b200/dist.py). */
/* out[i,:] = (ids[i] % R == rank) ? shard[ids[i] / R, :] : 0   — summed over ranks (reduce-scatter) this is the lookup */
int map_emb_gather_owned_f32(const float* shard, int64_t shard_rows, int D, const int64_t* ids, int64_t n_ids, int R,
                             int rank, float* out, map_stream_t stream);
/* keys[i] = owned ? ids[i] / R : sentinel  — local row ids for the dedup pipeline; every foreign occurrence lands in one
 * trailing segment (sentinel = shard_rows, a dummy row) */
int map_owned_keys(const int64_t* ids, int64_t n, int R, int rank, int64_t sentinel, int64_t* keys, map_stream_t stream);
/* partial[n,j] = owned(idx[n,j]) ? <q[n,:], emb_shard[idx/R,:]> + bias_shard[idx/R] : 0   (K1 = K+1 columns) */
int map_nce_scores_owned(const float* q, int64_t N, int P, int K1, const int64_t* idx, const float* emb_shard,
                         const float* bias_shard, int R, int rank, float* partial, map_stream_t stream);
/* the loss half of map_nce_fwd on complete scores: logits = scores - norm_term, loss_pos, dz (x grad_scale), acc_count */
int map_nce_loss_from_scores(const float* scores, const int64_t* idx, int64_t N, int K1, const float* logprob_noise,
                             float norm_term, int loss_type, float grad_scale, float* logits, float* loss_pos, float* dz,
                             int32_t* acc_count, map_stream_t stream);
/* d_q[n,:] = sum over owned j of dz[n,j] * emb_shard[idx[n,j]/R, :]   — summed over ranks this is d(loss)/d(query) */
int map_nce_dinput_owned(const float* dz, int64_t N, int P, int K1, const int64_t* idx, const float* emb_shard, int R,
                         int rank, float* d_q, map_stream_t stream);

/* ------------------------------------------------------------------ K13b row-sharded tables over NVLink peer memory
 * (the default multi-GPU path; the collective-based owner kernels above remain as the portable variant).
 * Every rank allocates its shards / compact gradients with map_p2p_alloc and maps the other ranks' buffers with
 * map_p2p_open (CUDA IPC, lazy peer access); kernels then take a HOST array of R device base pointers (R <= 8). */
int map_p2p_alloc(size_t bytes, void** ptr, unsigned char* handle64);  /* cudaMalloc (zero-filled) + 64-byte IPC handle */
int map_p2p_open(const unsigned char* handle64, void** ptr);           /* another process' allocation -> local address */
int map_p2p_close(void* ptr);
int map_p2p_free(void* ptr);
/* out[i,:] = shard[ids[i] % R][ids[i] / R, :] — replaces nn.Embedding forward (code/layers.py:98) on a sharded table: the
 * lookup itself is the exchange (remote rows are read over NVLink), there is no id / row all-to-all. */
int map_emb_gather_sharded_f32(const void* const* shard_ptrs, int R, int64_t V, int D, const int64_t* ids, int64_t n_ids,
                               float* out, int32_t* oob_flag, map_stream_t stream);
/* map_nce_fwd on row-sharded output tables (emb [V,P], bias [V,1] split by id % R); logprob_noise is replicated. */
int map_nce_fwd_sharded(const float* input, int64_t N, int P, int K, const int64_t* target, const int64_t* noise,
                        const void* const* emb_shards, const void* const* bias_shards, int R, const float* logprob_noise,
                        float norm_term, int loss_type, float grad_scale, float* logits, int64_t* ids_out, float* loss_pos,
                        float* dz, float* d_input, int32_t* acc_count, map_stream_t stream);
/* ids_out[n, 0] = target[n], ids_out[n, 1 + k] = noise[n, k] (the [N, K+1] id list map_nce_fwd emits as ids_out): lets the sort of
 * the NCE tables' gradient ids start as soon as the noise is drawn (code/nce/index_linear.py:79-83 builds the same cat). */
int map_nce_ids_concat(const int64_t* target, const int64_t* noise, int64_t N, int K, int64_t* ids_out, map_stream_t stream);
/* Stream-ordered barrier over the R ranks through peer memory: flag_ptrs[r] = rank r's uint32 flags[n_sites][8] from
 * map_p2p_alloc (zero-initialised), epochs = this rank's device uint32[n_sites] (zero-initialised, advanced by the kernel, so a
 * captured graph replays it), site = which barrier of the step's schedule (barriers of different sites may overlap in time),
 * error_word (may be NULL) is set to 1 if a peer does not arrive within seconds instead of hanging the GPU.  Everything earlier
 * kernels of `stream` wrote on any rank is visible to peer loads of kernels launched after the barrier. */
int map_p2p_barrier(const void* const* flag_ptrs, int R, int rank, int site, int n_sites, uint32_t* epochs, uint32_t* error_word,
                    map_stream_t stream);
/* Owner-side merge, step 1: scan the R per-rank unique-id lists (uniq_ptrs[s] = int64 ids ascending, n_unique_ptrs[s] =
 * device int32 count) and append the entries this rank owns: keys[k] = ((id / R) << ceil_log2(R)) | s, src[k] = s * cap + u,
 * *n_out = number of entries.  Then map_dedup_ids_ex(keys, R * cap, n_out, key_bits, ceil_log2(R), ...) and
 * map_segment_reduce_rows_ex(..., occ_map = src, row_ptrs = the ranks' compact gradients, rows_per_peer = cap). */
int map_owned_compact(const void* const* uniq_ptrs, const void* const* n_unique_ptrs, int R, int rank, int64_t cap, int64_t* keys,
                      int32_t* src, int32_t* n_out, map_stream_t stream);

/* Dense-gradient all-reduce over NVLink peer memory (the data-parallel gradient exchange the reference's NCCL initialisation at
 * code/arguments.py:74 stands for; replaces ncclAllReduce on the gradient buckets of the sharded step).  send_ptrs[q] = rank q's
 * peer-visible copy of the flat gradient.  map_p2p_reduce_f32: out[i] = sum over q = 0..R-1 (in that order on every rank: the
 * result is bit-identical everywhere) of send_ptrs[q][first + i], i in [0, count).  One shot: every rank reduces the whole range
 * into its local gradient.  Two shot: rank r reduces slice r in place (out = its own send buffer + first), then, after a barrier,
 * map_p2p_gather_slices_f32 collects out[i] = send_ptrs[i / slice][first + i].  first / count / slice in floats, multiples of 4.
 * The caller brackets the calls with map_p2p_barrier (copies complete before, reduced slices complete between). */
int map_p2p_reduce_f32(const void* const* send_ptrs, int R, int64_t first, int64_t count, float* out, map_stream_t stream);
int map_p2p_gather_slices_f32(const void* const* send_ptrs, int R, int64_t first, int64_t count, int64_t slice, float* out,
                              map_stream_t stream);

/* column sums: out[n] = sum_m X[m,n]  (bias gradients).  deterministic two-stage. */
int map_colsum_f32(const float* X, int64_t ldx, int64_t M, int N, float* out, void* workspace, size_t workspace_bytes,
                   map_stream_t stream);
size_t map_colsum_workspace_bytes(int64_t M, int N);
/* CrossNet backward elementwise stage (autograd of layers.py:200):  dU = G * X0 ; dX0_acc (+)= G * U */
int map_cross_bwd_pre(const float* G, int64_t ldg, const float* X0, int64_t ldx0, const float* U, int64_t ldu,
                      int64_t M, int N, int accumulate, float* dU, float* dX0_acc, map_stream_t stream);
/* out = a + b (+ c)  elementwise over [M,N] with row strides */
int map_add3_f32(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c, int64_t ldc, int64_t M, int N,
                 float* out, int64_t ldo, map_stream_t stream);
/* out = dy * (y > 0): ReLU backward outside a GEMM epilogue (autograd of nn.ReLU, layers.py:58) */
int map_relu_bwd_f32(const float* dy, int64_t lddy, const float* y, int64_t ldy, int64_t M, int N, float* out, int64_t ldo,
                     map_stream_t stream);
/* out[i] = x[i] * scalar_dev[0]: chain rule through a scalar loss whose upstream gradient lives on the device */
int map_scale_by_scalar_f32(const float* x, const float* scalar_dev, int64_t n, float* out, map_stream_t stream);
/* dst[m, 0..N) = src[m, 0..N) with independent row strides (padding a [N,K] weight to an aligned leading dimension) */
int map_copy2d_f32(const float* src, int64_t ld_src, int64_t M, int N, float* dst, int64_t ld_dst, map_stream_t stream);
/* out[i, :] = X[idx[i], :] for an int64 [n_rows, F] matrix — device-resident batcher replacing DataLoader collate +
 * OurDataset.__getitem__ (code/trainer.py:51-58, code/dataset.py:78-87) */
int map_gather_rows_i64(const int64_t* X, int64_t n_rows, int F, const int64_t* idx, int64_t n, int64_t* out,
                        map_stream_t stream);
/* out[N,M] = in[M,N]^T */
int map_transpose_f32(const float* in, int64_t ld_in, int64_t M, int64_t N, float* out, int64_t ld_out, map_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAP_B200_H */
