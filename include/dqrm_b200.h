/*
 * dqrm_b200.h -- C ABI of libdqrm_b200.so: the B200 (sm_100a) implementation of
 * DQRM's data-parallel hot path.
 *
 * The reference (YangZhou08/Deep_Quantized_Recommendation_Model_DQRM) is pure
 * Python over stock PyTorch ops: it has no FFI of its own, so each entry point
 * below names the reference Python function (file:line under the reference
 * tree) whose arithmetic it replaces.  The Python surface that binds these
 * (ctypes) keeps the reference's module / function names.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in any signature.  `stream` is
 *     a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - "dev" pointers are device memory owned by the caller; "host" pointers are
 *     small per-table metadata arrays read synchronously during the call (they
 *     are baked into the kernel parameters, so calls are CUDA-graph capturable).
 *   - every call is asynchronous on `stream`, never synchronises the device,
 *     and never allocates.  Workspaces are sized by the *_bytes queries.
 *   - return value: 0 = launched; <0 = -errno style (-EINVAL bad shape / bits /
 *     alignment, -E2BIG more than DQRM_MAX_TABLES tables or a table too large
 *     for the fast path, -ENOMEM workspace too small, -EIO CUDA launch error).
 *     dqrm_last_error() returns a thread-local description of the last failure.
 *   - data-dependent errors (index out of range) cannot be reported
 *     synchronously: kernels OR a DQRM_STATUS_* bit into the caller's `status`
 *     word (dev int32, zero it once) and clamp the offending index.
 *   - all floating point is IEEE fp32 with round-to-nearest-even, no FMA
 *     contraction on the quantisation / update paths (bit-exact codes and
 *     scales versus the reference); tables are fp32 [rows, dim] row-major with
 *     16-byte aligned base and dim % 4 == 0.
 *   - rows of a table are addressed by int64 on input (the reference's index
 *     dtype) and stored as int32 in de-duplicated row lists (rows < 2^31).
 */
#ifndef DQRM_B200_H
#define DQRM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DQRM_API __attribute__((visibility("default")))
#else
#define DQRM_API
#endif

/* `lr_dev` (dqrm_sgd_rows, dqrm_grad_merge_apply, dqrm_dense_apply, dqrm_dense_apply_gathered): when not NULL the
 * learning rate is read from this device fp32 scalar at kernel run time instead of the by-value `lr`, so a captured
 * CUDA graph follows an LR schedule (LRPolicyScheduler, dlrm_s_pytorch_comm_grad.py:221-255) without re-capture. */
#define DQRM_ABI_VERSION 5
#define DQRM_MAX_TABLES 64           /* tables per call (kernel-parameter descriptor size) */
#define DQRM_BWD_CTA_MAX_LOOKUPS 16384 /* per-table lookups handled by the single-CTA sort path */
#define DQRM_FOLD_BLOCK 64            /* duplicate-row gradients: left fold in lookup order; rows with more duplicates
                                         are folded in blocks of this many lookups, then the block sums left to right */

#define DQRM_STATUS_INDEX_RANGE 1    /* an index was <0 or >= rows (clamped) */
#define DQRM_STATUS_OFFSET_ORDER 2   /* offsets not monotone / outside the index segment */
#define DQRM_STATUS_CAPACITY 4       /* more unique rows than `capacity` */
#define DQRM_STATUS_P2P_TIMEOUT 8    /* a peer never signalled an exchange site within DQRM_P2P_TIMEOUT_S (default 30 s, 0 = wait
                                        for ever).  FATAL and sticky: while set, dqrm_grad_merge_apply and
                                        dqrm_dense_apply_gathered are no-ops (stale slots are never applied) */

DQRM_API int dqrm_abi_version(void);
DQRM_API const char* dqrm_last_error(void);

/* ------------------------------------------------------------------ (a1) --
 * Per-table symmetric scale: absmax = max|W| over the whole table, then
 * s = max(absmax, 1e-8) / (2^(bits-1)-1) and inv_s = 1.0f / s.
 * Replaces symmetric_linear_quantization_param_two
 *   (quantization_supp/quant_utils.py:141-194), called once per table per
 *   forward by QuantEmbeddingBagTwo.forward
 *   (quantization_supp/quant_modules_not_quantize_grad.py:337).
 * All tables are reduced by ONE launch.  `shard_world` > 1 restricts the scan
 * to the `shard_rank`-th contiguous slice of every table's rows (replicas are
 * identical, so max over ranks == full scan): then only `absmax` is meaningful
 * and the caller combines ranks with a MAX all-reduce before
 * dqrm_scale_from_absmax.
 *   weight   host array [num_tables] of dev pointers (fp32 [rows_k, dim])
 *   rows     host array [num_tables]
 *   absmax   dev [num_tables] out
 *   scale, inv_scale  dev [num_tables] out, or both NULL (absmax only)
 *   workspace dev, dqrm_scan_workspace_bytes(); must be zero before first use,
 *            every launch leaves it zero again.
 */
DQRM_API size_t dqrm_scan_workspace_bytes(int num_tables);
DQRM_API int dqrm_table_absmax_scale(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                            int bits, int shard_rank, int shard_world,
                            float* absmax, float* scale, float* inv_scale,
                            void* workspace, void* stream);
/* s = max(absmax,1e-8)/n ; inv = 1/s  for n_scales independent entries (quant_utils.py:189-192). */
DQRM_API int dqrm_scale_from_absmax(int n_scales, const float* absmax, int bits, float* scale, float* inv_scale,
                           void* stream);

/* Exact incremental form of (a1) -- SURVEY.md section 8 (f-1).  max|W| is maintained per block of
 * `block_rows` rows; after an update only the blocks holding an updated row are recomputed (from the table,
 * so decreases are exact), and the per-table scale is the reduction of the block maxima: run
 * dqrm_table_absmax_scale over the block-max arrays (each viewed as a [entries, 1] fp32 table).  The result
 * is bit-identical to a full rescan.  The caller must route every table mutation through
 * dqrm_grad_merge_apply / dqrm_sgd_rows + dqrm_blockmax_update, or rebuild.
 *   blockmax   host array [num_tables] of dev fp32 [dqrm_blockmax_entries(rows_k, block_rows)] (16-byte aligned)
 *   update     from the gathered exchange slots (gathered != NULL: world, capacity, bits as in
 *              dqrm_grad_merge_apply) or from a local row list (uniq_rows / uniq_count, gathered == NULL)
 */
DQRM_API int64_t dqrm_blockmax_entries(int64_t rows, int block_rows);
DQRM_API int dqrm_blockmax_build(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                 int block_rows, float* const* blockmax, void* stream);
DQRM_API int dqrm_blockmax_update(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                  int block_rows, float* const* blockmax,
                                  const void* gathered, int world, int64_t capacity, int bits,
                                  const int32_t* uniq_rows, const int32_t* uniq_count, void* stream);

/* Pipelined form of the period-1 rescan (a1).  The reference serialises a full min/max pass over every table in
 * front of every forward (quant_modules_not_quantize_grad.py:337 -> quant_utils.py:177-178).  The tables are
 * read-only between two updates, so the pass that yields the scale of step i+1 can run concurrently with step i:
 *   dqrm_blockmax_scan          (any stream, typically a low-priority one) reads this rank's shard of every table
 *                               once and writes one max|w| per block of `block_rows` rows;
 *   dqrm_blockmax_update_shard  (after the row update) recomputes the blocks that hold an updated row;
 *   dqrm_blockmax_reduce        block maxima -> absmax[T] (and scale / 1/scale when scale != NULL; with
 *                               shard_world > 1 pass scale == NULL, MAX-all-reduce absmax, then
 *                               dqrm_scale_from_absmax).
 * Every table byte is still read once per step, nothing is carried from one step to the next, and the scale is
 * bit-identical to dqrm_table_absmax_scale (max is exact and order-free).  Shards are balanced contiguous block
 * ranges (the get_my_slice rule, dlrm_s_pytorch_comm_grad.py:993-997, applied to blocks); the three calls must
 * use the same (shard_rank, shard_world).  dim must be a multiple of 4; workspace as dqrm_scan_workspace_bytes. */
DQRM_API int dqrm_blockmax_scan(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                int block_rows, float* const* blockmax, int shard_rank, int shard_world,
                                void* stream);
DQRM_API int dqrm_blockmax_update_shard(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                        int block_rows, float* const* blockmax,
                                        const void* gathered, int world, int64_t capacity, int bits,
                                        const int32_t* uniq_rows, const int32_t* uniq_count,
                                        int shard_rank, int shard_world, void* stream);
DQRM_API int dqrm_blockmax_reduce(int num_tables, const int64_t* rows, int block_rows, const float* const* blockmax,
                                  int shard_rank, int shard_world, int bits, float* absmax, float* scale,
                                  float* inv_scale, void* workspace, void* stream);

/* ------------------------------------------------------------------ (a3) --
 * Fused gather + sum-pool + fake-quantise + dequantise for all tables in one
 * launch.  Replaces QuantEmbeddingBagTwo.forward steps (ii)-(iv)
 *   (quant_modules_not_quantize_grad.py:367,378,393) with SymmetricQuantFunction
 *   (quant_utils.py:322-346, linear_quantize :75-101), as called per table by
 *   DLRM_Net.apply_emb (dlrm_s_pytorch_comm_grad.py:614-679).
 *   pooled = left fold of W[idx] over the bag, in index order
 *   q   = clamp(rint(inv_s * pooled), -2^(bits-1), 2^(bits-1)-1)
 *   out = q * s                       (scale == NULL: out = pooled, qm:395)
 *   indices  dev int64, all tables' lookups concatenated
 *   idx_begin host [num_tables+1]: table k owns indices[idx_begin[k] .. idx_begin[k+1])
 *   offsets  dev int64 [num_tables, bags]: bag starts RELATIVE to the table's
 *            segment (nn.EmbeddingBag offsets, no trailing end)
 *   out      dev fp32; element (k, b, d) at out[k*out_table_stride + b*out_bag_stride + d]
 *   codes    NULL, or dev [num_tables, bags, dim] int8 (bits<=8) / int16 (bits<=16)
 */
DQRM_API int dqrm_embbag_fwd(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                    const int64_t* indices, const int64_t* offsets, const int64_t* idx_begin, int64_t bags,
                    const float* scale, const float* inv_scale, int bits,
                    float* out, int64_t out_table_stride, int64_t out_bag_stride,
                    void* codes, int32_t* status, void* stream);

/* Bit-packed INT4 tables and the forward that reads them (north_star kernel 2; SURVEY.md section 8 f-4:
 * the packed checkpoint / serving format next to this path; the reference's own inference path uses ATen
 * quantized.embedding_bag_4bit_*, dlrm_s_pytorch.py:428-469).
 *   dqrm_table_pack_int4 : code = clamp(rint(inv_scale_k * w), -8, 7) (quant_utils.py:101,343); element d of a
 *                          row is stored in byte d/2, low nibble for even d.  packed[k] is dev [rows_k, dim/2].
 *   dqrm_embbag_fwd_int4 : out[b] = scale_k * sum_{l in bag b} code(row_l): exact integer pooling, one multiply.
 *                          Bit-identical to dqrm_embbag_fwd when every bag has one index (Criteo); for longer
 *                          bags it is quantise-then-pool (serving), not the training pool-then-quantise.
 * dim must be a multiple of 16 with dim/16 a power of two.
 */
DQRM_API int dqrm_table_pack_int4(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                  const float* inv_scale, uint8_t* const* packed, void* stream);
DQRM_API int dqrm_embbag_fwd_int4(int num_tables, const uint8_t* const* packed, const int64_t* rows, int dim,
                                  const int64_t* indices, const int64_t* offsets, const int64_t* idx_begin,
                                  int64_t bags, const float* scale, float* out, int64_t out_table_stride,
                                  int64_t out_bag_stride, int32_t* status, void* stream);

/* ------------------------------------------------- (a4 bwd, a5, a7 step 1-2) --
 * Sparse row gradient of (a3), de-duplicated: dy = (g*s)/s (autograd of
 * qm:393 then SymmetricQuantFunction.backward, quant_utils.py:348-363; no clip
 * mask), one value row per lookup (ATen sparse EmbeddingBag backward,
 * triggered at dlrm_s_pytorch_comm_grad.py:1938), then Tensor.coalesce
 * (sgd_quantized_gradients_parallel_comm.py:859): rows sorted ascending and
 * unique, duplicates summed as a left fold in original lookup order (rows with more than DQRM_FOLD_BLOCK
 * duplicates: blocks of DQRM_FOLD_BLOCK consecutive lookups folded left to right, then the block sums left to
 * right -- one fixed order, defined by the oracle's coalesce_spec; the reference's own fold order is
 * implementation-defined).
 * Optionally also the per-table gradient scale of quantize_emb_grad step 2
 * (sgd...parallel_comm.py:861): s_local = max(max|sums|,1e-8)/(2^(grad_bits-1)-1).
 *   dout       dev fp32, element (k,b,d) at dout[k*dout_table_stride + b*dout_bag_stride + d]
 *   fwd_scale  dev [num_tables] (the forward's s) or NULL for the full-precision forward (dy = g)
 *   capacity   slots per table in the outputs (>= lookups of the largest table)
 *   uniq_rows  dev int32 [num_tables, capacity]; uniq_count dev int32 [num_tables]
 *   grad_sums  dev fp32 [num_tables, capacity, dim]
 *   grad_scale_local dev [num_tables] or NULL (then grad_bits is ignored)
 * One CTA per table sorts (row, bag) keys in shared memory; tables with more
 * than DQRM_BWD_CTA_MAX_LOOKUPS lookups take the multi-block radix-sort path.
 * `workspace` (16-byte aligned, dqrm_bwd_workspace_bytes(num_tables, capacity, dim) bytes) holds the block sums
 * of the long rows, or the sort buffers of the radix path.
 */
DQRM_API size_t dqrm_bwd_workspace_bytes(int num_tables, int64_t max_lookups_per_table, int dim);
DQRM_API int dqrm_embbag_bwd(int num_tables, const int64_t* rows, int dim,
                    const int64_t* indices, const int64_t* offsets, const int64_t* idx_begin, int64_t bags,
                    const float* dout, int64_t dout_table_stride, int64_t dout_bag_stride,
                    const float* fwd_scale,
                    int64_t capacity, int32_t* uniq_rows, int32_t* uniq_count, float* grad_sums,
                    int grad_bits, float* grad_scale_local,
                    int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* Stand-alone (a7 step 2): s_local[k] from grad_sums / uniq_count. */
DQRM_API int dqrm_grad_absmax_scale(int num_tables, int dim, const float* grad_sums, const int32_t* uniq_count,
                           int64_t capacity, int bits, float* scale_local, void* stream);

/* ----------------------------------------------------------------- (a10) --
 * Un-quantised row update  W[row] += (-lr) * (sum * inv_world)
 * (weight_update_parallel_comm with emb_grad_quantized=False,
 *  sgd...parallel_comm.py:626; single process: torch.optim.SGD on the sparse grad).
 * With `momentum` != NULL applies row-wise sparse Adagrad instead
 * (optim/rwsadagrad.py:97-113): m[row] += mean(g^2); W[row] -= lr * g / (sqrt(m[row]) + eps).
 *   momentum  NULL, or host array [num_tables] of dev fp32 [rows_k] accumulators
 */
DQRM_API int dqrm_sgd_rows(int num_tables, float* const* weight, const int64_t* rows, int dim,
                  const int32_t* uniq_rows, const int32_t* uniq_count, const float* grad_sums, int64_t capacity,
                  float lr, const float* lr_dev, float inv_world, float* const* momentum, float eps, void* stream);

/* -------------------------------------------------------- (a5 + a10, fused) --
 * De-duplicating backward WITH the row update applied in place (the single-process path: ATen sparse EmbeddingBag
 * backward, dlrm_s_pytorch_comm_grad.py:1938, then torch.optim.SGD.step() on the sparse gradient,
 * dlrm_s_pytorch_single_gpu.py:1944-1946 / W.add_(-lr*grad), sgd...parallel_comm.py:626; with `momentum` the
 * row-wise sparse Adagrad of optim/rwsadagrad.py:97-113).  Same arguments as dqrm_embbag_bwd + dqrm_sgd_rows and the
 * same table bits as calling the two in turn: duplicates are summed first (the fixed fold order above), then
 * W[row] += (-lr) * (sum * inv_world).  On the radix-sort path the update runs inside the fold of the same kernel --
 * the table row is fetched beside the dOut gathers and the sums never travel through memory; tables with few lookups
 * take the single-CTA de-duplication followed by the row-update kernel.
 *   uniq_rows / uniq_count   out, as in dqrm_embbag_bwd (the rows that changed: scale tracker, INT4 shadow)
 *   grad_sums                scratch [num_tables, capacity, dim]; contents undefined afterwards
 */
DQRM_API int dqrm_embbag_bwd_sgd(int num_tables, float* const* weight, const int64_t* rows, int dim,
                    const int64_t* indices, const int64_t* offsets, const int64_t* idx_begin, int64_t bags,
                    const float* dout, int64_t dout_table_stride, int64_t dout_bag_stride,
                    const float* fwd_scale,
                    int64_t capacity, int32_t* uniq_rows, int32_t* uniq_count, float* grad_sums,
                    float lr, const float* lr_dev, float inv_world, float* const* momentum, float eps,
                    int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* ----------------------------------------------------------- (a7 steps 3-4) --
 * Quantise this rank's de-duplicated row gradients into its exchange slot.
 *   s_bar[k] = (sum_r gathered_scales[r][k]) * (1/world), summed in rank order
 *              (all_reduce(SUM) then mul_(1./N), sgd...parallel_comm.py:865-866)
 *   q        = clamp(rint((1/s_bar) * sums), -2^(bits-1), 2^(bits-1)-1)   (:869)
 * bits == 32 selects the UN-quantised exchange (emb_grad_quantized=False: coalesce, all-reduce, 1/N, W += -lr*g,
 * sgd...parallel_comm.py:319-329, 626): the slot carries the fp32 sums, gathered_scales is ignored, and
 * dqrm_grad_merge_apply sums them in rank order.
 * Slot layout (dqrm_slot_bytes; offsets from dqrm_slot_layout), fixed capacity
 * so one all-gather of `slot_bytes` per rank replaces the reference's Gloo
 * sparse all-reduce (:878):
 *   int32 count[num_tables] | int32 rows[num_tables][capacity] |
 *   int8 (bits<=8), int16 (bits<=16) or fp32 (bits==32) codes[num_tables][capacity][dim]
 *   gathered_scales dev [world] rows of num_tables floats, row r at r * scale_stride_elems (= num_tables for a
 *   dense [world, num_tables] array; the padded slot stride of a peer-arena site);  scale_mean dev [num_tables] out
 */
DQRM_API size_t dqrm_slot_bytes(int num_tables, int64_t capacity, int dim, int bits);
DQRM_API int dqrm_slot_layout(int num_tables, int64_t capacity, int dim, int bits,
                     size_t* rows_offset, size_t* codes_offset);
DQRM_API int dqrm_grad_pack(int num_tables, int dim, const float* grad_sums, const int32_t* uniq_rows,
                   const int32_t* uniq_count, int64_t capacity,
                   const float* gathered_scales, int64_t scale_stride_elems, int world, int bits,
                   void* slot, float* scale_mean, void* stream);

/* ------------------------------------------------------------------ (a8) --
 * Top-k row sparsification of the de-duplicated gradients, in place, between
 * coalesce and the gradient scale: per table keep the `topk` rows of largest
 * score ||g_row||^2 / dim (ties: lower row id); survivors stay in ascending row
 * order, uniq_count becomes min(count, topk), dropped rows are discarded.
 * north_star extension: the DQRM path itself only has "specified" sparsity
 * (rows touched by the batch); the only top-k in the reference tree is
 * training_imagenet_speedup.py:138,149 whose score/selection this follows.
 * Parity unpinned; topk >= count is the identity =
 * reference behaviour.  Recompute the scale afterwards with
 * dqrm_grad_absmax_scale.  Tables with more than DQRM_BWD_CTA_MAX_LOOKUPS
 * unique rows are rejected (-E2BIG).
 */
DQRM_API int dqrm_grad_topk(int num_tables, int dim, float* grad_sums, int32_t* uniq_rows, int32_t* uniq_count,
                   int64_t capacity, int64_t topk, void* stream);

/* ---------------------------------------------------- (a7 step 5, a9) --
 * Merge the `world` gathered slots and apply the SGD row update in place:
 *   for every row in the union of the ranks' row lists
 *     W[row] += (-lr) * ((float(sum_r q_r[row]) * (1/world)) * s_bar)
 * (sparse all-reduce + mul_(1./N), sgd...parallel_comm.py:878,885; update
 *  weight_update_parallel_comm :618,622).  Integer codes of coinciding rows are
 * summed exactly; each row is updated once.  Every rank runs this on the same
 * gathered bytes, so replicas stay bit-identical.
 *   gathered   dev, `world` slots back to back (all-gather output)
 *   updated_rows  NULL or dev int32 [num_tables, world*capacity]: union rows
 *                 (unordered); updated_count dev int32 [num_tables] (zeroed by the call)
 *   qbar       NULL or dev fp32 [num_tables, world*capacity, dim]: (sum q)*(1/world)
 *              aligned with updated_rows (the reference's .grad values after :885)
 */
DQRM_API int dqrm_grad_merge_apply(int num_tables, float* const* weight, const int64_t* rows, int dim,
                          const void* gathered, int world, int64_t capacity, int bits,
                          const float* scale_mean, float lr, const float* lr_dev,
                          int32_t* updated_rows, int32_t* updated_count, float* qbar,
                          int32_t* status, void* stream);

/* ----------------------------------------------------------------- (a14) --
 * Dot interaction: T = [x | ly_0 .. ly_{F-1}] per sample ([F+1, dim]),
 * Z = T T^t, R = [x | strict lower triangle of Z] (with `itself`, the diagonal too).
 * Replaces DLRM_Net.interact_features (dlrm_s_pytorch_comm_grad.py:701-725).
 *   ly   dev fp32, element (k,b,d) at ly[k*table_stride + b*bag_stride + d]
 *   R    dev fp32 [batch, dim + npairs]
 * Backward: dx, dly from dR (autograd of the same expression).  `ste_scale` (NULL or dev [num_tables]) fuses
 * the QAT EmbeddingBag's straight-through estimator into the epilogue: dly <- (dly * s_k) / s_k (autograd of
 * qm:393 then quant_utils.py:363); dqrm_embbag_bwd is then called with fwd_scale = NULL.
 */
DQRM_API int dqrm_interact_fwd(const float* x, const float* ly, int64_t ly_table_stride, int64_t ly_bag_stride,
                      int64_t batch, int num_tables, int dim, int itself, float* R, void* stream);
DQRM_API int dqrm_interact_bwd(const float* x, const float* ly, int64_t ly_table_stride, int64_t ly_bag_stride,
                      const float* dR, int64_t batch, int num_tables, int dim, int itself,
                      float* dx, float* dly, int64_t dly_table_stride, int64_t dly_bag_stride,
                      const float* ste_scale, void* stream);

/* --------------------------------------------------- (a3, north-star kernel 2) --
 * Packed-INT4 shadow rows in the TRAINING forward.  The reference quantises the POOLED vector
 * (quant_modules_not_quantize_grad.py:367,378,393); for a bag of one index Q(pooled) == Q(row), so the forward may read
 * the row's 4-bit codes (D/2 bytes) instead of its fp32 values (4D bytes) and return the same bits, as long as the
 * codes were made with the scale this forward uses.  shadow[k] is dev uint8 [rows_k, dim/2] (element d in byte d/2,
 * low nibble for even d; 16-byte aligned), shadow_scale dev fp32 [num_tables] = the scale each shadow was encoded
 * with (0 = never).  fp32 rows stay authoritative.
 *   dqrm_shadow_refresh     : tables whose scale differs BIT-WISE from shadow_scale are re-encoded from the fp32 rows
 *                             (flags: dev int32 [num_tables] scratch), shadow_scale <- scale.  Call after the scan.
 *   dqrm_shadow_update_rows : re-encode the rows a step updated (same row lists as dqrm_blockmax_update: the gathered
 *                             exchange slots, or uniq_rows/uniq_count) with the step's inv_scale.  Call after the update.
 *   dqrm_embbag_fwd_shadow  : dqrm_embbag_fwd (4-bit, int8 codes) that takes a bag's codes from the shadow when the bag
 *                             has one index and shadow_scale[k] == scale[k] bit-wise, and pools the fp32 rows otherwise
 *                             -- bit-identical outputs and codes either way.  Packed tables of <= 32 KiB are staged in
 *                             shared memory with 128-bit loads first.
 */
DQRM_API int dqrm_shadow_refresh(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                 const float* scale, const float* inv_scale, uint8_t* const* shadow,
                                 float* shadow_scale, int32_t* flags, void* stream);
DQRM_API int dqrm_shadow_update_rows(int num_tables, const float* const* weight, const int64_t* rows, int dim,
                                     const float* inv_scale, uint8_t* const* shadow, const void* gathered, int world,
                                     int64_t capacity, int bits, const int32_t* uniq_rows, const int32_t* uniq_count,
                                     void* stream);
DQRM_API int dqrm_embbag_fwd_shadow(int num_tables, const float* const* weight, uint8_t* const* shadow,
                                    const int64_t* rows, int dim, const int64_t* indices, const int64_t* offsets,
                                    const int64_t* idx_begin, int64_t bags, const float* scale, const float* inv_scale,
                                    const float* shadow_scale, float* out, int64_t out_table_stride,
                                    int64_t out_bag_stride, void* codes, int32_t* status, void* stream);

/* ----------------------------------------------------------------- (a15) --
 * QuantLinear weight/bias fake-quantisation, per output channel:
 *   s_row = max(max|W_row|,1e-8)/(2^(bits-1)-1); W_int = clamp(rint((1/s_row) W));
 *   b_int = clamp(rint((1/s_row) b))  -- bias uses the WEIGHT's scale and bits
 * (quant_modules_not_quantize_grad.py:125-154, quant_utils.py:196-220).
 */
DQRM_API int dqrm_linear_fakequant(const float* W, const float* b, int out_features, int in_features, int bits,
                          float* W_int, float* b_int, float* scale_row, void* stream);

/* Fused QuantLinear layers (same arithmetic as dqrm_linear_fakequant + F.linear + the activation
 * module that follows each layer in DLRM_Net.create_mlp, dlrm_s_pytorch_comm_grad.py:279-325):
 *   dqrm_mlp_fakequant_all : dqrm_linear_fakequant for every layer of the model in ONE launch; all array
 *                            arguments are host arrays [num_layers] (of dev pointers / sizes).
 *   dqrm_linear_fwd        : out = act((x W_int^t + b_int) * s_row)        act: 0 none, 1 relu, 2 sigmoid
 *   dqrm_linear_bwd        : g = dout * act'(out) * s_row ; dx = g W_int (dx may be NULL; dW/db may be NULL
 *                            so the two products can be issued on different streams) ;
 *                            dW (+)= (g^t x) / s_row ; db (+)= (sum_batch g) / s_row
 *                            (straight-through estimator, quant_utils.py:348-363); `accumulate` = 0 overwrites
 *                            (gradients freshly cleared, clear_gradients sgd:714), 1 adds to what is there.
 * `path` selects the contraction engine: DQRM_LINEAR_FFMA = fp32 FFMA (cluster split-K over distributed shared
 * memory), DQRM_LINEAR_TC = tcgen05 tensor cores with the fp32 operands split into TF32 terms inside the kernel
 * (W_int is exact in TF32; 2 MMAs per K-step for fwd / dx, 3 for dW; fp32 accumulation in TMEM; results agree with
 * the FFMA path to fp32 rounding), DQRM_LINEAR_AUTO = tensor cores from 256 batch rows.  Either way the summation
 * order is fixed (no atomics), so data-parallel replicas stay bit-identical.
 */
#define DQRM_LINEAR_AUTO 0
#define DQRM_LINEAR_FFMA 1
#define DQRM_LINEAR_TC 2
#define DQRM_LINEAR_FFMA_SERIAL 3 /* forward: the FFMA kernel's K-slices walked by ONE CTA (no cluster launch), same bits */
DQRM_API int dqrm_mlp_fakequant_all(int num_layers, const float* const* W, const float* const* b,
                                    const int32_t* out_features, const int32_t* in_features, int bits,
                                    float* const* W_int, float* const* b_int, float* const* scale_row, void* stream);
DQRM_API int dqrm_linear_fwd(const float* x, const float* W_int, const float* b_int, const float* scale_row,
                             int batch, int out_features, int in_features, int act, float* out, int path, void* stream);
DQRM_API int dqrm_linear_bwd(const float* x, const float* W_int, const float* scale_row, const float* dout,
                             const float* out, int batch, int out_features, int in_features, int act,
                             float* dx, float* dW, float* db, int accumulate, int path, void* stream);

/* ------------------------------------------------------------------ (a4) --
 * Stand-alone SymmetricQuantFunction.forward on a [rows, cols] fp32 matrix
 * (quant_utils.py:322-346): q = clamp(rint((1/s) * x), -2^(bits-1), 2^(bits-1)-1),
 * integer-valued fp32.  scale is dev [1] (scale_per_row = 0) or dev [rows]
 * (scale.view(-1,1), quant_utils.py:90-93).  `dequant` != NULL also writes q * s.
 */
DQRM_API int dqrm_fake_quant(const float* x, int64_t rows, int64_t cols, const float* scale, int scale_per_row,
                             int bits, float* q, float* dequant, void* stream);

/* ----------------------------------------------------------------- (a11) --
 * 8-bit per-channel quantised exchange of dense (MLP) gradients laid out in one
 * flat fp32 arena split into `num_chan` channels [chan_begin[c], chan_begin[c+1])
 * (a weight row is a channel; a whole bias vector is one channel):
 *   dqrm_dense_grad_scale : s_local[c] = max(max|g_c|,1e-8)/(2^(bits-1)-1)
 *        (quantize_linear_grad / quantize_bias_grad, sgd...parallel_comm.py:905-910, 945-947)
 *   dqrm_dense_grad_quant : s_bar = scale_sum * inv_world ; q = clamp(rint((1/s_bar) g))   (:912-919)
 *   dqrm_dense_apply      : p += ((-lr) * (code_sum * inv_world)) * s_bar
 *        (weight_update_parallel_comm :642-643 -- note the association differs from (a9))
 * chan_begin is dev int64 [num_chan+1]; codes are integer-valued fp32 so one
 * SUM all-reduce between quant and apply adds them exactly.
 * Error compensation (quantize_linear_grad / quantize_bias_grad with err_compensation=True,
 * sgd...parallel_comm.py:899-900,926-927,938-939,958-959; buffers error_compensation_weight/_bias,
 * quant_modules_not_quantize_grad.py:87,95): pass `error_comp` (same layout as grad) to dqrm_dense_grad_scale and
 * grad becomes grad + error_comp IN PLACE before the scale is taken; pass that compensated gradient and
 * `error_comp_out` to dqrm_dense_apply(_gathered) and error_comp_out = comp_grad - (code_sum * inv_world) * s_bar.
 * NULL = the reference's default (err_compensation=False).
 */
DQRM_API int dqrm_dense_grad_scale(float* grad, const float* error_comp, const int64_t* chan_begin, int num_chan, int bits,
                          float* scale_local, void* stream);
DQRM_API int dqrm_dense_grad_quant(const float* grad, const int64_t* chan_begin, int num_chan,
                          const float* scale_sum, float inv_world, int bits,
                          float* codes, float* scale_mean, void* stream);
DQRM_API int dqrm_dense_apply(float* param, const float* code_sum, const int64_t* chan_begin, int num_chan,
                     const float* scale_mean, float inv_world, float lr, const float* lr_dev, const float* comp_grad,
                     float* error_comp_out, void* stream);

/* Single rank (world 1): dqrm_dense_grad_scale + dqrm_dense_grad_quant (inv_world 1) + dqrm_dense_apply in ONE launch --
 * the same operations in the same order on the same buffers (scale_local, scale_mean, codes, grad (+= error_comp in
 * place), param, error_comp <- comp_grad - code * s_bar), so every result is bit-identical to the three calls; it
 * removes two launches from the tail of the step.  error_comp may be NULL.
 * Replaces quantize_linear_grad/quantize_bias_grad + the MLP half of weight_update_parallel_comm
 * (sgd_quantized_gradients_parallel_comm.py:892-961, 630-663) when there is nothing to exchange. */
DQRM_API int dqrm_dense_quant_apply_local(float* param, float* grad, float* error_comp, const int64_t* chan_begin,
                                 int num_chan, int bits, float* scale_local, float* codes, float* scale_mean,
                                 float lr, const float* lr_dev, void* stream);

/* BCE loss (mean reduction) and its gradient in one launch: torch.nn.BCELoss()(Z, T) + E.backward() of the
 * reference loop (loss_fn_wrap, dlrm_s_pytorch_comm_grad.py:192-211; :1938).
 *   *loss = mean((t-1)*max(log1p(-z),-100) - t*max(log z,-100));  dz = ((z-t) / max((1-z)*z, 1e-12)) * (1/n)
 * z, target, dz: dev fp32 [n] (dz may be NULL); loss: dev fp32 scalar.  Deterministic summation order. */
DQRM_API int dqrm_bce_loss_grad(const float* z, const float* target, int64_t n, float* loss, float* dz, void* stream);

/* ------------------------------------------- exchange over NVLink peer memory --
 * One-kernel all-gathers that replace the step's collectives when world > 1 (the reference: Gloo all-reduces
 * per table / per tensor, sgd_quantized_gradients_parallel_comm.py:865,878,913,925,949,957; this library's
 * portable form: 2 NCCL all-gathers + 3 all-reduces).  Every rank allocates one PEER ARENA and maps the arenas
 * of the other ranks of the box (CUDA IPC; the 64-byte handles travel through the caller's own channel, e.g.
 * torch.distributed.all_gather).  A SITE is a region at the same offset in every arena:
 *     ctl[16 B] | flag[world] u32 (padded to 16 B) | slot[world][slot_stride]      (dqrm_p2p_site_layout)
 * The producer kernel writes this rank's contribution into slot[rank] of its own arena; dqrm_p2p_allgather
 * copies it into slot[rank] of every peer with remote stores, signals, and waits for every peer's signal, so
 * that when it retires slot[0..world) of the LOCAL arena is complete.  The consumers below reduce the slots in
 * rank order (deterministic, identical on every rank).  Sequence numbers live on the device: the call is
 * CUDA-graph replayable.  Requirements: every rank calls the same sites in the same order on one stream, at
 * least two distinct sites per step (single buffering, see csrc/p2p.cu), arenas zero-initialised (p2p_alloc
 * does) and all ranks past a barrier after opening the handles.  A peer that never arrives makes the wait give
 * up after ~2 s with DQRM_STATUS_P2P_TIMEOUT in *status. */
DQRM_API int dqrm_p2p_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_64);
DQRM_API int dqrm_p2p_open(const void* ipc_handle_64, void** dev_ptr);
DQRM_API int dqrm_p2p_close(void* dev_ptr);
DQRM_API int dqrm_p2p_free(void* dev_ptr);
DQRM_API size_t dqrm_p2p_site_bytes(int world, size_t slot_bytes);
DQRM_API int dqrm_p2p_site_layout(int world, size_t slot_bytes, size_t* flag_off, size_t* data_off, size_t* slot_stride);
/* peer_base: host array [world] of the mapped arena bases (entry `rank` = the local arena) */
DQRM_API int dqrm_p2p_allgather(void* const* peer_base, int world, int rank, size_t site_off, size_t slot_bytes,
                                int32_t* status, void* stream);
/* (a11) quantize_linear_grad / quantize_bias_grad for all tensors, scales summed in rank order from the gathered
 * per-channel local scales (row r at r * scale_stride_elems), codes as int8 (bits <= 8) -- typically straight
 * into this rank's slot of the codes site. */
DQRM_API int dqrm_dense_grad_quant_gathered(const float* grad, const int64_t* chan_begin, int num_chan,
                                            const float* gathered_scales, size_t scale_stride_elems, int world,
                                            int bits, int8_t* codes, float* scale_mean, void* stream);
/* (a11 + a9) param += (-lr * ((sum_r codes_r) * (1/world))) * scale_mean  (sgd...parallel_comm.py:925,642-643,
 * 957,662-663); rank r's int8 codes at gathered_codes + r * code_stride_bytes. */
DQRM_API int dqrm_dense_apply_gathered(float* param, const int8_t* gathered_codes, size_t code_stride_bytes, int world,
                                       const int64_t* chan_begin, int num_chan, const float* scale_mean, float lr,
                                       const float* lr_dev,
                                       const float* comp_grad, float* error_comp_out, const int32_t* status,
                                       void* stream);
/* (a1, row-sharded scan) absmax = max over ranks of the gathered per-shard maxima, then scale and 1/scale. */
DQRM_API int dqrm_scale_from_absmax_gathered(int n_scales, const float* gathered_absmax, size_t stride_elems, int world,
                                             int bits, float* absmax, float* scale, float* inv_scale, void* stream);

/* (a11 + a9, MLP half) the WHOLE dense-gradient exchange of one step in one launch (csrc/dense_xchg.cu): replaces
 * dqrm_dense_grad_scale -> allgather -> dqrm_dense_grad_quant_gathered -> allgather -> dqrm_dense_apply_gathered
 * (quantize_linear_grad / quantize_bias_grad + the MLP part of weight_update_parallel_comm,
 * sgd_quantized_gradients_parallel_comm.py:892-961,630-663) with the same arithmetic in the same order, so the
 * parameters are bit-identical.  CTA b owns channels [cta_chan[b], cta_chan[b+1]) on every rank and exchanges only
 * with CTA b of the peers (no grid-wide step).  Every 8-byte word of a slot carries its payload and the step's sequence number and is
 * written with one 64-bit remote store / polled by the reader: no fence, no separate flag.  Sites (data areas of
 * dqrm_p2p_site_layout sites, byte offsets from the arena base, multiples of 16):
 *   scale slots  u64 [num_chan]                 payload = the rank's local fp32 scale of the channel
 *   code slots   u64 [cta_word[num_ctas]]       seven int8 codes + the low byte of the sequence number; CTA b's run
 *                                               starts at word cta_word[b]
 *   cta_chan, cta_word   dev int32 [num_ctas + 1], the same on every rank (cta_word[b+1] - cta_word[b] =
 *                  ceil(elements of run b / 7)); max_cta_elems / max_cta_chans bound one CTA's run
 *   seq            dev u32 [num_ctas], zero-initialised together with the arena; advanced by the kernel (replayable)
 *   error_comp     dev or NULL: grad += error_comp in place before the scales, residual written back (:899-900,926-927)
 * num_ctas <= 148: every CTA has to be resident, they wait for their peers.  Watchdog / status as dqrm_p2p_allgather;
 * a CTA that saw the timeout bit does not touch the parameters. */
DQRM_API size_t dqrm_dense_exchange_smem_bytes(int max_cta_elems, int max_cta_chans);
/* debug hook: dev u64 [num_ctas][8] (or NULL = off, the default): %globaltimer of thread 0 of every CTA at launch, after
 * the scale stores, after the scale wait, after the code stores, after the code wait and at the end -- where a launch spends its time */
DQRM_API int dqrm_dense_exchange_debug(uint64_t* stamps);
DQRM_API int dqrm_dense_exchange_apply(void* const* peer_base, int world, int rank, size_t scale_data_off,
                                       size_t scale_stride_bytes, size_t code_data_off, size_t code_stride_bytes,
                                       float* param, float* grad,
                                       float* error_comp, const int64_t* chan_begin, const int32_t* cta_chan,
                                       const int32_t* cta_word, int num_ctas, int max_cta_elems, int max_cta_chans, int bits, float* scale_mean,
                                       uint32_t* seq, float lr, const float* lr_dev, int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DQRM_B200_H */
