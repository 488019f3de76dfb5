/*
 * msda_b200.h -- C ABI of the B200-native multi-scale deformable attention library
 * (libmsda_b200.so, built from apollo-vision-net_b200/csrc for sm_100a).
 *
 * This is the drop-in boundary for the hot path of HankerSia/Apollo-Vision-Net:
 * every entry point below is what the reference's FFI for this path binds.  In the
 * reference the binding is mmcv's compiled `_ext` module, loaded at
 *   projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py:8-10
 * and called at :40-46 / :116-122 (forward) and :72-82 / :148-158 (backward).
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / ATen types.
 *  - every pointer is a DEVICE pointer unless its name ends in `_host`.
 *  - the caller owns and allocates every buffer (outputs and scratch included); the
 *    library never frees or retains a pointer past the call.
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls are
 *    asynchronous, stateless and re-entrant.
 *  - return value: 0 on success, a negative MSDA_ERR_* code on failure;
 *    msda_last_error() returns a thread-local message for the last failure
 *    (mirrors the AT_ASSERTM / AT_ERROR behaviour of the reference-era ops, cf.
 *    projects/mmdet3d_plugin/bevformer/backbones/ops_dcnv3/src/cuda/dcnv3_cuda.cu:28-53).
 *  - spatial_shapes / level_start_index arrive as DEVICE int64 arrays exactly as the
 *    reference passes them (transformer.py:251-254); they are read on the device,
 *    there is no host synchronisation anywhere in this library.
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 8: the row kernels' `partial` scratch holds eight reduction strips (64 + 8 * n * C floats); new entries
 *    grad_amax_scale_zero (scale + accumulator clear in one launch) and unscale_cast_strided (result written as a
 *    column block of a wider matrix).  Signatures of the older entries are unchanged since 7. */
#define MSDA_ABI_VERSION 8

/* element types of `value` / `out` (and optionally of locations / weights) */
#define MSDA_F32  0
#define MSDA_F16  1
#define MSDA_BF16 2

#define MSDA_OK                 0
#define MSDA_ERR_BAD_ARGUMENT  -1
#define MSDA_ERR_UNSUPPORTED   -2
#define MSDA_ERR_CUDA          -3

#define MSDA_MAX_LEVELS 16

int msda_abi_version(void);
const char* msda_last_error(void);

/* ---------------------------------------------------------------------------------
 * Operator boundary: replaces ext_module.ms_deform_attn_forward
 * (multi_scale_deformable_attn_function.py:40-46, :116-122).
 *
 *   value   (B, Nk, M, Dh)        value_dtype, contiguous
 *   shapes  (L, 2) int64 (h, w);  starts (L,) int64
 *   loc     (B, Nq, M, L, P, 2)   coord_dtype, normalised (x, y)
 *   attn    (B, Nq, M, L, P)      coord_dtype
 *   out     (B, Nq, M*Dh)         value_dtype, fully overwritten
 *
 * coord_dtype is MSDA_F32 or equal to value_dtype.  Accumulation is fp32.
 * im2col_step has no meaning for these kernels; it is validated like mmcv does
 * (B % min(B, im2col_step) == 0) and otherwise ignored.
 * ------------------------------------------------------------------------------- */
int msda_fwd(const void* value, const int64_t* shapes, const int64_t* starts,
             const void* loc, const void* attn, void* out,
             int B, int Nk, int M, int Dh, int L, int Nq, int P,
             int value_dtype, int coord_dtype, int im2col_step, void* stream);

/* ---------------------------------------------------------------------------------
 * Replaces ext_module.ms_deform_attn_backward
 * (multi_scale_deformable_attn_function.py:72-82, :148-158).
 *
 *   grad_out (B, Nq, M*Dh)       value_dtype, contiguous
 *   g_value  (B, Nk, M, Dh)      fp32 ACCUMULATOR: the kernel adds into it, the caller
 *                                zero-fills it first (the reference does the same with
 *                                torch.zeros_like, :144)
 *   g_loc    (B, Nq, M, L, P, 2) fp32, every element written exactly once
 *   g_attn   (B, Nq, M, L, P)    fp32, every element written exactly once
 * ------------------------------------------------------------------------------- */
int msda_bwd(const void* value, const int64_t* shapes, const int64_t* starts,
             const void* loc, const void* attn, const void* grad_out,
             float* g_value, float* g_loc, float* g_attn,
             int B, int Nk, int M, int Dh, int L, int Nq, int P,
             int value_dtype, int coord_dtype, int im2col_step, void* stream);

/* ---------------------------------------------------------------------------------
 * Host-buffer variants of the two calls above (the end-to-end leg of bench.py and
 * non-PyTorch callers): every tensor pointer is a HOST pointer (pinned memory gives
 * asynchronous copies), `scratch` is a DEVICE buffer of at least
 * msda_host_scratch_bytes(...) bytes.  Inputs are copied to the device, the kernel
 * runs, results are copied back, all on `stream`; the call returns after the stream
 * has been synchronised.
 * ------------------------------------------------------------------------------- */
int64_t msda_host_scratch_bytes(int B, int Nk, int M, int Dh, int L, int Nq, int P,
                                int value_dtype, int coord_dtype, int with_backward);
int msda_fwd_host(const void* value_host, const int64_t* shapes_host, const int64_t* starts_host,
                  const void* loc_host, const void* attn_host, void* out_host,
                  int B, int Nk, int M, int Dh, int L, int Nq, int P,
                  int value_dtype, int coord_dtype, void* scratch, int64_t scratch_bytes,
                  void* stream);
int msda_fwd_bwd_host(const void* value_host, const int64_t* shapes_host, const int64_t* starts_host,
                      const void* loc_host, const void* attn_host, const void* grad_out_host,
                      void* out_host, float* g_value_host, float* g_loc_host, float* g_attn_host,
                      int B, int Nk, int M, int Dh, int L, int Nq, int P,
                      int value_dtype, int coord_dtype, void* scratch, int64_t scratch_bytes,
                      void* stream);

/* ---------------------------------------------------------------------------------
 * BEV geometry: replaces BEVFormerEncoder.point_sampling (encoder.py:89-241) and the
 * per-camera nonzero() compaction of SpatialCrossAttention.forward
 * (spatial_cross_attention.py:135-139), with no host synchronisation.
 *
 *   ref_3d     (bs, D, HW, 3) fp32 pillar points in [0,1]^3 (get_reference_points,
 *              encoder.py:62-72), exactly the `reference_points` argument of the reference
 *   lidar2img  (bs, num_cam, 4, 4) fp32
 *   pc_range_host  6 doubles on the HOST (x0, y0, z0, x1, y1, z1)
 *   img_h/img_w    image size of camera 0 / sample 0 (the reference normalises every camera
 *              with it, encoder.py:196-226)
 *   ref_cam    (num_cam, bs, HW, D, 2) fp32   out
 *   bev_mask   (num_cam, bs, HW, D) uint8 (0/1) out  -- bit-exact w.r.t. the reference
 *   hit_bits   (bs, HW) uint32 out: bit i set <=> camera i sees query (any Z anchor)
 *   hit_index  (num_cam, HW) int32 out: ascending query indices seen by camera i in
 *              batch element 0 (the reference's quirk, :137); entries past the count
 *              are -1
 *   hit_count  (num_cam,) int32 out   (hit_index and hit_count may both be NULL: the fused
 *              kernels only read hit_bits, the ordered lists are for callers that want them)
 * ------------------------------------------------------------------------------- */
int bev_point_sampling(const float* ref_3d, const float* lidar2img, const double* pc_range_host,
                       float img_h, float img_w, int bs, int num_cam, int HW, int D,
                       float* ref_cam, uint8_t* bev_mask, uint32_t* hit_bits,
                       int32_t* hit_index, int32_t* hit_count, void* stream);

/* ---------------------------------------------------------------------------------
 * Fused spatial cross-attention core: everything between the Linear layers of
 * SpatialCrossAttention / MSDeformableAttention3D
 * (spatial_cross_attention.py:135-170 and :342-396): camera-hit gating, softmax over
 * L*P, Z-anchor sampling locations, multi-level bilinear sampling, the sum over
 * cameras and the division by the hit count -- without materialising
 * sampling_locations, the rebatched queries or the scatter.
 *
 *   value    (bs*num_cam, Nk, M, Dh)  value_dtype (output of value_proj)
 *   offsets  (bs, HW, M, L, P, 2)     coord_dtype raw output of sampling_offsets(query)
 *   logits   (bs, HW, M, L*P)         coord_dtype raw output of attention_weights(query)
 *            (coord_dtype = MSDA_F32 or value_dtype: a bf16 model feeds its Linear outputs as is)
 *   ref_cam, bev_mask, hit_bits       from bev_point_sampling (the kernels read the per-query
 *                                      camera bit field; bev_mask is accepted for the contract)
 *   bev_w    > 0: the HW queries are a row-major (HW / bev_w) x bev_w grid (2-D work tiles); 0: unknown
 *   slots    (bs, HW, M*Dh)           value_dtype out: sum over hit cameras / count
 *   attn_out (bs, HW, M, L*P)         fp32 out (softmax result, saved for backward) or NULL
 *   offsets_stride, logits_stride     elements between the rows of consecutive queries; 0 = dense
 *            (M*L*P*2 and M*L*P).  Non-zero strides let both tensors be column blocks of ONE
 *            Linear output -- sampling_offsets and attention_weights computed by a single GEMM
 *            over the concatenated weights.  The offsets stride
 *            must be even; g_offsets / g_logits of the backward use the same strides.
 * ------------------------------------------------------------------------------- */
int sca_fwd(const void* value, const int64_t* shapes, const int64_t* starts,
            const void* offsets, const void* logits, const float* ref_cam,
            const uint8_t* bev_mask, const uint32_t* hit_bits, void* slots, float* attn_out,
            int bs, int num_cam, int Nk, int M, int Dh, int L, int P, int D, int HW,
            int bev_w, int value_dtype, int coord_dtype, int64_t offsets_stride,
            int64_t logits_stride, void* stream);

/*   g_slots   (bs, HW, M*Dh) value_dtype   gradient w.r.t. `slots`
 *   g_value   (bs*num_cam, Nk, M, Dh) accumulator, zero-filled by the caller: fp32
 *             (accum_dtype = MSDA_F32), or fp16 (accum_dtype = MSDA_F16, 16-bit value dtypes only)
 *             holding gradient * (*accum_scale), a power of two from grad_amax_scale(); the fp16
 *             form halves the L2 sectors per update (red.global.add.noftz.v4.f16x2) and is
 *             turned into the value dtype by unscale_cast()
 *   g_offsets (bs, HW, M, L, P, 2) coord_dtype, fully written
 *   g_logits  (bs, HW, M, L*P) coord_dtype, fully written (softmax backward included)        */
int sca_bwd(const void* value, const int64_t* shapes, const int64_t* starts,
            const void* offsets, const void* logits, const float* ref_cam,
            const uint8_t* bev_mask, const uint32_t* hit_bits, const void* g_slots,
            void* g_value, void* g_offsets, void* g_logits,
            int bs, int num_cam, int Nk, int M, int Dh, int L, int P, int D, int HW,
            int bev_w, int value_dtype, int coord_dtype, int64_t offsets_stride,
            int64_t logits_stride, int accum_dtype, const float* accum_scale,
            void* g_value_tail, int tail_copies, int tail_pixels,
            void* coarse_records, const int32_t* hit_index, const int32_t* hit_count, void* stream);
/*   coarse_records, hit_index, hit_count (optional, all three or none; 16-bit value dtypes, Dh = 32,
 *             P <= 8): grad_value of the COARSE pyramid levels -- the longest suffix of at most two levels
 *             holding <= 2048 pixels together: 6 % of the pixels and 46 % of the updates at the base
 *             config -- is not scattered with reductions (the reference-era design: one global atomicAdd
 *             per corner and channel, ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:106-146 and mmcv's
 *             ms_deform_attn_col2im_bilinear) but accumulated on the tensor cores: the kernel writes one
 *             16-byte record per coarse sample into coarse_records (sca_coarse_workspace_bytes() bytes,
 *             contents undefined on entry and exit) and a second kernel computes, per (camera, head),
 *             patch[pixel][channel] += W^T[pixel][row] . g_slots[row][channel] with tcgen05.mma, the
 *             fp32 patch living in tensor memory, walking hit_index / hit_count (bev_hit_lists() or
 *             bev_point_sampling()).  The patch sums are added to g_value once per CTA.
 *   g_value_tail (optional, fp16 accumulator only): (tail_copies, bs*num_cam, tail_pixels, M, Dh) fp16,
 *             zero-filled: replicas of the LAST tail_pixels pixels of every value map -- the coarse
 *             pyramid levels, whose slots receive hundreds of updates each.  A CTA adds into replica
 *             blockIdx % (tail_copies + 1) (0 = g_value itself), which divides the number of fp16
 *             roundings a slot's running sum sees; unscale_cast() adds the replicas back.  NULL / 0:
 *             everything goes to g_value.                                                          */

/* Bytes of sca_bwd's coarse_records for these sizes; 0 when the tensor-core pass does not support them
 * (then pass NULL and everything is scattered with reductions). */
int64_t sca_coarse_workspace_bytes(int bs, int num_cam, int HW, int M, int Dh, int P, int value_dtype);

/* Per-camera ordered hit lists from the camera bit field (what bev_point_sampling writes when given
 * hit_index / hit_count): hit_bits (bs, HW) uint32 -- batch element 0 decides, the reference's quirk --
 * -> hit_index (num_cam, HW) int32 ascending, -1 padded; hit_count (num_cam,) int32. */
int bev_hit_lists(const uint32_t* hit_bits, int num_cam, int HW, int32_t* hit_index, int32_t* hit_count,
                  void* stream);

/* ---------------------------------------------------------------------------------
 * Fused temporal self-attention / decoder cross-attention core
 * (temporal_self_attention.py:204-279, decoder.py:299-350): softmax (with optional
 * logit clamp), location = ref + offset / (W_l, H_l), sampling, and the mean over the
 * `Q` queue entries (Q = 2 for TSA, Q = 1 for CustomMSDeformableAttention).
 *
 *   value    (bs*Q, Nk, M, Dh)        value_dtype; batch index = b*Q + j
 *   offsets  (bs, Nq, M, Q, L, P, 2)  coord_dtype raw Linear output
 *   logits   (bs, Nq, M, Q, L*P)      coord_dtype raw Linear output
 *   ref      (bs*Q, Nq, L, 2)         fp32 reference points
 *   out      (bs, Nq, M*Dh)           value_dtype: (1/Q) * sum_j MSDA_j
 *   clamp    < 0 : no clamp; otherwise logits are clamped to [-clamp, clamp]
 *   bev_w    > 0 when the Nq queries are the cells of a (Nq / bev_w) x bev_w grid in row-major
 *            order (TSA): work is tiled in 2-D patches of that grid; 0 otherwise (decoder)
 * ------------------------------------------------------------------------------- */
int tsa_fwd(const void* value, const int64_t* shapes, const int64_t* starts,
            const void* offsets, const void* logits, const float* ref, void* out,
            int bs, int Q, int Nk, int M, int Dh, int L, int P, int Nq, int bev_w,
            float clamp, int value_dtype, int coord_dtype, int64_t offsets_stride,
            int64_t logits_stride, void* stream);

int tsa_bwd(const void* value, const int64_t* shapes, const int64_t* starts,
            const void* offsets, const void* logits, const float* ref, const void* g_out,
            void* g_value, void* g_offsets, void* g_logits,
            int bs, int Q, int Nk, int M, int Dh, int L, int P, int Nq, int bev_w,
            float clamp, int value_dtype, int coord_dtype, int64_t offsets_stride,
            int64_t logits_stride, int accum_dtype, const float* accum_scale, void* stream);

/* Helpers of the fp16 gradient accumulator.
 *   grad_amax_scale: ws is a zero-initialised fp32 scratch of >= 64 floats (left zeroed except
 *     ws[16]); on return (stream order) ws[16] = 2^floor(log2(limit / max|g|)) (1 when g is all zero),
 *     i.e. a single scaled contribution is <= limit in fp16.  Overflow is excluded BY CONSTRUCTION when
 *     limit <= 32768 / R with R = the number of (query, head) rows that can add into one slot of a value
 *     map (= the number of queries): the attention weights of a row sum to one and bilinear weights are
 *     <= 1, so a row adds at most max|g| to any slot and a slot's sum stays <= 32768 < 65504 whatever the
 *     signs and locations.  The host passes limit = min(4, 32768 / Nq).
 *   unscale_cast: out[i] = (out_dtype)((acc_f16[i] + replicas) / *scale).  `tail` = the replicas
 *     of sca_bwd's g_value_tail (or NULL): acc is n / map_elems value maps of map_elems elements
 *     (Nk*M*Dh); the last tail_elems (tail_pixels*M*Dh) of each also sum tail_copies replica maps.
 *     colsum_out (C,) out_dtype or NULL: additionally the sums over all rows of the (n / C, C) view of
 *     `out` -- the bias gradient of the value projection -- saving a pass over the tensor; `partial` is
 *     the row kernels' zero-initialised scratch (64 + 8 * C floats, see below); C = 8 * (a divisor of 256).
 *     overflow_flag (may be NULL): a device int32 that is OR-ed with 1 when a non-finite accumulator
 *     slot is met -- a sticky word the host can poll at its leisure (defence in depth behind the
 *     scale bound above; nothing synchronises on the good path). */
int grad_amax_scale(const void* g, int64_t n, int dtype, float limit, float* ws, void* stream);
/* The same, and `zero` (zero_bytes, a multiple of 16, 16-byte aligned) is cleared by the same launch: the
 * accumulator the scale is for, instead of a fill launch of its own (ABI 8). */
int grad_amax_scale_zero(const void* g, int64_t n, int dtype, float limit, float* ws, void* zero,
                         int64_t zero_bytes, void* stream);
int unscale_cast(const void* acc_f16, void* out, const float* scale, int64_t n, int out_dtype,
                 const void* tail, int tail_copies, int64_t map_elems, int64_t tail_elems,
                 void* colsum_out, float* partial, int C, int32_t* overflow_flag, void* stream);
/* The same with `out` a column block of a wider row-major matrix (ABI 8): row r of the (n / C, C) view of the
 * result is written at out + r * out_row_stride elements (0 = contiguous; no tail replicas in this form).
 * The hoisted value projections of the encoder collect the gradients of all layers' projections side by side
 * this way, so that the feature gradient and the weight gradients are one GEMM each. */
int unscale_cast_strided(const void* acc_f16, void* out, const float* scale, int64_t n, int out_dtype,
                         const void* tail, int tail_copies, int64_t map_elems, int64_t tail_elems,
                         void* colsum_out, float* partial, int C, int32_t* overflow_flag,
                         int64_t out_row_stride, void* stream);

/* ---------------------------------------------------------------------------------
 * Row-wise companions of the attention kernels inside a BEVFormer layer (SURVEY.md section
 * 8f rank 3, "encoder remainder"; the reference builds them through mmcv bricks,
 * custom_base_transformer_layer.py:142-161): LayerNorm forward / backward over (rows, C)
 * activations, and the column sum that is the bias gradient of a Linear layer.
 *
 *   x, y, dy, dx  (rows, C) dtype;  gamma, beta (C,) dtype;  mean, rstd (rows,) fp32
 *   dgamma_dbeta  (2, C) dtype out: row 0 = d gamma, row 1 = d beta
 *   partial       fp32 scratch of 64 + 8 * 2 * C floats (ABI 8; 64 + 8 * n * C for a kernel with n output
 *                 rows), ALL ZERO on entry and left all zero: a ticket counter plus eight fp32 strips;
 *                 every CTA reduces its partial sums into one of them (reductions on one line serialise
 *                 in L2, hence several strips) and the last CTA to finish sums the strips into the
 *                 output (one launch); one scratch buffer per stream
 *   C must be 128, 256, 512 or 1024 for LayerNorm; a multiple of 16 bytes per row for colsum.
 * ------------------------------------------------------------------------------- */
int rowops_workspace_rows(void);
int ln_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
           int64_t rows, int C, float eps, int dtype, void* stream);
int ln_bwd(const void* x, const void* dy, const void* gamma, const float* mean, const float* rstd,
           void* dx, void* dgamma_dbeta, float* partial, int64_t rows, int C, int dtype, void* stream);
/* Post-norm block "y = LayerNorm(linear_out + residual)" of a transformer layer
 * (custom_base_transformer_layer.py:142-161 applied after spatial_cross_attention.py:173 /
 * temporal_self_attention.py:289 / the FFN's identity add) in one pass each way:
 *   ln_residual_fwd: s = (dtype)(x + residual) is written to sum_out (may alias x) and
 *     normalised; bit-identical to an add kernel followed by ln_fwd.
 *   ln_bwd_dxsum: ln_bwd that also returns the column sums of dx -- the bias gradient of the Linear
 *     layer that produced x -- as a third row: dgamma_dbeta_dxsum is (3, C), partial holds
 *     64 + 8 * 3 * C floats.                                                                      */
int ln_residual_fwd(const void* x, const void* residual, const void* gamma, const void* beta,
                    void* sum_out, void* y, float* mean, float* rstd, int64_t rows, int C, float eps,
                    int dtype, void* stream);
int ln_bwd_dxsum(const void* x, const void* dy, const void* gamma, const float* mean,
                 const float* rstd, void* dx, void* dgamma_dbeta_dxsum, float* partial, int64_t rows,
                 int C, int dtype, void* stream);
int colsum(const void* x, void* out, float* partial, int64_t rows, int C, int dtype, int out_dtype,
           void* stream);
/* ReLU backward fused with the bias gradient of the Linear layer in front of the activation (the FFN's
 * first Linear, custom_base_transformer_layer.py:142-155): dx = (y > 0 ? dy : 0) for the activation's
 * forward output y, and colsum_out (C,) = column sums of dx; one pass instead of two.  dx must not
 * alias dy or y.                                                                                   */
int relu_bwd_colsum(const void* dy, const void* y, void* dx, void* colsum_out, float* partial,
                    int64_t rows, int C, int dtype, int out_dtype, float scale, void* stream);
/*   scale: factor on the surviving entries -- 1, or 1 / (1 - p) when y came from relu_dropout_fwd (y is
 *   zero where the unit was dropped, so y > 0 selects exactly the kept, active entries).           */

/* ---------------------------------------------------------------------------------
 * Dropout of the training configuration, fused into the row kernels.  The reference trains with
 * dropout 0.1 after the output projections of TemporalSelfAttention / SpatialCrossAttention
 * (temporal_self_attention.py:285-289, spatial_cross_attention.py:171-173) and ffn_drop 0.1 after the
 * FFN's activation and after its last Linear (mmcv FFN; bev_base_occ.py:127).  Masks are counter-based
 * (Philox4x32-10 over (element / 8, call site, step), keyed by a seed; 16 random bits per element) and
 * are RECOMPUTED in the backward: no mask tensor exists.
 *   rng_state  uint64[2] on the device: (seed, step); the caller advances `step` on the device once
 *              per training step (so a replayed CUDA graph draws new masks)
 *   key_save   uint64[2] on the device: the forward stores the (seed, step) it used; the matching
 *              backward call takes it as `key`
 *   site       distinguishes the dropout instances of one step;  p in (0, 1)
 *   ln_residual_dropout_fwd: s = (dtype)(dropout(x) + residual), normalised (ln_residual_fwd with the
 *              dropout of the block's last Linear output)
 *   ln_bwd_dxsum_dropout: dx = d s (the residual's gradient) and dx_masked = d x = mask * d s / (1 - p)
 *              (the Linear output's gradient); the third output row holds the column sums of dx_masked
 *   relu_dropout_fwd: x <- dropout(relu(x)) in place (the FFN's Linear-ReLU-Dropout); n % 8 == 0
 *   dropout_keep_mask: the keep mask itself as bytes, for tests (the i-th byte belongs to element i of
 *              a tensor of the given dtype processed under (key, site, p))
 * ------------------------------------------------------------------------------- */
int ln_residual_dropout_fwd(const void* x, const void* residual, const void* gamma, const void* beta,
                            void* sum_out, void* y, float* mean, float* rstd, int64_t rows, int C, float eps,
                            int dtype, const void* rng_state, void* key_save, uint32_t site, float p,
                            void* stream);
int ln_bwd_dxsum_dropout(const void* x, const void* dy, const void* gamma, const float* mean,
                         const float* rstd, void* dx, void* dx_masked, void* dgamma_dbeta_dxsum,
                         float* partial, int64_t rows, int C, int dtype, const void* key, uint32_t site,
                         float p, void* stream);
int relu_dropout_fwd(void* x, int64_t n, int dtype, const void* rng_state, void* key_save, uint32_t site,
                     float p, void* stream);
int dropout_keep_mask(uint8_t* mask, int64_t n, int dtype, const void* key, uint32_t site, float p,
                      void* stream);

/* Number of kernel launches this library has enqueued since load (all entry points);
 * bench.py reports the delta over its timed region as "gpu_launches". */
int64_t msda_launch_count(void);

/* ---------------------------------------------------------------------------------
 * Pre-processing in front of the encoder (reference
 * PerceptionTransformer.get_bev_features, transformer.py:119-298).
 *
 * bev_flatten_level: one pyramid level of the image features
 *   feat (bs, num_cam, C, hw) dtype  ->  rows [start, start + hw) of
 *   feat_flatten (num_cam, Nk, bs, C) dtype, with non-finite values zeroed (:246-247 without the
 *   reference's device->host `isfinite().all()` check) and cams_embeds (num_cam, C; may be NULL)
 *   and level_embed (C,) added (:249-253).
 * bev_rotate_nearest: the ego-motion rotation of the previous BEV (:182-203) =
 *   torchvision.transforms.functional.rotate(nearest, no expand, zero fill), all samples at once:
 *   prev_bev, out (bev_h*bev_w, bs, C) dtype (must not alias);
 *   theta (bs, 6) fp32: the inverse affine matrix of torchvision's rotate, rows divided by
 *   (bev_w / 2, bev_h / 2); xs (bev_w,), ys (bev_h,) fp32: the pixel-centre coordinates
 *   linspace(-n/2 + 0.5, n/2 - 0.5, n) (the host builds all three exactly as torchvision does).
 * ------------------------------------------------------------------------------- */
int bev_flatten_level(const void* feat, const void* cams_embeds, const void* level_embed,
                      void* feat_flatten, int bs, int num_cam, int C, int hw, int64_t Nk,
                      int64_t start, int dtype, void* stream);
int bev_rotate_nearest(const void* prev_bev, void* out, const float* theta, const float* xs,
                       const float* ys, int bs, int bev_h, int bev_w, int C, int dtype, void* stream);

/* ---------------------------------------------------------------------------------
 * Weight and bias gradient of a Linear layer y = x W^T + b in one launch (16-bit dtypes):
 *   dW[o, i] = sum_n dy[n, o] * x[n, i]      db[o] = sum_n dy[n, o]   (db may be NULL)
 *   dy (N, O), x (N, I), dW (O, I), db (O,): dtype (MSDA_BF16 / MSDA_F16), row-major, 16-byte
 *   aligned, O and I multiples of 8.  Shaped for the transformer layer's backward: small O x I,
 *   huge N.  The row dimension is split over the whole grid, tensor-core MMAs accumulate in fp32,
 *   partial tiles are reduced in `workspace`: fp32, linear_wgrad_workspace_floats(O, I) floats,
 *   ALL ZERO on entry and left all zero (one scratch per stream).  Replaces a split-K library GEMM +
 *   its reduction kernel + the separate column sum of the bias gradient.
 * ------------------------------------------------------------------------------- */
int64_t linear_wgrad_workspace_floats(int O, int I);
int linear_wgrad(const void* dy, const void* x, void* dW, void* db, float* workspace, int64_t N,
                 int O, int I, int dtype, void* stream);

/* ---------------------------------------------------------------------------------
 * Small-sequence multi-head self-attention of the decoders: replaces the
 * torch.nn.MultiheadAttention core behind mmcv's MultiheadAttention in
 *   projects/mmdet3d_plugin/maptrv2/modules/decoder.py:129-188 (inter-vector attention over the vectors with
 *   the one-to-one / one-to-many block mask, intra-vector attention over a vector's points) and in the
 *   detection decoder's DetrTransformerDecoderLayer (bevformer/modules/decoder.py:50-126 builds it).
 *
 *   out = dropout_p(softmax(scale * q k^T + mask)) v        per (group, head)
 *
 * A problem is one (group g, head h): S tokens x Dh.  Token s of group g is row
 *     s * seq_stride + (g / n_lo) * hi_stride + (g % n_lo) * lo_stride
 * of the matrices q / k / v / out (dtype; element (row, h * Dh + d) at row * ld + h * Dh + d), so the
 * (num_query, bs, C) activations of a decoder layer are attended in place under either grouping of
 * decoder.py:131-185 without the permute + contiguous copies.  ld* in elements; pointers and row strides
 * 16-byte aligned; Dh in {32, 64} takes the tensor-core kernels for 16-bit dtypes (mma.sync, fp32
 * accumulate, K / V of a problem staged once per CTA), any other supported case (fp32; Dh in {8, 16, 32, 64},
 * S <= 3072) a one-warp-per-row FMA kernel.  impl: 0 = choose, 1 = force the FMA path, 2 = require the
 * tensor-core path.  mha_impl() returns what 0 would choose (0 = unsupported).
 *   mask_bits / mask_bits_t  (S, ceil(S / 32)) uint32 from mha_pack_mask(): bit k of row q set = query q may
 *              not attend to key k (the reference's boolean attn_mask, shared by all groups and heads);
 *              the second matrix is the transpose; NULL = no mask.  A row with no admissible key yields 0.
 *   lse        (G * H, S) fp32, written by the forward: log sum exp of the row's scaled scores.
 *   delta      (G * H, S) fp32 scratch of the backward (rowsum(grad_out * out)).
 *   dq, dk, dv have the token layout of q, k, v and their own row strides ld_dq / ld_dk / ld_dv (0 = those of
 *              q / k / v; one (rows, 3C) buffer can receive dq | dk | dv); every element of a token row's head
 *              slice is written.
 *   dropout    on the attention weights, counter-based and recomputed in the backward (rng_state,
 *              key_save / key, site, p: as in ln_residual_dropout_fwd).  mha_keep_mask() writes the keep
 *              mask ((G * H, S, S) bytes) for tests.
 * ------------------------------------------------------------------------------- */
int mha_impl(int G, int H, int S, int Dh, int dtype);
int mha_pack_mask(const uint8_t* mask, int S, uint32_t* bits, uint32_t* bits_t, void* stream);
int mha_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int64_t ldq, int64_t ldk,
            int64_t ldv, int64_t ldo, const uint32_t* mask_bits, int G, int H, int S, int Dh,
            int64_t seq_stride, int64_t hi_stride, int64_t lo_stride, int n_lo, float scale, int dtype,
            int impl, const void* rng_state, void* key_save, uint32_t site, float p, void* stream);
int mha_bwd(const void* q, const void* k, const void* v, const void* out, const void* grad_out,
            const float* lse, float* delta, void* dq, void* dk, void* dv, int64_t ldq, int64_t ldk,
            int64_t ldv, int64_t ldo, int64_t ld_dq, int64_t ld_dk, int64_t ld_dv,
            const uint32_t* mask_bits, const uint32_t* mask_bits_t, int G, int H,
            int S, int Dh, int64_t seq_stride, int64_t hi_stride, int64_t lo_stride, int n_lo, float scale,
            int dtype, int impl, const void* key, uint32_t site, float p, void* stream);
int mha_keep_mask(uint8_t* keep, int P, int S, const void* key, uint32_t site, float p, void* stream);

/* ---------------------------------------------------------------------------------
 * DCNv3: replaces the compiled `DCNv3` extension of the reference,
 *   dcnv3_forward / dcnv3_backward (projects/mmdet3d_plugin/bevformer/backbones/ops_dcnv3/src/dcnv3.h:21-59,
 *   src/cuda/dcnv3_cuda.cu:28-173), bound by functions/dcnv3_func.py:16, :41-46, :55-60.
 *
 *   input   (N, H_in, W_in, group * group_channels)              dtype, channels-last, contiguous
 *   offset  (N, H_out, W_out, group * kernel_h * kernel_w * 2)   dtype, (w, h) pairs, kernel_w outer loop
 *   mask    (N, H_out, W_out, group * kernel_h * kernel_w)       dtype
 *   output  (N, H_out, W_out, group * group_channels)            dtype, fully overwritten
 *   H_out = (H_in + 2 pad_h - (dilation_h (kernel_h - 1) + 1)) / stride_h + 1 (validated, as the reference does)
 *   scratch dcnv3_scratch_floats() fp32 values (sampling positions in pixels; an fp32 copy of the mask for
 *           16-bit dtypes; the level table), caller-owned, rewritten by every call
 * Backward: grad_input is an fp32 ACCUMULATOR the caller zero-fills (the reference allocates it with
 *   at::zeros_like, dcnv3_cuda.cu:112); grad_offset / grad_mask fp32, every element written once.
 * The reference's im2col_step batching has no meaning here and is not an argument.  offset_scale > 0.
 * ------------------------------------------------------------------------------- */
int64_t dcnv3_scratch_floats(int N, int H_out, int W_out, int group, int kernel_h, int kernel_w, int dtype);
int dcnv3_fwd(const void* input, const void* offset, const void* mask, void* output, float* scratch, int N,
              int H_in, int W_in, int H_out, int W_out, int kernel_h, int kernel_w, int stride_h, int stride_w,
              int pad_h, int pad_w, int dilation_h, int dilation_w, int group, int group_channels,
              float offset_scale, int dtype, void* stream);
int dcnv3_bwd(const void* input, const void* offset, const void* mask, const void* grad_output,
              float* grad_input, float* grad_offset, float* grad_mask, float* scratch, int N, int H_in, int W_in,
              int H_out, int W_out, int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
              int dilation_h, int dilation_w, int group, int group_channels, float offset_scale, int dtype,
              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
