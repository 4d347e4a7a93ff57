/* iif_b200.h -- C ABI of the B200-native IIF classifier head.
 *
 * Drop-in boundary for the one hot path of kostas1515/iif: fc_cls GEMM -> per-class IIF logit
 * scale -> softmax-CE / sigmoid-BCE (forward + backward) and the label histogram that produces
 * the IIF weight vector.  The reference has no FFI (it is pure Python on torch); every entry point
 * below replaces the implicit ATen / cuBLAS dispatch made at the cited reference line
 * (paths relative to the reference root; cls/ = classification/, seg/ = instance_segmentation/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; the caller allocates every output
 *   - all matrices are row-major with an explicit leading dimension (elements, not bytes)
 *   - calls are stateless, stream-ordered and re-entrant across streams / ranks
 *   - return value: 0 = ok, <0 = argument error (IIF_E*), >0 = cudaError_t of the launch
 *   - `stream` is a cudaStream_t (CUstream) passed as void*
 *   - bf16 tensors are passed as void* (uint16 storage, torch.bfloat16 compatible)
 */
#ifndef IIF_B200_H_
#define IIF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IIF_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define IIF_API __attribute__((visibility("default")))
#else
#define IIF_API
#endif

enum {
  IIF_OK = 0,
  IIF_EINVAL = -1,       /* null pointer / negative size / bad enum */
  IIF_EALIGN = -2,       /* pointer or leading dimension violates the documented alignment */
  IIF_EUNSUPPORTED = -3, /* shape outside the supported range (e.g. C > 32768 for the row kernels) */
  IIF_EWORKSPACE = -4,   /* workspace too small */
  IIF_EDRIVER = -5       /* cuTensorMapEncodeTiled unavailable / failed */
};

/* IIF weighting variants, cls/custom.py:16-23 (CSV column names in seg/lvis_files/idf_1204.csv:1
 * use "prob" for REL). */
enum {
  IIF_VARIANT_RAW = 0,    /* ln(N/f)                 */
  IIF_VARIANT_SMOOTH = 1, /* ln((N+1)/(f+1)) + 1     */
  IIF_VARIANT_REL = 2,    /* ln((N-f)/f)             */
  IIF_VARIANT_NORMIT = 3, /* -ndtri(f/N)             */
  IIF_VARIANT_GOMBIT = 4, /* -ln(-ln(1 - f/N))       */
  IIF_VARIANT_BASE2 = 5,  /* log2(N/f)               */
  IIF_VARIANT_BASE10 = 6  /* log10(N/f)              */
};

enum { IIF_DTYPE_F32 = 0, IIF_DTYPE_BF16 = 1 };

/* Row-norm -> operand multiplier of the normalised classifiers (see iif_row_scale_from_norm). */
enum {
  IIF_NORM_NORMED = 0, /* T / (n^p + eps): mmdet NormedLinear / IIFNormedLinear, normed_predictor.py:36-40,70-76 */
  IIF_NORM_COS = 1,    /* T / (1 + n):     CosNorm_Classifier features, cls/resnet_cifar.py:67-69              */
  IIF_NORM_UNIT = 2    /* T / max(n, eps): unit rows (CosNorm weights :71, F.normalize)                        */
};

IIF_API int iif_abi_version(void);
IIF_API const char* iif_error_string(int code);
/* Number of kernels this library has launched since load (all streams); bench.py's gpu_launches. */
IIF_API uint64_t iif_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * (d) histograms + IIF weight vector
 * ------------------------------------------------------------------------------------------- */

/* counts[c] += #{i : labels[i] == c}, c in [0,C); labels outside [0,C) are ignored.
 * Replaces the O(N*C) numpy loops cls/imbalanced_dataset.py:112,127.  `counts` is accumulated
 * into (zero it first); integer, bit-exact. */
IIF_API int iif_hist_labels_i64(const int64_t* labels, int64_t n, int64_t* counts, int64_t num_classes,
                        void* stream);

/* Per-annotation (image_id, category) pairs -> instance_freq[c] (#annotations) and img_freq[c]
 * (#distinct images holding c): the two frequency columns of seg/lvis_files/idf_1204.csv (cols
 * 15-16; set-dedup semantics of seg/mmdet/datasets/dataset_wrappers.py:245-252).
 * `bitmap_ws`: num_classes * ceil(num_images/32) uint32 words, ZEROED by the caller; both outputs
 * are accumulated into.  Pairs with category outside [0,C) or image outside [0,num_images) are
 * ignored. */
IIF_API int iif_hist_images_dedup_i64(const int64_t* image_ids, const int64_t* categories, int64_t n,
                              int64_t num_images, int64_t num_classes, int64_t* img_freq,
                              int64_t* instance_freq, uint32_t* bitmap_ws, void* stream);
IIF_API size_t iif_hist_images_dedup_ws_bytes(int64_t num_images, int64_t num_classes);

/* counts[C] -> weight vector, float64 arithmetic rounded ONCE to float32 (cls/custom.py:14-26;
 * closed forms of the CSV columns read at seg/mmdet/models/losses/iif_loss.py:47-50).
 *   total   : N; <= 0 means sum(counts) (classification).  The CSVs use #images / sum(instance_freq).
 *   norm_p  : > 0 divides by the p-norm of the fp32 vector (iif_norm, cls/custom.py:25-26)
 *   out_f32 : [C] required;  out_f64 : [C] optional un-normalised float64 values (may be NULL) */
IIF_API int iif_weights_from_counts(const int64_t* counts, int64_t num_classes, int64_t total, int variant,
                            double norm_p, float* out_f32, double* out_f64, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (b) fused softmax cross-entropy forward + backward with the IIF logit scale
 * ------------------------------------------------------------------------------------------- */

/* One pass over the logits.  With a = z * iif (per column), p = softmax(a):
 *   l_i      = -cw[y_i] * w_i * log p_{i,y_i}          (0 when y_i == ignore_index or y_i outside [0,C))
 *   loss_i   = scale * l_i                              [B]   optional
 *   loss_sum = sum_i loss_i  (fixed-order, deterministic) [1] optional
 *   dz_ic    = iif_c * scale * cw[y_i] * w_i * (p_ic - [c == y_i])       optional, fp32 and/or bf16
 *   argmax_i = first index of max_c z_ic (RAW logits), rank_i = #{c: z_ic > z_iy} + #{c<y: z_ic == z_iy}
 * Replaces `pred*iif` + F.cross_entropy + weight_reduce_loss + their autograd:
 * cls/custom.py:28-36; seg/mmdet/models/losses/iif_loss.py:187-200, losses/utils.py:42-55.
 * scale = 1/B (cls 'mean'), 1 ('sum'/'none'), loss_weight/avg_factor or loss_weight/B (mmdet).
 * acc_counts[0..1] (optional) = #{rank_i < 1}, #{rank_i < 5} (cls/utils.py:165-179).
 * `scratch`: iif_loss_scratch_bytes(B) device bytes whose first int32 is zero on entry (left zero on
 * exit): per-CTA partial sums for the deterministic grid-wide reduction; required with loss_sum /
 * acc_counts.
 * Alignment: z and dz rows may use any leading dimension; 128-bit access is used when the base
 * pointers are 16-byte aligned and ldz / lddz / C are multiples of 4 (8 for bf16 dz).
 * Supported C: 1..32768. */
IIF_API int iif_softmax_ce_fwd_bwd(const float* z, int64_t ldz, const float* iif, const int64_t* label,
                           const float* class_weight, const float* sample_weight,
                           int64_t ignore_index, float scale, int64_t B, int64_t C, float* loss_i,
                           float* loss_sum, float* dz_f32, int64_t lddz_f32, void* dz_bf16,
                           int64_t lddz_bf16, float* lse, int32_t* argmax, int32_t* rank,
                           int32_t* acc_counts, int32_t* scratch, void* stream);
IIF_API size_t iif_loss_scratch_bytes(int64_t B);

/* Dual-label (Mixup) form: loss_i = scale * [lam * l_i(label_a) + (1 - lam) * l_i(label_b)] with l_i as above,
 * dz its gradient -- ONE pass over the logits where cls/custom.py:116-117 (Mixup.mixup_criterion) runs the
 * criterion twice.  argmax / rank refer to label_a.  Needs the 128-bit path (C % 4 == 0, aligned rows):
 * returns IIF_EUNSUPPORTED otherwise and the caller evaluates the two terms separately. */
IIF_API int iif_softmax_ce_mixup_fwd_bwd(const float* z, int64_t ldz, const float* iif, const int64_t* label_a,
                                 const int64_t* label_b, float lam, const float* class_weight,
                                 const float* sample_weight, int64_t ignore_index, float scale, int64_t B,
                                 int64_t C, float* loss_i, float* loss_sum, float* dz_f32, int64_t lddz_f32,
                                 void* dz_bf16, int64_t lddz_bf16, int32_t* argmax, int32_t* rank,
                                 int32_t* acc_counts, int32_t* scratch, void* stream);

/* out = softmax(z * iif) per row (seg/mmdet/models/losses/iif_loss.py:76) or, with
 * softmax == 0, out = z * iif (cls/custom.py:38, infer=True).  argmax/rank (optional) are taken
 * on the ADJUSTED logits here (cls/train.py:104-106). */
IIF_API int iif_scaled_activation(const float* z, int64_t ldz, const float* iif, int softmax, int64_t B,
                          int64_t C, float* out, int64_t ldo, const int64_t* label, int32_t* argmax,
                          int32_t* rank, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (b') sigmoid BCE forward + backward (no IIF scale; no one-hot tensor is materialised)
 * ------------------------------------------------------------------------------------------- */

/* t_ic = [c == y_i] for 0 <= y_i < C and y_i != ignore_index; valid_i = (y_i >= 0 && y_i != ignore_index)
 *   e_ic   = (1 - t) z + (1 + (pw_c - 1) t) * softplus(-z)                  (pw = pos_weight, optional)
 *   wgt_ic = valid_i * w_i * colw_c                                          (w, colw optional)
 *   loss_elem = scale * wgt * e [B,C] optional;  loss_i = row sums [B] optional;  loss_sum [1] optional
 *   dz = scale * wgt * ((1 - t) - (1 + (pw - 1) t) (1 - sigmoid(z)))
 * Replaces _expand_onehot_labels + F.binary_cross_entropy_with_logits + weight_reduce_loss:
 * seg/mmdet/models/losses/cross_entropy_loss.py:53-111 (used by CrossEntropyLoss(use_sigmoid) and
 * FasaIIFLoss(use_sigmoid), fasa_iif_loss.py:35-36) and cls FocalLoss(gamma=0), cls/custom.py:61-73
 * (colw = per-class weights; scale = 1/(B*C) for 'mean', 1/B for 'sum'). */
IIF_API int iif_sigmoid_bce_fwd_bwd(const float* z, int64_t ldz, const int64_t* label,
                            const float* pos_weight, const float* col_weight,
                            const float* sample_weight, int64_t ignore_index, float scale,
                            int64_t B, int64_t C, float* loss_elem, int64_t ldl, float* loss_i,
                            float* loss_sum, float* dz_f32, int64_t lddz_f32, void* dz_bf16,
                            int64_t lddz_bf16, int32_t* scratch, void* stream);

/* Focal form of the same kernel (cls FocalLoss with gamma > 0, cls/custom.py:74-89):
 *   p = sigmoid(z), q = t p + (1-t)(1-p),  e_ic = -log(q) (1-q)^gamma * (alpha > 0 ? (t alpha + (1-t)(1-alpha)) : 1)
 * with wgt / scale / outputs as above (no pos_weight).  Evaluated in logit space (softplus), so it stays
 * accurate where the reference's fp32 sigmoid -> log saturates.  gamma must be > 0 (gamma = 0 is
 * iif_sigmoid_bce_fwd_bwd), alpha <= 0 means "no class balance". */
IIF_API int iif_sigmoid_focal_fwd_bwd(const float* z, int64_t ldz, const int64_t* label, float gamma, float alpha,
                              const float* col_weight, const float* sample_weight, int64_t ignore_index,
                              float scale, int64_t B, int64_t C, float* loss_elem, int64_t ldl, float* loss_i,
                              float* loss_sum, float* dz_f32, int64_t lddz_f32, void* dz_bf16,
                              int64_t lddz_bf16, int32_t* scratch, void* stream);

/* Sigmoid BCE with ALREADY-EXPANDED labels: `target` is a dense [B,C] fp32 matrix of (soft) targets in [0,1] -- the
 * branch binary_cross_entropy takes when pred.dim() == label.dim(), seg/mmdet/models/losses/cross_entropy_loss.py:100-106.
 *   e = (1 - t) z + (1 + (pw - 1) t) softplus(-z);  loss_elem = scale * w * e;  dz = scale * w * d e / d z
 * `weight`: element weights [B,C] (ldw >= C), a per-row vector [B] (ldw == 0), or NULL.  loss_i = row sums. */
IIF_API int iif_sigmoid_bce_dense_fwd_bwd(const float* z, int64_t ldz, const float* target, int64_t ldt,
                                  const float* pos_weight, const float* weight, int64_t ldw, float scale,
                                  int64_t B, int64_t C, float* loss_elem, int64_t ldl, float* loss_i,
                                  float* loss_sum, float* dz_f32, int64_t lddz, int32_t* scratch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FASA bookkeeping around loss_cls and the many / median / low-shot accuracy (8f-2)
 * ------------------------------------------------------------------------------------------- */

/* cum_labels[c] += #{i : label_i == c};  cum_losses[c] += sum_{i : label_i == c} rowsum(loss_i)  -- the python loop over
 * label.unique() with its .item() syncs at seg/mmdet/models/losses/fasa_iif_loss.py:154-160.  `loss` is [B] (loss_cols
 * = 1) or [B, loss_cols] (sigmoid mode: the reference sums the row).  A negative label indexes from the end like the
 * reference's python indexing; labels outside [-num_bins, num_bins) (where the reference raises IndexError) are skipped.
 * Deterministic (one warp per class, fixed-order sums). */
IIF_API int iif_class_accumulate(const int64_t* label, const float* loss, int64_t ldl, int64_t loss_cols, int64_t B,
                         int64_t num_bins, float* cum_losses, float* cum_labels, void* stream);

/* FasaBBoxHead.fa_update / fa_update_push (seg/mmdet/models/roi_heads/bbox_heads/fasa_bbox_head.py:118-148): for every
 * class c present in `label`, mean and variance (unbiased for n > 1) of the class's feature rows, folded into the running
 * statistics: first sighting (feature_used[c] == 0) stores them and bumps feature_used[c], later ones blend with
 * `decay` (decay * new + (1 - decay) * old).  x [B,D] fp32; feature_mean / feature_var [num_bins, D] (ldm); B <= 8192.
 * `ws_zeroed`: iif_class_feature_stats_ws_bytes(num_bins) bytes, zero on entry (left zero on exit). */
IIF_API int iif_class_feature_stats(const float* x, int64_t ldx, const int64_t* label, int64_t B, int64_t D,
                            int64_t num_bins, float decay, float* feature_mean, float* feature_var, int64_t ldm,
                            float* feature_used, int32_t* ws_zeroed, void* stream);
IIF_API size_t iif_class_feature_stats_ws_bytes(int64_t num_bins);

/* shot_acc (cls/per_shot_acc.py:62-105): per-class test / correct counts of (preds, labels) -- integer, bit-exact --
 * and the mean class accuracy over the classes PRESENT in `labels` whose TRAIN count is > many_shot_thr (out3[0]),
 * < low_shot_thr (out3[2]) or in between (out3[1]); 0 for an empty group.  class_acc [C] (optional): correct / test,
 * -1 for classes absent from `labels`.  preds int32 (the argmax output of the loss rows), labels int64. */
IIF_API int iif_shot_accuracy(const int32_t* preds, const int64_t* labels, int64_t n, const int64_t* train_counts,
                      int64_t C, int64_t many_shot_thr, int64_t low_shot_thr, int64_t* test_counts,
                      int64_t* correct_counts, double* out3, double* class_acc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * small elementwise helpers on [rows, cols] matrices
 * ------------------------------------------------------------------------------------------- */

/* out[i,c] = in[i,c] * g[i * g_stride]  (g == NULL -> 1).  g_stride 0 = one device scalar (the
 * upstream autograd grad of a reduced loss), 1 = per-row vector (reduction='none').  Also the
 * fp32 -> bf16 operand cast (out_dtype = IIF_DTYPE_BF16). */
IIF_API int iif_scale_rows(const float* in, int64_t ldi, const float* g, int64_t g_stride, int64_t rows,
                   int64_t cols, void* out, int out_dtype, int64_t ldo, void* stream);

/* data[0..n) *= *g_dev in place (fp32 or bf16 storage); a no-op launch when *g_dev == 1.  Applies the upstream scalar
 * of autograd to gradients that the one-launch head step already formed in the forward call. */
IIF_API int iif_scale_inplace(void* data, int dtype, int64_t n, const float* g_dev, void* stream);

/* db[c] = alpha * sum_i dz[i,c]   (alpha_dev == NULL -> 1; fixed-order, deterministic).
 * The bias gradient of AddmmBackward (a10). */
IIF_API int iif_colsum(const void* dz, int dz_dtype, int64_t lddz, const float* alpha_dev, int64_t rows,
               int64_t cols, float* db, void* stream);

/* ---------------------------------------------------------------------------------------------
 * normalised classifiers: the reference normalises the OPERANDS of fc_cls and then calls
 * F.linear (seg/mmdet/models/utils/normed_predictor.py:36-40,70-76; cls/resnet_cifar.py:66-77); so does
 * this path, with three row kernels around the head's GEMMs.  All fp32, rows with any leading dimension.
 * ------------------------------------------------------------------------------------------- */

/* Per row i: n = | pre_i x_i |_2 (pre optional: the IIF weight of a class row), r = r_mode(n; T, p, eps):
 *   a[i] = pre_i r(n)                 (y_i = a[i] x_i is the normalised operand row)
 *   c[i] = pre_i^3 r'(n) / n          (optional; backward: dx_i = a[i] g_i + c[i] (x_i . g_i) x_i)
 *   norm[i] = n                       (optional) */
IIF_API int iif_row_scale_from_norm(const float* x, int64_t ldx, int64_t rows, int64_t cols, const float* pre, int mode,
                            float temperature, float power, float eps, float* a, float* c, float* norm, void* stream);
/* out[i] = u_i . v_i */
IIF_API int iif_row_dot(const float* u, int64_t ldu, const float* v, int64_t ldv, int64_t rows, int64_t cols, float* out,
                void* stream);
/* out_i = a[i] u_i + (b[i] b2[i]) v_i   (a, b, b2 optional = 1; v == NULL drops the second term) */
IIF_API int iif_rows_axpby(const float* u, int64_t ldu, const float* a, const float* v, int64_t ldv, const float* b,
                   const float* b2, int64_t rows, int64_t cols, float* out, int64_t ldo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (a)/(c) fc_cls GEMMs.  *_bf16: tcgen05.mma (TMEM accumulators, TMA-fed), bf16 operands, fp32
 * accumulation.  *_f32: FFMA, fp32 operands (the 1e-5 parity mode).
 * `ws`: 16-byte aligned workspace of at least iif_gemm_ws_bytes(...) bytes: a 4 KB header of split-K
 * arrival counters followed by the fp32 partial tiles.  The HEADER MUST BE ZERO when the workspace is
 * first used (cudaMemset it once at allocation); every launch leaves it zero again, so one workspace
 * serves any number of stream-ordered launches.  One workspace per stream (not re-entrant across
 * streams).
 * bf16 alignment: base pointers 16 bytes; leading dimensions of bf16 operands multiples of 8.
 * ------------------------------------------------------------------------------------------- */

/* Z[B,C] = X[B,D] W[C,D]^T + bias (nn.Linear / fc_cls forward: cls/resnet_pytorch.py:219,293;
 * cls/resnet_cifar.py:192,211; seg/.../bbox_heads/bbox_head.py:118, convfc_bbox_head.py:188).
 * Epilogue: + bias[c] (optional); z (raw, optional) and zs = z * col_scale[c] (optional, the IIF
 * adjusted logits of cls/custom.py:38) are written from the same accumulator. */
IIF_API int iif_linear_fwd_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                        const float* col_scale, float* z, int64_t ldz, float* zs, int64_t ldzs,
                        int64_t B, int64_t D, int64_t C, void* ws, size_t ws_bytes, void* stream);
IIF_API int iif_linear_fwd_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                       const float* col_scale, float* z, int64_t ldz, float* zs, int64_t ldzs,
                       int64_t B, int64_t D, int64_t C, void* stream);

/* fp32 parity mode on the tensor cores: expand an fp32 operand into six bf16 copies along the contraction dimension
 * (v = v_h + v_m + v_l; A side: l m h m h h, B side: h m l h m h) so that ONE call of the bf16 GEMM entry points above /
 * below with a six times longer K forms  A_l B_h + A_m B_m + A_h B_l + A_m B_h + A_h B_m + A_h B_h  in fp32 (smallest
 * terms first) -- the product to ~2^-24 relative, i.e. within the 1e-5 bar, at tensor-core speed (csrc/split3.cu).
 *   k_along_rows = 0: in [rows, cols] -> out [rows, 6 * pad8(cols)] (copy k at column k * pad8(cols); padding zero)
 *   k_along_rows = 1: in [rows, cols] -> out [6 * pad8(rows), cols] (copy k at row k * pad8(rows); padding zero)
 *   side_b: 0 = the A operand's term order, 1 = the B operand's. */
IIF_API int iif_split3_bf16(const float* in, int64_t ldi, int64_t rows, int64_t cols, int k_along_rows, int side_b,
                    void* out, int64_t ldo, void* stream);

/* dX[B,D] = alpha * dZ[B,C] W[C,D]   (AddmmBackward, a10).  alpha_dev: device scalar or NULL. */
IIF_API int iif_linear_bwd_dx_bf16(const void* dz, int64_t lddz, const void* w, int64_t ldw,
                           const float* alpha_dev, void* dx, int dx_dtype, int64_t lddx, int64_t B,
                           int64_t D, int64_t C, void* ws, size_t ws_bytes, void* stream);
IIF_API int iif_linear_bwd_dx_f32(const float* dz, int64_t lddz, const float* w, int64_t ldw,
                          const float* alpha_dev, float* dx, int64_t lddx, int64_t B, int64_t D,
                          int64_t C, void* stream);

/* dW[C,D] = alpha * dZ[B,C]^T X[B,D]   (AddmmBackward, a10). */
IIF_API int iif_linear_bwd_dw_bf16(const void* dz, int64_t lddz, const void* x, int64_t ldx,
                           const float* alpha_dev, float* dw, int64_t lddw, int64_t B, int64_t D,
                           int64_t C, void* ws, size_t ws_bytes, void* stream);
IIF_API int iif_linear_bwd_dw_f32(const float* dz, int64_t lddz, const float* x, int64_t ldx,
                          const float* alpha_dev, float* dw, int64_t lddw, int64_t B, int64_t D,
                          int64_t C, void* stream);

/* dX, dW and db of AddmmBackward in ONE launch (they share dZ and together fill the SMs):
 *   dX[B,D] = alpha dZ W (skipped when dx == NULL: frozen backbone, cls/train.py:123-145),
 *   dW[C,D] = alpha dZ^T X,  db[C] = alpha sum_i dZ[i,:] (optional; computed on the tensor cores). */
IIF_API int iif_linear_bwd_bf16(const void* dz, int64_t lddz, const void* x, int64_t ldx, const void* w, int64_t ldw,
                        const float* alpha_dev, void* dx, int dx_dtype, int64_t lddx, float* dw, int64_t lddw,
                        float* db, int64_t B, int64_t D, int64_t C, void* ws, size_t ws_bytes, void* stream);

/* Debug hooks (tools/tc_timing.py).  iif_debug_timing: when `buf` (device, 16 int64 per CTA of the next
 * tensor-core launches) is non-NULL every CTA records %globaltimer at its phase boundaries; NULL switches
 * it off.  iif_debug_capacity: resident-CTA capacity of the current device for the tensor-core kernel
 * (the split-K rendezvous is only enabled for grids that fit it). */
IIF_API void iif_debug_timing(long long* buf);
IIF_API void iif_debug_timing_fused(long long* buf);       /* one-launch step: 32 int64 per CTA, see tools/fused_timing.py */
IIF_API int iif_debug_fused_plan(int64_t B, int64_t D, int64_t C, int need_dx, int sms, int* out12);   /* host only: the one-launch step's plan */
IIF_API void iif_debug_timing_allreduce(long long* buf);   /* 8 int64 per CTA: start, after handshake, after data, end */
IIF_API int iif_debug_capacity(int* detail6 /* host, optional: occupancy API, by smem, by regs, regs, smem/SM, static smem */);

/* Split-K and the loss-fused backward launch rendezvous INSIDE a grid, which is only safe while every CTA
 * of the grid can be resident at once.  A kernel that overlaps these launches and itself blocks on other
 * GPUs (iif_allreduce_mean_f32) must not be able to starve them of resident-CTA slots: reserve its
 * footprint here (1 slot per CTA of <= 256 threads, 2 per larger CTA; process-wide, default 0) and the
 * planner only counts on the remaining slots (falling back to unsplit / unfused launches when a grid no
 * longer fits). */
IIF_API int iif_gemm_reserve_slots(int slots);
/* By default the launches that rendezvous inside a grid are only made as COOPERATIVE launches (the driver
 * guarantees co-residency whatever else runs on the device), which bounds such grids by the runtime's own
 * occupancy answer.  A caller that guarantees exclusive use of the SMs while the head's launches run may allow
 * the larger grids of round 1 (up to this library's own resident-CTA bound, minus iif_gemm_reserve_slots). */
IIF_API int iif_gemm_assume_exclusive(int on);

/* Workspace (bytes) the three bf16 GEMMs of a head of this shape may need (max over the three). */
IIF_API size_t iif_gemm_ws_bytes(int64_t B, int64_t D, int64_t C);

/* ---------------------------------------------------------------------------------------------
 * whole head, one call: fc_cls -> IIF softmax-CE fwd+bwd -> dX, dW, db   (bf16 GEMM mode)
 * ------------------------------------------------------------------------------------------- */
typedef struct iif_head_args {
  /* inputs */
  const void* x;  int64_t ldx;        /* bf16 [B,D] */
  const void* w;  int64_t ldw;        /* bf16 [C,D] */
  const float* bias;                  /* [C] or NULL */
  const float* iif;                   /* [C] or NULL (plain CE) */
  const int64_t* label;               /* [B] */
  const float* class_weight;          /* [C] or NULL */
  const float* sample_weight;         /* [B] or NULL */
  int64_t ignore_index;
  float scale;                        /* see iif_softmax_ce_fwd_bwd */
  int64_t B, D, C;
  /* outputs (any may be NULL to skip, except z and dz_bf16 which the chain needs) */
  float* z;       int64_t ldz;        /* raw logits fp32 [B,C] */
  float* loss_i;                      /* [B] */
  float* loss_sum;                    /* [1] */
  void* dz_bf16;  int64_t lddz;       /* bf16 [B,lddz], lddz % 8 == 0 */
  void* dx;       int dx_dtype; int64_t lddx;   /* [B,D]; NULL = frozen backbone (--decoup / selectp=1) */
  float* dw;      int64_t lddw;       /* fp32 [C,D] */
  float* db;                          /* fp32 [C] or NULL */
  int32_t* argmax; int32_t* rank; int32_t* acc_counts;
  /* scratch */
  int32_t* scratch;                   /* iif_loss_scratch_bytes(B), first int32 zero */
  void* ws; size_t ws_bytes;          /* iif_gemm_ws_bytes(B,D,C) */
  int32_t flags;                      /* IIF_HEAD_* */
} iif_head_args;

/* Launches the 3 kernels of one head step on `stream` (fc_cls GEMM; fused loss; dX+dW+db GEMM group),
 * chained by programmatic dependent launch (cls/train.py:66-77 collapsed to the head;
 * seg/.../bbox_head.py:118 + :269-274 + autograd). */
#define IIF_HEAD_NO_FUSED_LOSS 1      /* keep the loss rows in their own launch (3 launches per step) */
#define IIF_HEAD_NO_PERSISTENT 2      /* do not use the one-launch persistent step (csrc/head_fused.cu) */
#define IIF_HEAD_LOW_REGS 4           /* the 128-register build of the one-launch step: room for TWO all-reduce lanes of
                                         128-thread CTAs on every SM next to it (set by iif_pipeline_set_allreduce) */
IIF_API int iif_head_fwd_bwd_bf16(const iif_head_args* args, void* stream);
/* The loss rows + AddmmBackward in ONE launch: every CTA of the backward launch first computes its share
 * of the softmax-CE rows (reads args->z, writes loss_i / dz_bf16 / argmax / rank), the grid meets at a
 * counter in the workspace header, then dX, dW, db are formed from dZ out of L2.  The X / W operand tiles
 * are requested before the loss rows run (after the programmatic-dependency wait).  Returns IIF_EUNSUPPORTED (nothing launched) when the shape
 * does not qualify: grid above the resident-CTA capacity, C > 4096 or C % 4 != 0, unaligned rows. */
IIF_API int iif_loss_linear_bwd_bf16(const iif_head_args* args, void* stream);
/* Number of launches iif_head_fwd_bwd_bf16 makes for these arguments (2 or 3); negative = argument error. */
IIF_API int iif_head_launches(const iif_head_args* args);

/* ---------------------------------------------------------------------------------------------
 * host-batch pipeline: the head step with the step's features / labels in (pinned) HOST memory.
 * `slot_args[i]` describes slot i's DEVICE buffers exactly as for iif_head_fwd_bwd_bf16 (x and label
 * are the device staging buffers the host batch is copied into; w, bias, iif are shared parameters;
 * ws may be shared by all slots).  Per submit: H2D of x [B,D] bf16 and label [B] int64 on a copy
 * stream, the launches of the step (two, or three) on the library's single compute stream, D2H of the loss
 * scalar on a third stream -- slot i+1's copy overlaps slot i's kernels.  This is the data-loader ->
 * criterion(output, target) -> loss.item() sequence of cls/train.py:66-77 for the head alone.
 * ------------------------------------------------------------------------------------------- */
typedef struct iif_pipeline iif_pipeline;
IIF_API int iif_pipeline_create(iif_pipeline** out, const iif_head_args* slot_args, int nslots);
/* Enqueue one step on `slot` (waits, on the device, for the slot's previous step). host_x: [B,D] bf16
 * contiguous, host_label: [B] int64, host_loss: 4 bytes; all three should be pinned. Never blocks. */
IIF_API int iif_pipeline_submit(iif_pipeline* p, int slot, const void* host_x, const int64_t* host_label,
                                float* host_loss);
/* The same step with the inputs already in the slot's device buffers (no copies, no loss read-back:
 * iif_pipeline_wait on such a slot returns IIF_EINVAL until it is submitted with host batches again). */
IIF_API int iif_pipeline_submit_device(iif_pipeline* p, int slot);
/* Make `stream` wait, on the device, for everything enqueued so far on ALL of the pipeline's streams
 * (compute, copies, every comm lane). */
IIF_API int iif_pipeline_join(iif_pipeline* p, void* stream);
/* Data-parallel runs: after every step, all-reduce(mean) the slot's gradient slice with
 * iif_allreduce_mean_f32 on the pipeline's comm stream (arguments as there; slot i's slice starts at
 * slot_offsets_elems[i]); the slot is not reused before its all-reduce has finished.  num_lanes (1..4)
 * comm streams take the all-reduces of consecutive steps in turn, so that many can be in flight at once
 * (an 8 MB all-reduce is a chain of NVLink round trips longer than the step it overlaps). */
IIF_API int iif_pipeline_set_allreduce(iif_pipeline* p, void* const* peer_bufs_dev, void* const* peer_flags_dev,
                                       void* multicast_ptr, int rank, int world, const int64_t* slot_offsets_elems,
                                       int64_t n_elems, int num_ctas, int num_threads, int num_lanes);
/* The pipeline's streams (cudaStream_t), e.g. to record timing events on them; any pointer may be NULL. */
IIF_API int iif_pipeline_get_streams(iif_pipeline* p, void** h2d, void** compute, void** d2h, void** comm);
/* Staged mode -- ONE driver call per step.  iif_pipeline_enable_staged gives every slot pinned HOST staging
 * buffers (iif_pipeline_staging: x [B,D] bf16, label [B] int64, and the 4-byte loss, which the loss kernel
 * stores straight into mapped pinned memory) and captures one CUDA graph per slot: this slot's two launches
 * with the H2D copy of the NEXT slot's staged batch as a parallel branch.  Contract: slots are walked
 * round-robin and, when slot k is submitted, slot k+1's staging already holds the following batch (a data
 * loader one batch ahead); an out-of-order submit is still correct, it only pays an un-overlapped copy.
 * iif_pipeline_wait(slot) then blocks until that step's loss is in *host_loss.  Not combined with
 * iif_pipeline_set_allreduce. */
IIF_API int iif_pipeline_enable_staged(iif_pipeline* p);
IIF_API int iif_pipeline_staging(iif_pipeline* p, int slot, void** host_x, int64_t** host_label, float** host_loss);
IIF_API int iif_pipeline_submit_staged(iif_pipeline* p, int slot);
/* One step of EVERY slot (0 .. nslots-1, in order) from the slots' staging buffers as ONE graph launch: slot k's
 * launch with the H2D of slot k+1's staged batch as a parallel branch, each slot's loss event an external event node
 * (iif_pipeline_wait(slot) works per slot).  One driver call per nslots steps.  Contract: every slot's staging holds
 * its batch at the call and is not rewritten until iif_pipeline_wait(slot) has returned for that slot. */
IIF_API int iif_pipeline_submit_staged_ring(iif_pipeline* p);
/* Block until the slot's latest step has delivered its loss to host_loss. */
IIF_API int iif_pipeline_wait(iif_pipeline* p, int slot);
/* Make `stream` wait for the slot's latest step (its gradients are then complete) ... */
IIF_API int iif_pipeline_stream_wait_step(iif_pipeline* p, int slot, void* stream);
/* ... and keep the slot's buffers untouched until the work queued so far on `stream` (the
 * all-reduce of its gradients) has finished. */
IIF_API int iif_pipeline_hold_slot(iif_pipeline* p, int slot, void* stream);
IIF_API int iif_pipeline_sync(iif_pipeline* p);
IIF_API void iif_pipeline_destroy(iif_pipeline* p);

/* ---------------------------------------------------------------------------------------------
 * (e) all-reduce(mean) of the flat fp32 gradient buffer [dW | db] over NVLink peer memory
 * (replaces the DDP bucket all-reduce of fc.weight.grad / fc.bias.grad: cls/train.py:231-234,
 * seg/mmdet/apis/train.py:81-85; mean over ranks = DDP semantics).
 *   peer_bufs_dev  : DEVICE array [world] of the ranks' buffer base pointers (symmetric memory: every
 *                    rank's buffer is mapped into every process; same offset on every rank)
 *   peer_flags_dev : DEVICE array [world] of the ranks' flag arrays, iif_allreduce_flag_bytes() each,
 *                    ZERO when first used (monotonic launch counters; never reset them afterwards)
 *   multicast_ptr  : NVLS multicast mapping of the buffer, or NULL (then plain peer loads / stores)
 *   offset_elems, n_elems : the slice of the buffer to reduce, both multiples of 4
 *   lane           : all-reduces of one rank that may be IN FLIGHT AT THE SAME TIME (different streams)
 *                    must use different lanes (separate flag sets); calls on one lane must be stream-ordered
 * Collective: every rank must launch it with the same arguments (its own rank aside), stream-ordered
 * after the rank's own writes to the slice.  In place; the result is bit-identical on every rank.
 * ------------------------------------------------------------------------------------------- */
IIF_API int iif_allreduce_mean_f32(void* const* peer_bufs_dev, void* const* peer_flags_dev, void* multicast_ptr,
                                   int rank, int world, int64_t offset_elems, int64_t n_elems, int num_ctas,
                                   int num_threads /* 0 = defaults */, int lane /* 0..3 */, void* stream);
IIF_API size_t iif_allreduce_flag_bytes(void);
/* Bound of the kernel's waits on OTHER ranks, wall-clock milliseconds (default 600 000, or the environment
 * variable IIF_B200_PEER_TIMEOUT_S at first use; 0 = wait forever).  Past it the kernel reports and traps. */
IIF_API int iif_allreduce_set_timeout_ms(int64_t ms);

#ifdef __cplusplus
}
#endif
#endif /* IIF_B200_H_ */
