/*
 * pds_b200.h -- C-ABI of the B200-native stereo cost-volume pipeline that sits
 * behind practical_deep_stereo.network.PdsNetwork.forward.
 *
 * The reference has NO FFI: its only seam is Python nn.Module duck typing
 * (network.py:17-24, matching.py:17-29).  This header is therefore the surface
 * a maintainer would bind from the reference's Python modules (ctypes stub in
 * INTEGRATION.md); each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     its name ends in _host; tensors are contiguous, PyTorch layout
 *     (NCHW / NCDHW), float32 unless a dtype argument says otherwise;
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and
 *     returns without synchronising; no call allocates device memory except
 *     the *_create functions (weights in kernel layout) -- scratch memory is
 *     provided by the caller through (workspace, workspace_bytes);
 *   - return value: PDS_OK or an error code; pds_last_error() gives the
 *     message for the calling thread (the Python layer raises ValueError /
 *     RuntimeError from it, mirroring network.py:28-31, estimator.py:34-41).
 */
#ifndef PDS_B200_H_
#define PDS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDS_B200_VERSION 100

enum pds_status {
  PDS_OK = 0,
  PDS_ERR_INVALID_ARGUMENT = 1, /* bad shape / parameter (-> ValueError)      */
  PDS_ERR_CUDA = 2,             /* a CUDA runtime / driver call failed         */
  PDS_ERR_WORKSPACE = 3,        /* workspace too small or misaligned           */
  PDS_ERR_UNSUPPORTED = 4       /* valid request this build cannot serve       */
};

enum pds_dtype { PDS_F32 = 0, PDS_BF16 = 1 };

/* Arithmetic of the convolution stacks.
 *   PDS_PRECISION_FP32      CUDA-core FFMA, fp32 accumulate (bit-comparable to
 *                           the reference up to summation order)
 *   PDS_PRECISION_BF16X3    tcgen05 tensor cores, operands split in three bf16
 *                           terms (24 significand bits), 6 partial products
 *   PDS_PRECISION_BF16X2    two bf16 terms (16 bits), 3 partial products
 *   PDS_PRECISION_BF16      plain bf16 operands, fp32 accumulate
 *   PDS_PRECISION_FP16X2    two IEEE-half terms (22 significand bits), 3 partial
 *                           products: fp32-grade results at half the bf16x3 cost
 *   PDS_PRECISION_FP16      plain half operands, fp32 accumulate
 * In every tensor-core mode accumulation is fp32, one accumulator per order of
 * magnitude of the partial products, summed in fp32 in the epilogue.         */
enum pds_precision {
  PDS_PRECISION_FP32 = 0,
  PDS_PRECISION_BF16X3 = 1,
  PDS_PRECISION_BF16X2 = 2,
  PDS_PRECISION_BF16 = 3,
  PDS_PRECISION_FP16X2 = 4,
  PDS_PRECISION_FP16 = 5
};

int pds_version(void);
const char* pds_status_string(int status);
const char* pds_last_error(void);

/* ---- measurement hooks (bench.py) -----------------------------------------
 * pds_launch_count: kernels launched by this library since load.
 * Profiler: when enabled, every kernel launch is bracketed by CUDA events on
 * its own stream; pds_profiler_read(i, ...) returns the i-th kernel class
 * (name, launches, summed device milliseconds), 0 when i is out of range.   */
unsigned long long pds_launch_count(void);
void pds_profiler_enable(int on);
void pds_profiler_reset(void);
int pds_profiler_read(int index, char* name, int name_len,
                      unsigned long long* launches, double* milliseconds);
/* Algorithmic work of the same kernel class, summed over its launches: FLOPs
 * counted as the reference's dense contraction (2 * taps * Cin * Cout * outputs,
 * whatever the executed precision) and the minimum HBM bytes (each input read
 * once, each output written once).                                            */
int pds_profiler_read_work(int index, double* flops, double* bytes);

/* ---- a1: Matching.forward, data movement (matching.py:50-63) -------------
 * Builds, for every disparity d in [0, D), the tensor the reference hands to
 * `operation`: cat[left, shift_d(right)] with shift_d(right)[x] = right[x-d]
 * and zeros for x < d (matching.py:12-13,56-60).
 *   left, right : (B, C, H, W)
 *   volume      : (B, D, 2C, H, W)  == (B*D, 2C, H, W) batched operation input
 */
int pds_matching_concat(const void* left, const void* right, void* volume,
                        int B, int C, int H, int W, int D, int dtype,
                        void* stream);

/* th.stack(matching_signatures, dim=2) (matching.py:63) for a batched
 * operation output: in (B*D, F, H, W) -> out (B, F, D, H, W). */
int pds_matching_stack(const void* in, void* out, int B, int F, int D, int H,
                       int W, int dtype, void* stream);

/* f4 (training): adjoints of the two kernels above, i.e. what autograd
 * accumulates through the reference's per-disparity pad / slice / cat nodes
 * (matching.py:53-60) and through th.stack (matching.py:63):
 *   grad_left [b,c,y,x] = sum_d grad_volume[b,d,c,y,x]
 *   grad_right[b,c,y,x] = sum_{d: x+d<W} grad_volume[b,d,C+c,y,x+d]
 * summed over d in ascending order in fp32 (deterministic).
 * pds_matching_unstack: in (B, F, D, H, W) -> out (B*D, F, H, W). */
int pds_matching_concat_backward(const void* grad_volume, void* grad_left,
                                 void* grad_right, int B, int C, int H, int W,
                                 int D, int dtype, void* stream);
int pds_matching_unstack(const void* in, void* out, int B, int F, int D, int H,
                         int W, int dtype, void* stream);

/* f4 (training): InstanceNorm{2,3}d(affine, eps)(LeakyReLU_slope(x)) -- the
 * tail of every Conv -> LeakyReLU -> InstanceNorm block (network_blocks.py:
 * 47-58, 88-131) -- forward and backward on fp32 (N, C, L) tensors (L = product
 * of the spatial extents), slope = 1 for a plain InstanceNorm; gamma / beta may
 * be null (affine=False).
 *   forward : y; mean_rstd (N*C, 2) float is kept for the backward; sums is
 *             (N*C, 2) double scratch.
 *   backward: dx; on return sums (N*C, 2) double holds, per (sample, channel),
 *             sum dy and sum dy * zhat: d beta / d gamma are their sums over
 *             the samples. */
int pds_instance_norm_forward(const float* x, const float* gamma,
                              const float* beta, float* y, float* mean_rstd,
                              double* sums, int N, int C, long long L,
                              float eps, float slope, void* stream);
int pds_instance_norm_backward(const float* x, const float* dy,
                               const float* gamma, const float* mean_rstd,
                               float* dx, double* sums, int N, int C,
                               long long L, float slope, void* stream);

/* ---- a2: MatchingOperation.forward over all disparities ------------------
 * (matching.py:69-112 applied by the loop of matching.py:53-62.)
 * Weights are given in the reference's own layout, in state_dict() order of
 * MatchingOperation: conv0.w (F,2C,3,3), conv0.b, then per residual block
 * 2 x [conv.w (F,F,3,3), conv.b, in.gamma, in.beta], then conv_last.w
 * (S,F,3,3), conv_last.b; all float32 device pointers, copied and re-laid-out
 * at create time.                                                            */
typedef struct pds_matching_op pds_matching_op;

int pds_matching_op_create(pds_matching_op** op, const float* const* params,
                           int n_params, int descriptor_features /* C=64 */,
                           int features /* F=64 */, int signature_features /* 8 */,
                           int residual_blocks /* 2 */, int precision,
                           void* stream);
void pds_matching_op_destroy(pds_matching_op* op);
size_t pds_matching_op_workspace_bytes(const pds_matching_op* op, int B, int H,
                                       int W, int D);
/* left, right (B, C, H, W) -> signatures (B, S, D, H, W), D = md + 1. */
int pds_matching_op_forward(pds_matching_op* op, const float* left,
                            const float* right, float* signatures, int B,
                            int H, int W, int D, void* workspace,
                            size_t workspace_bytes, void* stream);

/* ---- a3: Regularization.forward (regularization.py:94-126) ---------------
 * params: state_dict() order of Regularization (74 tensors for F = 8).       */
typedef struct pds_regularization pds_regularization;

int pds_regularization_create(pds_regularization** reg,
                              const float* const* params, int n_params,
                              int features /* 8 */, int precision,
                              void* stream);
void pds_regularization_destroy(pds_regularization* reg);
size_t pds_regularization_workspace_bytes(const pds_regularization* reg, int B,
                                          int D, int H, int W);
/* signatures (B, F, D, H, W), shortcut (B, F, H, W) -> cost (B, 2D, 4H, 4W) */
int pds_regularization_forward(pds_regularization* reg,
                               const float* signatures, const float* shortcut,
                               float* cost, int B, int D, int H, int W,
                               void* workspace, size_t workspace_bytes,
                               void* stream);

/* Regularization.forward + SubpixelMap.__call__ + SizeAdapter.unpad in one call
 * (regularization.py:125-126 -> estimator.py:59-91 -> size_adapter.py:51-52, the
 * tail of PdsNetwork.forward in eval mode, network.py:49-52): the last layer of
 * the hourglass feeds the estimator's running arg-max state directly and the
 * (B, 2D, 4H, 4W) cost volume is never written.  Results are bit-identical to
 * pds_regularization_forward followed by pds_subpixel_map.
 * disparity (B, 4H - crop_top, 4W - crop_left) float32; argmax may be NULL.
 * half_support_window / disparity_step as in pds_subpixel_map (window radius
 * hsw / step <= 4, else PDS_ERR_UNSUPPORTED).                                 */
int pds_regularization_forward_disparity(pds_regularization* reg,
                                         const float* signatures,
                                         const float* shortcut, float* disparity,
                                         int64_t* argmax, int B, int D, int H,
                                         int W, int half_support_window,
                                         int disparity_step, int crop_top,
                                         int crop_left, void* workspace,
                                         size_t workspace_bytes, void* stream);

/* Individually tested reference blocks (test/test_regularization.py:13-28):
 * ContractionBlock3d.forward (regularization.py:28-31) and
 * ExpansionBlock3d.forward (regularization.py:54-57); params = the block's
 * state_dict() order (8 tensors each), fp32 CUDA-core arithmetic.            */
size_t pds_contraction_block_workspace_bytes(int B, int C, int D, int H, int W);
int pds_contraction_block_forward(const float* const* params, const float* in,
                                  float* down, float* smooth, int B, int C,
                                  int D, int H, int W, void* workspace,
                                  size_t workspace_bytes, void* stream);
size_t pds_expansion_block_workspace_bytes(int B, int C, int D, int H, int W);
int pds_expansion_block_forward(const float* const* params, const float* in,
                                const float* skip, float* out, int B, int C,
                                int D, int H, int W, void* workspace,
                                size_t workspace_bytes, void* stream);

/* ---- a6: Embedding.forward (embedding.py:46-65) ----------------------------
 * InstanceNorm2d(3) -> 2 x [conv5x5 s2 + LeakyReLU + IN] -> residual blocks ->
 * descriptor; shortcut = conv3x3 block of the descriptor.  params: state_dict()
 * order of Embedding (28 tensors by default).  Tensor-core precisions only
 * (PDS_ERR_UNSUPPORTED for PDS_PRECISION_FP32: callers keep their own fp32 path).
 * images (n, in_features, H, W), H and W multiples of 4 -> descriptor
 * (n, features, H/4, W/4); shortcut (n_shortcut, shortcut_features, H/4, W/4) is
 * computed for the first n_shortcut samples only (the reference computes and
 * discards it for the right image, network.py:40); shortcut may be NULL when
 * n_shortcut == 0.  Left and right images are simply samples of one batch.     */
typedef struct pds_embedding pds_embedding;

int pds_embedding_create(pds_embedding** emb, const float* const* params,
                         int n_params, int in_features /* 3 */,
                         int features /* 64 */, int shortcut_features /* 8 */,
                         int residual_blocks /* 2 */, int precision,
                         void* stream);
void pds_embedding_destroy(pds_embedding* emb);
size_t pds_embedding_workspace_bytes(const pds_embedding* emb, int n, int H,
                                     int W);
int pds_embedding_forward(pds_embedding* emb, const float* images,
                          float* descriptor, float* shortcut, int n,
                          int n_shortcut, int H, int W, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---- f3: the input path in front of the embedding ---------------------------
 * The images as the caller holds them, UN-padded: dataset.py:67-72 decodes to
 * interleaved uint8 (cv2) and converts to float CHW on the host; this entry
 * takes either form.  SizeAdapter.pad (size_adapter.py:29-43: pad_top rows and
 * pad_left columns of zeros, padded extent (h + pad_top) x (w + pad_left), both
 * multiples of 4) and the first InstanceNorm2d (embedding.py:32, statistics
 * over the padded image) are fused into the kernel that writes the first
 * convolution's operands.  The batch is images_a (n_a samples) followed by
 * images_b (n_b samples; may be NULL with n_b == 0): left and right images
 * without a concatenation copy.  Outputs and workspace as pds_embedding_forward
 * with n = n_a + n_b and the PADDED extent.                                   */
enum pds_image_layout {
  PDS_IMAGE_F32_NCHW = 0, /* float32 (n, C, h, w)                     */
  PDS_IMAGE_U8_NCHW = 1,  /* uint8   (n, C, h, w)                     */
  PDS_IMAGE_U8_NHWC = 2   /* uint8   (n, h, w, C): cv2 / decoder order */
};
int pds_embedding_forward_images(pds_embedding* emb, const void* images_a,
                                 int n_a, const void* images_b, int n_b,
                                 int layout, int h, int w, int pad_top,
                                 int pad_left, float* descriptor,
                                 float* shortcut, int n_shortcut,
                                 void* workspace, size_t workspace_bytes,
                                 void* stream);

/* ---- a4: SubpixelMap.__call__ (estimator.py:45-91) -----------------------
 * cost (B, D, H, W) of `dtype` -> disparity (B, H-crop_top, W-crop_left)
 * float32; the crop is SizeAdapter.unpad (size_adapter.py:51-52) fused into
 * the store.  argmax may be NULL, else (B, H-crop_top, W-crop_left) int64
 * (th.max indices: lowest index on ties, first NaN wins).
 * half_support_window >= 1, disparity_step >= 1 and hsw % step == 0, else
 * PDS_ERR_INVALID_ARGUMENT (estimator.py:34-41).                             */
int pds_subpixel_map(const void* cost, float* disparity, int64_t* argmax,
                     int B, int D, int H, int W, int half_support_window,
                     int disparity_step, int crop_top, int crop_left,
                     int dtype, void* stream);

/* ---- f4 (evaluation half): errors.compute_absolute_error and
 * errors.compute_n_pixels_error (errors.py:9-74) in one pass --------------------
 * estimated / ground_truth: `count` float32 disparities (any shape, contiguous);
 * ground truth +-inf marks "unknown" (dataset.py:74-81).  pixelwise_abs and
 * pixelwise_n_pixels (each NULL or `count` floats) receive |e - g| resp.
 * [|e - g| > n] with 0 where the ground truth is unknown.  sums (device, 3
 * doubles): sum of |e - g| over known locations, number of known locations,
 * number of known locations with |e - g| > n -- the mean absolute error is
 * sums[0] / sums[1], the n-pixel error 100 * sums[2] / sums[1] (0 when sums[1]
 * is 0, as the reference returns).                                             */
int pds_disparity_errors(const float* estimated, const float* ground_truth,
                         float* pixelwise_abs, float* pixelwise_n_pixels,
                         size_t count, float n, double* sums, void* stream);

/* ---- f4 (training half): loss.SubpixelCrossEntropy (loss.py:16-78) -------------
 * The reference loops over the disparity axis in Python (log_softmax of the whole
 * volume, then per index an un-normalised Laplace target and two accumulations);
 * here the forward is ONE pass over similarities (B, D, H, W) float32 contiguous,
 * the backward another one.  ground_truth (B, H, W): +-inf marks "unknown"
 * locations (they contribute nothing); weights: NULL or (B, H, W).
 * Forward outputs per location: entropy (0 where unknown), lse = log-sum-exp of
 * the similarities, sum_pt = sum_d P_target(d); sums (device, 2 doubles):
 * sum of w * entropy and sum of w over known locations (w = 1 without weights).
 * The loss is sums[0] / (sums[1] + 1e-15) with weights, sums[0] / sums[1]
 * without (loss.py:74-78).
 * Backward: upstream = d L / d loss (device, one float);
 * grad_similarities (B, D, H, W) = upstream * c * (softmax(s)_d - P_target(d) /
 * sum_pt), c = w / (sum w + 1e-15) resp. 1 / N; grad_weights: NULL or (B, H, W) =
 * upstream * (entropy - loss) / (sum w + 1e-15).                                  */
int pds_subpixel_cross_entropy_forward(const float* similarities, const float* ground_truth,
                                       const float* weights, float* entropy, float* lse, float* sum_pt,
                                       double* sums, int B, int D, int H, int W, float diversity,
                                       int disparity_step, void* stream);
int pds_subpixel_cross_entropy_backward(const float* similarities, const float* ground_truth,
                                        const float* weights, const float* entropy, const float* lse,
                                        const float* sum_pt, const double* sums, const float* upstream,
                                        float* grad_similarities, float* grad_weights, int B, int D,
                                        int H, int W, float diversity, int disparity_step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDS_B200_H_ */
